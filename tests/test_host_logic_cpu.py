"""CPU: the HOST side of the product (autograd wiring in ops.py, grouping / batching / index logic in
model.py, the loop body in trainer.py, state_dict compatibility) against the golden fixtures of the real
reference.  The CUDA kernels are replaced by their torch statements from tests/emul.py — this validates
everything except the kernels themselves, which the -m gpu tests cover through the C ABI."""
import json
import os

import pytest
import torch

from tests.conftest import GOLDEN, load_golden
from tests.helpers import golden_state, golden_inputs, digest_close, template_state
from tests import emul
import rd_b200.config as rd_config
from rd_b200.trainer import Trainer, build_model, default_active, apply_fix_pretrain, LOSS_KEYS


@pytest.fixture()
def emulated():
    saved = emul.install()
    yield
    emul.uninstall(saved)


def _cfg_from_fixture(fx):
    cfg = rd_config.default_config(precision="fp32")
    cfg.update(fx["cfg"])
    cfg["precision"] = "fp32"
    cfg = rd_config.derive(cfg)
    for k in ("input_output_act", "target_output_act"):     # fixtures pin the constructor arguments themselves
        if k in fx["cfg"]:
            cfg[k] = fx["cfg"][k]
    return cfg


def test_state_dict_matches_reference():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as f:
        ref = json.load(f)
    cfg = rd_config.default_config(precision="fp32")
    model = build_model(cfg, "cpu")
    sd = model.state_dict()
    assert list(sd.keys()) == [e["key"] for e in ref["keys"]]
    for e in ref["keys"]:
        assert list(sd[e["key"]].shape) == e["shape"], e["key"]
    names = [n for n, _ in model.named_parameters()]
    assert names == [e["key"] for e in ref["keys"] if e["param"]]
    # initialisation follows the same distributions (xavier_normal_ CondConv weights, zero CondConv bias, ...)
    for n, p in model.named_parameters():
        st = ref["init_stats"][n]
        if p.numel() >= 4096:
            assert abs(float(p.std()) - st["std"]) <= 0.1 * st["std"] + 1e-6, n
        if n.endswith(".bias") and st["std"] == 0.0 and p.numel() > 1:
            assert float(p.abs().max()) == 0.0, n


def test_state_dict_m2():
    with open(os.path.join(GOLDEN, "state_dict_keys_m2.json")) as f:
        ref = json.load(f)
    cfg = rd_config.default_config(precision="fp32", contrast_list=["T1", "T2"], dataset_name="NCANDA")
    model = build_model(cfg, "cpu")
    assert list(model.state_dict().keys()) == [e["key"] for e in ref["keys"]]


def _run_step(fx_name, emulated_unused=None):
    fx = load_golden(fx_name + ".pt")
    cfg = _cfg_from_fixture(fx)
    model = build_model(cfg, "cpu")
    model.load_state_dict(golden_state(fx, model))
    model.train(fx["training"])
    apply_fix_pretrain(model, cfg)
    tr = Trainer(model, cfg, fx["B"], use_graph=False)
    batch, eps = golden_inputs(fx)
    tr.load_batch(batch, eps, tuple(fx["pair"]))
    return fx, cfg, model, tr


# fixtures whose masks / freezing leave MORE parameters without a gradient than the name-based static rule (default_active) says:
# an all-missing contrast whose decoder half no counted loss term reads (Q4 / Q10), the frozen stage-1 networks (fix_pretrain)
DATA_DEPENDENT_NONE = ("step_m4_b2_skip", "stage2_fused_zd_b1", "stage2_fused_brats_b2")


@pytest.mark.parametrize("name", ["step_m4_b2", "step_m2_b2", "stage2_m4_b2", "variants_m4_b2", "shared_m4_b2", "stage2_u_m4_b2",
                                  "step_m4_b2_skip", "step_m4_b2_kl_p2", "stage2_fused_zd_b1", "stage2_fused_brats_b2",
                                  "stage2_saca_m4_b2", "stage2_ssaca_m4_b2"])
def test_train_iteration_matches_reference(emulated, name):
    fx, cfg, model, tr = _run_step(name)
    out = tr.forward_losses(with_y=fx["with_y"], keep=True)
    L = out["losses"]
    for k, v in fx["losses"].items():
        assert abs(float(L[k]) - v) <= 2e-4 * max(1.0, abs(v)), (k, float(L[k]), v)
    T = out["tensors"]
    B, M = fx["B"], fx["M"]
    g = fx["tensors"]
    for i in range(M):
        digest_close(T["S"][i * B:(i + 1) * B].permute(0, 3, 1, 2), g["si"][i], 2e-3, 1e-5, "si[%d]" % i)
        digest_close(T["z_mean"][i * B:(i + 1) * B], g["z_mean"][i], 2e-3, 1e-5, "z_mean[%d]" % i)
        digest_close(T["x_fake"][i * B:(i + 1) * B].permute(0, 3, 1, 2), g["x_fake"][i], 2e-3, 1e-5, "x_fake[%d]" % i)
    for t in range(M * (M - 1)):
        digest_close(T["x_fake_mix"][t * B:(t + 1) * B].permute(0, 3, 1, 2), g["x_fake_mix"][t], 2e-3, 1e-5, "x_mix[%d]" % t)
    if fx["with_y"]:
        for i in range(M):
            digest_close(T["y_fake_list"][i * B:(i + 1) * B].permute(0, 3, 1, 2), g["y_fake_list"][i], 2e-3, 1e-5, "y[%d]" % i)
        digest_close(T["y_fake_fused"].permute(0, 3, 1, 2), g["y_fake_fused"], 2e-3, 1e-5, "y_fused")
    # backward + clip exactly as the trainer does it
    L["all"].backward()
    fp = tr.fp
    import rd_b200.kernels as K
    K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
    assert abs(float(fp.scalars[0]) - fx["grad_norm"]) <= 2e-3 * fx["grad_norm"]
    K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)
    active = dict(zip(fp.names, fp.active_mask))
    for (n, p) in model.named_parameters():
        d = fx["grads"][n]
        if name in DATA_DEPENDENT_NONE:
            assert d is None or (active[n] and p.requires_grad), "the reference has a gradient for the inactive parameter " + n
        else:
            assert (d is not None) == active[n], "active-parameter rule differs from the reference for " + n
        if d is not None:
            digest_close(p.grad, d, 5e-3, 2e-7, "grad:" + n)
        else:
            assert float(p.grad.abs().max()) == 0.0, n
    sd = model.state_dict()
    for k, d in fx["buffers"].items():
        digest_close(sd[k], d, 1e-3, 1e-6, "buf:" + k)


def test_full_reference_tensors_every_pixel(emulated):
    """Host logic against FULL tensors of the unmodified reference (tests/golden/step_m4_b2_full_tensors.pt: s_0, x-hat_0 and the first
    cross-reconstruction of the first slice), every element — not only the digests the other fixtures keep."""
    fx, cfg, model, tr = _run_step("step_m4_b2_full")
    ref = load_golden("step_m4_b2_full_tensors.pt")
    with torch.no_grad():
        T = tr.forward_losses(keep=True)["tensors"]
    got = {"si0": T["S"][0:1], "x_fake0": T["x_fake"][0:1], "x_fake_mix0": T["x_fake_mix"][0:1]}
    for k, r in ref.items():
        a = got[k].permute(0, 3, 1, 2).float()
        assert a.shape == r.shape
        assert float((a - r).abs().max()) <= 2e-4 * float(r.abs().max()), k


def test_inference_matches_reference(emulated):
    fx, cfg, model, tr = _run_step("infer_m4_b2")
    with torch.no_grad():
        out = tr.forward_losses(with_y=True, keep=True)
    for k, v in fx["losses"].items():
        assert abs(float(out["losses"][k]) - v) <= 2e-4 * max(1.0, abs(v)), (k, float(out["losses"][k]), v)
    digest_close(out["tensors"]["y_fake_fused"].permute(0, 3, 1, 2), fx["tensors"]["y_fake_fused"], 2e-3, 1e-5, "y_fused")


@pytest.mark.parametrize("dedup", [False, True])
def test_batched_inference_sweep_equals_per_subset_passes(emulated, dedup):
    """Host logic of rd_b200.inference.SweepRunner (CPU twin of the GPU test): missing-modality subsets as ONE batched pass — optionally only
    the distinct rows — against the reference's schedule, one pass per subset over its present contrasts (src/main_missing.py:349,
    src/util.py:580-613, src/model.py:3135-3157, 3239-3258).  A subset list with unequal group sizes also walks the group-padding branch."""
    import rd_b200.kernels as K
    import rd_b200.ops as ops_mod
    from rd_b200.inference import SweepRunner
    fx, cfg, model, tr = _run_step("infer_m4_b2")
    batch, _ = golden_inputs(fx)
    model.eval()
    B, M, C = fx["B"], fx["M"], model.in_num_ch
    subsets = [1, 6, 10, 11, 15]                # {0}, {1,2}, {1,3}, {0,1,3}, all: contrast 1 in four subsets, contrast 2 in two
    sw = SweepRunner(model, B, use_graph=False, subsets=subsets, dedup=dedup)
    sw.load(batch["inputs"], batch["mask_img"])
    out = sw.sweep()
    assert out.shape[0] == sum(bin(s).count("1") for s in subsets) * B
    inputs, mask_img = batch["inputs"].float(), batch["mask_img"].float()
    ones_img = torch.ones_like(mask_img)
    worst = 0.0
    with torch.no_grad():
        for k, sub in enumerate(sw.subsets):
            present = [m for m in range(M) if (sub >> m) & 1]
            r = len(present)
            X = torch.empty((r * B, model.input_size[0], model.input_size[1], C), dtype=model.cdtype)
            for q, m in enumerate(present):
                K.nchw_to_nhwc(inputs, X[q * B:(q + 1) * B], m * C, C)
            types = [model._types_all[m] for m in present]
            feats = model.anatomy_encoder_enc_list[0].nhwc(X, types)
            logits = model.anatomy_encoder_dec.nhwc(feats, types)
            mi = mask_img if 0 in present else ones_img
            S = ops_mod.masked_softmax(logits, mi if model.others.get("softmax_remove_mask", False) else None)
            rows, _, _ = ops_mod.fuse_gather(S, torch.ones(B, r), B, r)
            y, _ = model.output_decoder.nhwc(model.fuse_rows(rows))
            got = sw.subset_output(k)
            assert got.shape == y.shape, (got.shape, y.shape)
            worst = max(worst, float((got.float() - y.float()).abs().max()) / max(float(y.float().abs().max()), 1e-30))
    assert worst <= 1e-5, worst


def test_list_api_matches_stacked_path(emulated):
    """The reference-signature list methods (what main_missing.py calls) give the same numbers as the trainer."""
    fx, cfg, model, tr = _run_step("step_m4_b2")
    B, M, C = fx["B"], fx["M"], 7
    with torch.no_grad():
        xs = [tr.inputs[:, m * C:(m + 1) * C] for m in range(M)]
        si = model.compute_anatomy_encoding(xs, tr.mask_img)
        model._eps_override = tr.eps
        zi, mu, lv = model.compute_modality_encoding(xs, si, phase="train")
        xf = model.reconstruct_input_si_zi(si, zi)
        xm = model.reconstruct_input_si_zj(si, zi)
        lx = model.compute_recon_loss_x_list(xs, xf, tr.mask, p=1)
        lm = model.compute_recon_loss_x_mix_list(xs, xm, tr.mask, p=1)
        model._pair_override = tuple(fx["pair"])
        ls = model.compute_similarity_s_loss(si, tr.mask)
        lz = model.compute_similarity_z_loss(zi, tr.mask)
    g = fx["tensors"]
    assert tuple(si[0].shape) == (B, 4, 160, 192)
    for i in range(M):
        digest_close(si[i], g["si"][i], 2e-3, 1e-5, "si")
        digest_close(xf[i], g["x_fake"][i], 2e-3, 1e-5, "x_fake")
    for k, v in (("recon_x", lx), ("recon_x_mix", lm), ("sim_s", ls), ("sim_z", lz)):
        assert abs(float(v) - fx["losses"][k]) <= 2e-4 * max(1.0, abs(fx["losses"][k])), k


def test_optimizer_matches_torch_adam(emulated):
    """clip + Adam(amsgrad, wd 1e-5) over the flat buffers == torch.optim.Adam on the same grads, 3 steps."""
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(37, 5)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3, 3))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = torch.optim.Adam([ref[0], ref[2]], lr=2e-4, weight_decay=1e-5, amsgrad=True)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = ps
    from rd_b200.trainer import FlatParams
    import rd_b200.kernels as K
    fp = FlatParams(Holder())
    fp.set_active([True, False, True])
    hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 0, 0])
    for step in range(3):
        gs = [torch.randn_like(p) * 3 for p in ps]
        for p, r, g in zip(ps, ref, gs):
            p.grad.copy_(g)
            r.grad = g.clone()
        ref[1].grad = None
        total = torch.nn.utils.clip_grad_norm_([ref[0], ref[2]], 1.0)
        opt.step()
        ps[1].grad.zero_()
        K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
        K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)
        K.adam_amsgrad(fp.flat, fp.grad, fp.m, fp.v, fp.vmax, fp.segments, fp.nseg, hyper)
        assert abs(float(fp.scalars[0]) - float(total)) < 1e-4
    for p, r in zip(ps, ref):
        assert torch.allclose(p.detach(), r.detach(), rtol=1e-5, atol=1e-6)


def test_bf16_channel_padding_wiring(emulated):
    """bf16 mode zero-pads 4- / 7-channel tensors to 8 for the tensor-core gathers; with the kernels emulated in
    torch (bf16 storage, fp32 math) the step must still agree with the reference within bf16 noise."""
    fx = load_golden("step_m2_b2.pt")
    cfg = _cfg_from_fixture(fx)
    cfg["precision"] = "bf16"
    model = build_model(cfg, "cpu")
    model.load_state_dict(golden_state(fx, model))
    model.train(True)
    tr = Trainer(model, cfg, fx["B"], use_graph=False)
    batch, eps = golden_inputs(fx)
    tr.load_batch(batch, eps, tuple(fx["pair"]))
    out = tr.forward_losses(keep=True)
    L = out["losses"]
    for k in ("recon_x", "recon_x_mix", "all"):
        assert abs(float(L[k]) - fx["losses"][k]) <= 3e-2 * max(1.0, abs(fx["losses"][k])), (k, float(L[k]), fx["losses"][k])
    assert out["tensors"]["S"].shape[-1] == 4 and out["tensors"]["x_fake"].shape[-1] == 7
    L["all"].backward()
    import rd_b200.kernels as K
    fp = tr.fp
    K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
    assert abs(float(fp.scalars[0]) - fx["grad_norm"]) <= 0.1 * fx["grad_norm"]


def test_prefetch_on_cpu_loads_the_batch(emulated):
    """Without a CUDA device Trainer.prefetch is load_batch (no copy stream): the static buffers hold the staged batch."""
    fx, cfg, model, tr = _run_step("step_m4_b2")
    batch, eps = golden_inputs(fx)
    other = dict(batch)
    other["inputs"] = batch["inputs"] * 0.5
    tr.prefetch(other, eps, (1, 3))
    assert torch.equal(tr.inputs, other["inputs"].float())
    assert [int(v) for v in tr.pair] == [1, 3]
    assert torch.equal(tr.eps, torch.cat([e.reshape(fx["B"], -1) for e in eps], 0))


def test_same_size_resize_is_the_identity(emulated):
    """ops.bilinear to the tensor's own size returns its argument (src == dst, l1 == 0 in both align conventions): the full-resolution
    SPADE block's resize of the anatomy code costs nothing, forward or backward."""
    from rd_b200 import ops
    x = torch.randn(2, 5, 6, 4, requires_grad=True)
    for align in (False, True):
        assert ops.bilinear(x, 5, 6, align) is x
    y = ops.bilinear(x, 10, 12, False)
    assert tuple(y.shape) == (2, 10, 12, 4)


@pytest.mark.parametrize("name", ["step_m4_b2", "step_m2_b2"])
def test_composed_decoder_tail_matches_separate_convs(emulated, name):
    """ops.COMPOSE_OUT: sp6.out (3x3) composed with the decoder's 1x1 `out` into one grouped convolution (nothing lies between
    them, src/model.py:2606-2612) gives the reference's losses, images and EVERY parameter gradient — including both composed
    layers' experts, routing and biases through the weight-space chain rule."""
    from rd_b200 import ops
    res = []
    old_flag = ops.COMPOSE_OUT
    for flag in (False, True):
        ops.COMPOSE_OUT = flag
        try:
            fx, cfg, model, tr = _run_step(name)
            out = tr.forward_losses(with_y=fx["with_y"], keep=True)
            L = out["losses"]
            L["all"].backward()
            res.append(({k: float(v) for k, v in L.items()}, out["tensors"]["x_fake"].detach().clone(), tr.fp.grad.clone(),
                        list(tr.fp.names), list(tr.fp.offsets)))
        finally:
            ops.COMPOSE_OUT = old_flag
    (la, xa, ga, names, offs), (lb, xb, gb, _, _) = res
    for k, v in fx["losses"].items():
        assert abs(lb[k] - v) <= 2e-4 * max(1.0, abs(v)), (k, lb[k], v)
        assert abs(lb[k] - la[k]) <= 1e-5 * max(1.0, abs(la[k])), (k, la[k], lb[k])
    assert torch.allclose(xa, xb, rtol=1e-4, atol=1e-5)
    # per-parameter comparison: relative to the parameter's own gradient magnitude
    bounds = offs + [ga.numel()]
    for i, n in enumerate(names):
        a, b = ga[bounds[i]:bounds[i + 1]], gb[bounds[i]:bounds[i + 1]]
        scale = float(a.abs().max())
        assert float((a - b).abs().max()) <= 2e-4 * scale + 1e-7, (n, scale, float((a - b).abs().max()))


@pytest.mark.parametrize("name", ["step_m4_b2", "shared_m4_b2"])
def test_fused_spade_convolution_matches_separate_passes(emulated, name):
    """ops.spade_conv with the modulation fused into the gamma|beta convolution (rd_conv2d_fwd_spade; its autograd path saves gamma only
    and runs rd_spade_modulate_bwd_g before the convolution's backward) gives the reference's losses and EVERY parameter gradient, like
    the separate convolution + modulation passes."""
    res = []
    for flag in (False, True):
        emul.SPADE_FUSE = flag
        try:
            fx, cfg, model, tr = _run_step(name)
            out = tr.forward_losses(with_y=fx["with_y"], keep=True)
            L = out["losses"]
            L["all"].backward()
            res.append(({k: float(v) for k, v in L.items()}, out["tensors"]["x_fake"].detach().clone(), tr.fp.grad.clone(),
                        list(tr.fp.names), list(tr.fp.offsets)))
        finally:
            emul.SPADE_FUSE = False
    (la, xa, ga, names, offs), (lb, xb, gb, _, _) = res
    for k, v in fx["losses"].items():
        assert abs(lb[k] - v) <= 2e-4 * max(1.0, abs(v)), (k, lb[k], v)
    assert torch.allclose(xa, xb, rtol=1e-4, atol=1e-5)
    bounds = offs + [ga.numel()]
    for i, n in enumerate(names):
        a, b = ga[bounds[i]:bounds[i + 1]], gb[bounds[i]:bounds[i + 1]]
        scale = float(a.abs().max())
        assert float((a - b).abs().max()) <= 2e-4 * scale + 1e-7, (n, scale, float((a - b).abs().max()))


def test_adam_skips_parameters_without_gradient_like_torch(emulated):
    """torch.optim.Adam leaves a parameter whose grad is None untouched (no update, no moment decay, its own step counter does not
    advance).  With contrast 3 missing in every row the private decoder half input_decoder_list.3 is not reached by any counted
    loss term (fixture step_m4_b2_skip: 52 extra grad-None parameters in the reference) -> bit-unchanged after the iteration."""
    fx, cfg, model, tr = _run_step("step_m4_b2_skip")
    tr.accum_every = 1
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    batch, eps = golden_inputs(fx)
    tr.train_iteration(batch, eps, tuple(fx["pair"]))
    steps = dict(zip(tr.fp.names, tr.fp.param_steps.tolist()))
    for n, p in model.named_parameters():
        ref_none = fx["grads"][n] is None
        if ref_none:
            assert torch.equal(p.detach(), before[n]), n
            assert steps[n] == 0.0, n
        else:
            assert steps[n] == 1.0, n
            assert not torch.equal(p.detach(), before[n]), n
    sd = tr.optimizer_state_dict()
    idx = {n: k for k, n in enumerate(tr.fp.names)}
    assert idx["input_decoder_list.3.sp4.gamma.weight"] not in sd["state"]
    assert idx["input_decoder_list.2.sp4.gamma.weight"] in sd["state"]


def test_optimizer_state_round_trips_with_torch_adam(emulated):
    """export_adam_state / import_adam_state against torch.optim.Adam(amsgrad=True, weight_decay=1e-5): same per-parameter step /
    exp_avg / exp_avg_sq / max_exp_avg_sq after steps in which one parameter has grad None, torch loads the exported dict and both
    continue identically; importing torch's dict reproduces the flat buffers."""
    from rd_b200.trainer import FlatParams, export_adam_state, import_adam_state
    import rd_b200.kernels as K
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(9, 4)), torch.nn.Parameter(torch.randn(6)), torch.nn.Parameter(torch.randn(2, 3))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = torch.optim.Adam(ref, lr=2e-4, weight_decay=1e-5, amsgrad=True)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = ps
    fp = FlatParams(Holder())
    fp.set_active([True, True, True])
    hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 1 - 0.9, 1 - 0.999])

    def step(skip_b):
        gs = [torch.randn_like(p) for p in ps]
        for k, (p, r, g) in enumerate(zip(ps, ref, gs)):
            if k == 1 and skip_b:
                p.grad.zero_()
                r.grad = None
            else:
                p.grad.copy_(g)
                r.grad = g.clone()
        opt.step()
        K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1e9)
        K.clip_adam_amsgrad_gated(fp.flat, fp.grad, fp.m, fp.v, fp.vmax, fp.segments, fp.seg_param, fp.nseg, fp.partial, fp.param_flags,
                                  fp.param_steps, hyper, None, True)
    step(True)
    step(False)
    assert fp.param_steps.tolist() == [2.0, 1.0, 2.0]
    for p, r in zip(ps, ref):
        assert torch.allclose(p.detach(), r.detach(), rtol=1e-6, atol=1e-7)
    sd, tsd = export_adam_state(fp, hyper), opt.state_dict()
    assert sorted(sd["state"].keys()) == sorted(tsd["state"].keys())
    for k in tsd["state"]:
        assert float(sd["state"][k]["step"]) == float(tsd["state"][k]["step"])
        for f in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            assert torch.allclose(sd["state"][k][f], tsd["state"][k][f], rtol=1e-5, atol=1e-9), (k, f)
    opt2 = torch.optim.Adam([torch.nn.Parameter(r.detach().clone()) for r in ref], lr=1.0, amsgrad=True)
    opt2.load_state_dict(sd)                      # torch accepts the exported dict as is
    assert opt2.param_groups[0]["lr"] == pytest.approx(2e-4) and opt2.param_groups[0]["weight_decay"] == pytest.approx(1e-5)
    # import torch's dict into fresh flat buffers
    qs = [torch.nn.Parameter(r.detach().clone()) for r in ref]

    class Holder2(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = qs
    fp2 = FlatParams(Holder2())
    fp2.set_active([True, True, True])
    hyper2 = torch.zeros(8)
    import_adam_state(fp2, hyper2, tsd)
    assert torch.allclose(fp2.m, fp.m) and torch.allclose(fp2.v, fp.v) and torch.allclose(fp2.vmax, fp.vmax)
    assert fp2.param_steps.tolist() == [2.0, 1.0, 2.0] and float(hyper2[0]) == pytest.approx(2e-4)


def test_ragged_last_batch_and_shape_checks(emulated):
    """The reference's loaders keep a smaller final batch (src/util.py:706).  A trainer built for B rows runs it eagerly on its own
    rows (same losses as a trainer built for that size) instead of broadcasting it into the B-row buffers; wrong shapes raise."""
    fx, cfg, model, tr = _run_step("step_m4_b2")
    batch, eps = golden_inputs(fx)
    one = {k: (v[:1].clone() if torch.is_tensor(v) else v[:1]) for k, v in batch.items()}
    eps1 = [e[:1].clone() for e in eps]
    tr.accum_every = 10 ** 6                               # no optimizer step: compare the losses of the same weights
    lv = tr.train_iteration(one, eps1, tuple(fx["pair"])).clone()
    assert tr.B == fx["B"] and tuple(tr.inputs.shape) == (fx["B"], 28, 160, 192) and tr.iter == 1
    model2 = build_model(cfg, "cpu")
    model2.load_state_dict(golden_state(fx, model2))
    tr2 = Trainer(model2, cfg, 1, use_graph=False)
    tr2.accum_every = 10 ** 6
    lv2 = tr2.train_iteration(one, eps1, tuple(fx["pair"]))
    assert torch.allclose(lv, lv2, rtol=1e-5, atol=1e-6), (lv, lv2)
    bad = dict(batch)
    bad["inputs"] = batch["inputs"][:, :21]
    with pytest.raises(ValueError):
        tr.load_batch(bad, eps, (0, 1))
    three = {k: (torch.cat([v, v[:1]], 0) if torch.is_tensor(v) else v) for k, v in batch.items()}
    with pytest.raises(ValueError):
        tr.load_batch(three, eps, (0, 1))
