"""CPU: the oracle restatement (oracle/rd_oracle.py) against the fixtures written by the REAL reference
(oracle/make_golden.py, run where /root/reference exists).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden
from tests.helpers import golden_state, golden_inputs, digest_close
from oracle.rd_oracle import RDOracle, clone_state, param_keys, clip_grad_norm


def _run_case(name, batched):
    fx = load_golden(name + ".pt")
    state = clone_state(golden_state(fx))
    batch, eps = golden_inputs(fx)
    orc = RDOracle(state, fx["cfg"], training=fx["training"], batched_condconv=batched)
    ctx = torch.enable_grad() if fx["training"] else torch.no_grad()
    with ctx:
        out = orc.forward_losses(batch["inputs"], batch["targets"], batch["mask"], batch["mask_img"], eps,
                                 tuple(fx["pair"]), with_y=fx["with_y"], keep=True)
    for k, v in fx["losses"].items():
        assert abs(float(out[k]) - v) <= 1e-4 * max(1.0, abs(v)), (k, float(out[k]), v)
    T = out["tensors"]
    for k, d in fx["tensors"].items():
        if d is None:
            assert T[k] is None
        elif isinstance(d, list):
            for n, dd in enumerate(d):
                digest_close(T[k][n], dd, 2e-3, 1e-5, "%s[%d]" % (k, n))
        else:
            digest_close(T[k], d, 2e-3, 1e-5, k)
    if fx["training"]:
        out["all"].backward()
        pk = param_keys(state)
        grads = {k: state[k].grad for k in pk}
        gn = float(clip_grad_norm(list(grads.values()), 1.0))
        assert abs(gn - fx["grad_norm"]) <= 1e-3 * fx["grad_norm"]
        for k, d in fx["grads"].items():
            assert (d is None) == (grads[k] is None), k
            if d is not None:
                digest_close(grads[k], d, 5e-3, 1e-7, "grad:" + k)
        for k, d in fx["buffers"].items():
            digest_close(state[k], d, 1e-3, 1e-6, "buf:" + k)


@pytest.mark.parametrize("name", ["step_m4_b2"])
def test_oracle_step_faithful(name):
    _run_case(name, batched=False)


@pytest.mark.parametrize("name", ["step_m2_b2", "infer_m4_b2"])
def test_oracle_step_batched_identity(name):
    # Q2: the per-sample CondConv loop equals one batched conv (exact algebraic identity)
    _run_case(name, batched=True)


def test_oracle_loss_cases():
    fx = load_golden("loss_cases.pt")
    g = torch.Generator().manual_seed(fx["seed"])
    B, M, C, H, W = fx["B"], fx["M"], fx["C"], fx["H"], fx["W"]
    gt = [torch.randn(B, C, H, W, generator=g) for _ in range(M)]
    xs = [torch.randn(B, C, H, W, generator=g) for _ in range(M)]
    xm = [torch.randn(B, C, H, W, generator=g) for _ in range(M * (M - 1))]
    zs = [torch.randn(B, 16, generator=g) for _ in range(M)]
    zn = [torch.randn(B, 16, generator=g) for _ in range(M)]
    ss = [torch.softmax(torch.randn(B, 4, 160, 192, generator=g), 1) for _ in range(M)]
    tgt = torch.randint(0, 4, (B, 1, H, W), generator=g).float()
    ys = [torch.randn(B, 4, H, W, generator=g) for _ in range(M)]
    y1 = [torch.randn(B, 1, H, W, generator=g) for _ in range(M)]
    from oracle.rd_oracle import DEFAULT_CFG
    orc = RDOracle({}, DEFAULT_CFG)
    for row in fx["rows"]:
        mask = torch.tensor(row["mask"], dtype=torch.float32)
        vals = {
            "recon_x_p1": orc.recon_loss_x_list(gt, xs, mask, 1), "recon_x_p2": orc.recon_loss_x_list(gt, xs, mask, 2),
            "recon_x_mix_p1": orc.recon_loss_x_mix_list(gt, xm, mask, 1),
            "recon_x_mix_p2": orc.recon_loss_x_mix_list(gt, xm, mask, 2),
            "latent_z": orc.latent_z_loss(zs, zn, mask), "sim_s": orc.similarity_s_loss(ss, mask, tuple(row["pair"])),
            "sim_z": orc.similarity_z_loss(zs, mask), "recon_y_list_p1": orc.recon_loss_y_list(tgt, y1, mask, 1),
            "seg_y_list": orc.segmentation_loss_y_list(tgt, ys, mask),
        }
        if mask.sum() > 0:
            vals["kl"] = orc.kl_loss_list_standard(zs, zn, mask)
        for k, v in row["values"].items():
            assert abs(float(vals[k]) - v) <= 1e-5 * max(1.0, abs(v)), (row["mask"], k, float(vals[k]), v)
