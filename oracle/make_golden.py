"""TEST INFRASTRUCTURE — pin oracle/rd_oracle.py against the UNMODIFIED reference and write the
committed fixtures in tests/golden/.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py            # validates + (re)writes tests/golden/*.pt, *.json
    python oracle/make_golden.py --check    # validates only

What is compared (reference vs restatement, same weights / inputs / eps / (i,j)):
  * every loss term of the loop body src/main_missing.py:165-251, the clipped-grad norm (:272),
  * every parameter gradient (incl. which parameters get grad None),
  * s_i, z_mean, z_log_var, z, x_fake (M), x_fake_mix (M(M-1)), y_fake_list, y_fake_fused,
  * BN running statistics after the step,
  * loss functions alone over many masks (index / skip logic Q3, Q4, empty-mask paths).
The fixtures store losses, per-tensor digests (shape, float64 sum and abs-sum, a strided sample),
never full tensors, so they stay small.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader                      # noqa: E402
from oracle.params import synth_fill_              # noqa: E402
from oracle.rd_oracle import (DEFAULT_CFG, RDOracle, clone_state, param_keys, train_iteration)  # noqa: E402
import rd_b200.data as rd_data                     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(t: torch.Tensor, n: int = 192) -> dict:
    t = t.detach().to(torch.float64).reshape(-1)
    step = max(1, t.numel() // n)
    return {"shape": None, "numel": t.numel(), "sum": float(t.sum()), "abssum": float(t.abs().sum()),
            "sample": t[::step][:n].to(torch.float32).clone()}


def digest_of(t: torch.Tensor, n: int = 192) -> dict:
    d = digest(t, n)
    d["shape"] = list(t.shape)
    return d


def maxdiff(a: torch.Tensor, b: torch.Tensor):
    a, b = a.detach().double(), b.detach().double()
    ad = (a - b).abs().max().item() if a.numel() else 0.0
    rel = ad / max(b.abs().max().item(), 1e-30) if b.numel() else 0.0
    return ad, rel


def cfg_for(M: int, **kw) -> dict:
    cfg = json.loads(json.dumps(DEFAULT_CFG))
    if M == 2:
        cfg["contrast_list"] = ["T1", "T2"]
        cfg["dataset_name"] = "NCANDA"
    cfg.update(kw)
    return cfg


def reference_iteration(model, cfg, batch, eps_list, pair, with_y):
    """Loop body src/main_missing.py:165-272 around the real reference model."""
    ref = ref_loader.load_reference_model_module()
    M = len(cfg["contrast_list"])
    C = 2 * cfg["block_size"] + 1
    inputs, targets, mask, mask_img = batch["inputs"], batch["targets"], batch["mask"], batch["mask_img"]
    xs = [inputs[:, m * C:(m + 1) * C] for m in range(M)]
    calls = {"n": 0}

    def sample(self, z_mean, z_log_var):  # inject eps (Q8); later (cycle) calls use zeros
        k = calls["n"]
        calls["n"] += 1
        eps = eps_list[k] if k < M else torch.zeros_like(z_mean)
        return z_mean + eps * torch.exp(0.5 * z_log_var)

    orig_sample = type(model).sample
    orig_choice = np.random.choice
    type(model).sample = sample
    np.random.choice = lambda n, k, replace=False: np.array(pair)
    try:
        model.zero_grad(set_to_none=True)
        si = model.compute_anatomy_encoding(xs, mask_img)
        zi, mu, lv = model.compute_modality_encoding(xs, si, phase="train" if model.training else "test")
        x_self = model.reconstruct_input_si_zi(si, zi)
        x_mix = model.reconstruct_input_si_zj(si, zi)
        y_list = y_fused = None
        if with_y or cfg["lambda_recon_y"] > 0:
            y_list = model.reconstruct_output_si(si)
        if with_y or cfg["lambda_recon_y_fused"] > 0:
            y_fused = model.reconstruct_output_si_fused(si, mask)
        L = {}
        loss = 0
        brats = cfg["dataset_name"] == "BraTS"
        if cfg["lambda_recon_y"] > 0:
            L["recon_y"] = (model.compute_segmentation_loss_y_list(targets, y_list, mask) if brats
                            else model.compute_recon_loss_y_list(targets, y_list, mask, p=cfg["p"]))
            loss = loss + cfg["lambda_recon_y"] * L["recon_y"]
        if cfg["lambda_recon_y_fused"] > 0:
            L["recon_y_fused"] = (model.compute_segmentation_loss_y(targets, y_fused) if brats
                                  else model.compute_recon_loss_y(targets, y_fused, p=cfg["p"]))
            loss = loss + cfg["lambda_recon_y_fused"] * L["recon_y_fused"]
        if cfg["lambda_recon_x"] > 0:
            L["recon_x"] = model.compute_recon_loss_x_list(xs, x_self, mask, p=cfg["p"])
            loss = loss + cfg["lambda_recon_x"] * L["recon_x"]
        if cfg["lambda_recon_x_mix"] > 0:
            L["recon_x_mix"] = model.compute_recon_loss_x_mix_list(xs, x_mix, mask, p=cfg["p"])
            loss = loss + cfg["lambda_recon_x_mix"] * L["recon_x_mix"]
        if cfg["lambda_kl"] > 0:
            L["kl"] = model.compute_kl_loss_list_standard(mu, lv, mask)
            loss = loss + cfg["lambda_kl"] * L["kl"]
        mu_new = None
        if cfg["lambda_latent_z"] > 0:
            si_new = model.compute_anatomy_encoding(x_self, mask_img)
            _, mu_new, _ = model.compute_modality_encoding(x_self, si_new, phase="train" if model.training else "test")
            L["latent_z"] = model.compute_latent_z_loss(mu, mu_new, mask)
            loss = loss + cfg["lambda_latent_z"] * L["latent_z"]
        if cfg["lambda_sim_s"] > 0:
            L["sim_s"] = model.compute_similarity_s_loss(si, mask)
            loss = loss + cfg["lambda_sim_s"] * L["sim_s"]
        if cfg["lambda_sim_z"] > 0:
            L["sim_z"] = model.compute_similarity_z_loss(zi, mask)
            loss = loss + cfg["lambda_sim_z"] * L["sim_z"]
        L["all"] = loss
        gn = None
        grads = {}
        if model.training:
            loss.backward()
            gn = float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0))
            grads = {k: p.grad for k, p in model.named_parameters()}
    finally:
        type(model).sample = orig_sample
        np.random.choice = orig_choice
    losses = {k: float(v) for k, v in L.items()}
    tensors = {"si": si, "zi": zi, "z_mean": mu, "z_log_var": lv, "x_fake": x_self, "x_fake_mix": x_mix,
               "y_fake_list": y_list, "y_fake_fused": y_fused, "z_mean_new": mu_new}
    return losses, grads, gn, tensors


def compare_tensors(name, ref_t, ora_t, tol, report):
    if ref_t is None:
        assert ora_t is None, name
        return
    if isinstance(ref_t, (list, tuple)):
        assert len(ref_t) == len(ora_t), name
        for k, (a, b) in enumerate(zip(ref_t, ora_t)):
            compare_tensors("%s[%d]" % (name, k), a, b, tol, report)
        return
    assert tuple(ref_t.shape) == tuple(ora_t.shape), (name, ref_t.shape, ora_t.shape)
    ad, rel = maxdiff(ora_t, ref_t)
    report.append((name, ad, rel))
    assert rel <= tol or ad <= 1e-6, "%s: abs %.3e rel %.3e" % (name, ad, rel)


def step_case(name, M, B, mask_rows, pair, with_y, zero_border, seed, training=True, cfg_kw=None, write=True, full_tensors=False):
    cfg = cfg_for(M, **(cfg_kw or {}))
    torch.manual_seed(10)
    np.random.seed(10)
    model = ref_loader.build_reference_model(cfg, "cpu")
    synth_fill_(model.state_dict(), seed=1234)
    model.train(training)
    FROZEN = ("anatomy_encoder_enc_list.", "anatomy_encoder_dec.", "modality_encoder_list.", "input_decoder_list.")
    if cfg.get("fix_pretrain") and cfg.get("continue_train"):      # src/main_missing.py:104-116: stage-1 parts frozen
        for k, p in model.named_parameters():
            if k.startswith(FROZEN):
                p.requires_grad = False
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = rd_data.synthetic_batch(B, M, cfg["block_size"], cfg["input_height"], cfg["input_width"], seed=seed,
                                    missing=mask_rows, zero_border=zero_border)
    eps = rd_data.synthetic_eps(B, M, cfg["z_size"], seed=seed + 1)
    t0 = time.time()
    r_losses, r_grads, r_gn, r_t = reference_iteration(model, cfg, batch, eps, pair, with_y)
    t_ref = time.time() - t0
    ostate = clone_state(state0)
    if cfg.get("fix_pretrain") and cfg.get("continue_train"):
        for k in param_keys(ostate):
            if k.startswith(FROZEN):
                ostate[k].requires_grad_(False)
    orc = RDOracle(ostate, cfg, training=training)
    t0 = time.time()
    if training:
        o_losses, o_grads, o_gn, o_t = train_iteration_with_y(orc, batch, eps, pair, with_y)
    else:
        with torch.no_grad():
            out = orc.forward_losses(batch["inputs"], batch["targets"], batch["mask"], batch["mask_img"], eps, pair,
                                     with_y=with_y, keep=True)
        o_losses = {k: float(v) for k, v in out.items() if k != "tensors"}
        o_grads, o_gn, o_t = {}, None, out["tensors"]
    t_orc = time.time() - t0
    report = []
    for k, v in r_losses.items():
        assert abs(o_losses[k] - v) <= 2e-5 * max(1.0, abs(v)), (name, k, o_losses[k], v)
    if training:
        assert abs(o_gn - r_gn) <= 1e-4 * max(1.0, r_gn), (o_gn, r_gn)
        for k, g in r_grads.items():
            og = o_grads[k]
            assert (g is None) == (og is None), "grad None mismatch for " + k
            if g is not None:
                compare_tensors("grad:" + k, g, og, 2e-3, report)
    for k in r_t:
        compare_tensors(k, r_t[k], o_t[k], 1e-3, report)
    new_state = model.state_dict()
    for k in new_state:
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            compare_tensors("buf:" + k, new_state[k].float(), ostate[k].float(), 1e-4, report)
    worst = sorted(report, key=lambda r: -r[2])[:5]
    print("[%s] ref %.1fs oracle %.1fs  losses %s  gn %s" % (name, t_ref, t_orc, {k: round(v, 6) for k, v in r_losses.items()}, r_gn))
    print("   worst rel diffs:", [(n, "%.2e" % a, "%.2e" % r) for n, a, r in worst])
    if not write:
        return
    fx = {"name": name, "cfg": cfg, "B": B, "M": M, "mask_rows": mask_rows, "pair": list(pair), "with_y": with_y,
          "zero_border": zero_border, "seed": seed, "training": training, "param_seed": 1234,
          "losses": r_losses, "grad_norm": r_gn,
          "grads": {k: (None if g is None else digest_of(g, 64)) for k, g in r_grads.items()},
          "tensors": {}, "buffers": {}}
    for k, v in r_t.items():
        if v is None:
            fx["tensors"][k] = None
        elif isinstance(v, (list, tuple)):
            fx["tensors"][k] = [digest_of(t) for t in v]
        else:
            fx["tensors"][k] = digest_of(v)
    for k in new_state:
        if k.endswith("running_mean") or k.endswith("running_var"):
            fx["buffers"][k] = digest_of(new_state[k].float(), 32)
    torch.save(fx, os.path.join(GOLD, name + ".pt"))
    if full_tensors:
        # one fixture also keeps FULL reference tensors (not digests) of the first contrast / first slice: the anatomy code s_0 and the
        # self-reconstruction x-hat_0, every pixel — the GPU tests compare them element by element (~1.3 MB)
        torch.save({"si0": r_t["si"][0][0:1].detach().float().clone(), "x_fake0": r_t["x_fake"][0][0:1].detach().float().clone(),
                    "x_fake_mix0": r_t["x_fake_mix"][0][0:1].detach().float().clone()}, os.path.join(GOLD, name + "_tensors.pt"))


def train_iteration_with_y(orc, batch, eps, pair, with_y):
    pk = param_keys(orc.P)
    out = orc.forward_losses(batch["inputs"], batch["targets"], batch["mask"], batch["mask_img"], eps, pair,
                             with_y=with_y, keep=True)
    out["all"].backward()
    from oracle.rd_oracle import clip_grad_norm
    grads = {k: orc.P[k].grad for k in pk}
    gn = float(clip_grad_norm(list(grads.values()), 1.0))
    return {k: float(v) for k, v in out.items() if k != "tensors"}, grads, gn, out["tensors"]


def loss_cases(write=True):
    """Loss functions alone over many masks: the integer skip/index logic (Q3, Q4, Q10)."""
    cfg = cfg_for(4)
    torch.manual_seed(10)
    model = ref_loader.build_reference_model(cfg_for(4, target_model_name="U+SA"), "cpu")
    orc = RDOracle({}, cfg)
    g = torch.Generator().manual_seed(77)
    B, M, C, H, W = 3, 4, 7, 32, 48
    gt = [torch.randn(B, C, H, W, generator=g) for _ in range(M)]
    xs = [torch.randn(B, C, H, W, generator=g) for _ in range(M)]
    xm = [torch.randn(B, C, H, W, generator=g) for _ in range(M * (M - 1))]
    zs = [torch.randn(B, 16, generator=g) for _ in range(M)]
    zn = [torch.randn(B, 16, generator=g) for _ in range(M)]
    ss = [torch.softmax(torch.randn(B, 4, 160, 192, generator=g), 1) for _ in range(M)]
    tgt = torch.randint(0, 4, (B, 1, H, W), generator=g).float()
    ys = [torch.randn(B, 4, H, W, generator=g) for _ in range(M)]
    y1 = [torch.randn(B, 1, H, W, generator=g) for _ in range(M)]
    rows = []
    masks = []
    all_rows = [[(v >> k) & 1 for k in range(M)] for v in range(16)]
    gm = torch.Generator().manual_seed(5)
    for trial in range(40):
        idx = torch.randint(0, 16, (B,), generator=gm).tolist()
        masks.append([all_rows[i] for i in idx])
    masks += [[[0] * 4] * 3, [[1] * 4] * 3, [[1, 0, 0, 0]] * 3, [[0, 1, 1, 0], [1, 0, 0, 1], [1, 1, 0, 0]]]
    pairs = [(0, 1), (2, 0), (3, 1), (1, 2)]
    for n, mrows in enumerate(masks):
        mask = torch.tensor(mrows, dtype=torch.float32)
        pair = pairs[n % len(pairs)]
        orig_choice = np.random.choice
        np.random.choice = lambda a, k, replace=False: np.array(pair)
        try:
            sim_s = model.compute_similarity_s_loss(ss, mask)
        finally:
            np.random.choice = orig_choice
        ref_vals = {
            "recon_x_p1": float(model.compute_recon_loss_x_list(gt, xs, mask, p=1)),
            "recon_x_p2": float(model.compute_recon_loss_x_list(gt, xs, mask, p=2)),
            "recon_x_mix_p1": float(model.compute_recon_loss_x_mix_list(gt, xm, mask, p=1)),
            "recon_x_mix_p2": float(model.compute_recon_loss_x_mix_list(gt, xm, mask, p=2)),
            "latent_z": float(model.compute_latent_z_loss(zs, zn, mask)),
            "sim_s": float(sim_s),
            "sim_z": float(model.compute_similarity_z_loss(zs, mask)),
            "recon_y_list_p1": float(model.compute_recon_loss_y_list(tgt, y1, mask, p=1)),
            "seg_y_list": float(model.compute_segmentation_loss_y_list(tgt, ys, mask)),
        }
        if mask.sum() > 0:
            ref_vals["kl"] = float(model.compute_kl_loss_list_standard(zs, zn, mask))
        ora_vals = {
            "recon_x_p1": float(orc.recon_loss_x_list(gt, xs, mask, 1)),
            "recon_x_p2": float(orc.recon_loss_x_list(gt, xs, mask, 2)),
            "recon_x_mix_p1": float(orc.recon_loss_x_mix_list(gt, xm, mask, 1)),
            "recon_x_mix_p2": float(orc.recon_loss_x_mix_list(gt, xm, mask, 2)),
            "latent_z": float(orc.latent_z_loss(zs, zn, mask)),
            "sim_s": float(orc.similarity_s_loss(ss, mask, pair)),
            "sim_z": float(orc.similarity_z_loss(zs, mask)),
            "recon_y_list_p1": float(orc.recon_loss_y_list(tgt, y1, mask, 1)),
            "seg_y_list": float(orc.segmentation_loss_y_list(tgt, ys, mask)),
        }
        if mask.sum() > 0:
            ora_vals["kl"] = float(orc.kl_loss_list_standard(zs, zn, mask))
        for k, v in ref_vals.items():
            assert abs(ora_vals[k] - v) <= 1e-6 * max(1.0, abs(v)), (n, k, ora_vals[k], v)
        rows.append({"mask": mrows, "pair": list(pair), "values": ref_vals})
    # fusion gather order (Q3): which (b, m) rows are selected, and their order
    fus = []
    for mrows in masks[:12]:
        mask = torch.tensor(mrows, dtype=torch.float32)
        tag = torch.stack([torch.stack([torch.full((1, 1, 1), float(b * 10 + m)) for m in range(M)], 0) for b in range(B)], 0)
        sel = tag[mask == 1].flatten().tolist()
        fus.append({"mask": mrows, "order": sel})
    print("[loss_cases] %d masks x %d loss terms agree (<=1e-6 rel)" % (len(rows), len(rows[0]["values"])))
    if write:
        torch.save({"seed": 77, "B": B, "M": M, "C": C, "H": H, "W": W, "rows": rows, "fusion_order": fus},
                   os.path.join(GOLD, "loss_cases.pt"))


def datafeed_case(write=True):
    """f-1 / f-3: the REAL reference's ZeroDoseDataset.__getitem__ (src/util.py:471-566) and compute_segmentation_metrics
    (src/util.py:946-992) on seeded inputs, against their restatements in oracle/metrics_oracle.py (exact), stored as a small fixture.
    (compute_reconstruction_metrics needs scikit-image, absent here: unpinned, see oracle/metrics_oracle.py.)"""
    from oracle.metrics_oracle import assemble_sample, compute_segmentation_metrics
    ref = ref_loader.load_reference_model_module()          # `from util import *` puts the util names into the model module
    contrasts = ["T1", "T1c", "T2", "T2_FLAIR"]
    g = np.random.RandomState(3)
    data, subj = {}, ["a", "b", "c"]
    H, W, D = 160, 192, 155
    for s in subj:
        for c in contrasts:
            if not (s == "b" and c == "T2"):
                v = g.randn(H, W, 12).astype(np.float32)
                v[:2] = 0
                data[s + "/" + c] = np.tile(v, (1, 1, 13))[:, :, :D]
        if s != "c":
            data[s + "/seg"] = np.tile(g.randint(0, 5, (H, W, 12)).astype(np.float32), (1, 1, 13))[:, :, :D]
    subj_list = np.array(["a", "b", "c", "a", "b", "c"])
    idx_list = np.array([0, 77, 100, 151, 3, 148])
    with contextlib.redirect_stdout(io.StringIO()):
        ds = ref.ZeroDoseDataset("BraTS", data, subj_list, idx_list, None, block_size=3, contrast_list=contrasts, dropoff=True)
    rows = []
    np.random.seed(5)
    ref_items = [ds[k] for k in range(len(subj_list))]
    np.random.seed(5)
    for k, it in enumerate(ref_items):
        assert it is not None, k
        present = np.array([1 if subj_list[k] + "/" + c in data else 0 for c in contrasts])
        drop = None
        if present.sum() > 1 and np.random.rand() > 0.8:
            drop = int(np.random.choice(np.where(present == 1)[0], 1)[0])
        mine = assemble_sample(data, str(subj_list[k]), int(idx_list[k]), contrasts, 3, "BraTS", drop_idx=drop)
        for key in ("inputs", "targets", "mask", "mask_img"):
            assert np.array_equal(np.asarray(it[key], dtype=np.float64), np.asarray(mine[key], dtype=np.float64)), (k, key)
        assert int(it["slice_idx"]) == mine["slice_idx"]
        rows.append({"subj": str(subj_list[k]), "idx": int(idx_list[k]), "drop": -1 if drop is None else drop, "slice_idx": int(it["slice_idx"]),
                     "mask": [int(v) for v in it["mask"]], "inputs_sum": float(np.asarray(it["inputs"], dtype=np.float64).sum()),
                     "inputs_abssum": float(np.abs(np.asarray(it["inputs"], dtype=np.float64)).sum()),
                     "targets_sum": float(np.asarray(it["targets"], dtype=np.float64).sum()),
                     "mask_img_sum": float(np.asarray(it["mask_img"]).sum())})
    gs = np.random.RandomState(11)
    tgt = gs.randint(0, 4, (5, 1, 40, 48)).astype(np.float32)
    pred = gs.randn(5, 4, 40, 48).astype(np.float32)
    pred[0, :3] = -1.0                                   # an image without any positive prediction: the +1 smoothing decides
    r = ref.compute_segmentation_metrics(tgt, pred)
    m = compute_segmentation_metrics(tgt, pred)
    assert np.array_equal(np.array(r["dice"]), np.array(m["dice"])) and np.array_equal(np.array(r["iou"]), np.array(m["iou"]))
    print("[datafeed] %d dataset items and 5 segmentation-metric rows agree exactly with the reference" % len(rows))
    if write:
        torch.save({"rows": rows, "seg": {"dice": [float(v) for v in r["dice"]], "iou": [float(v) for v in r["iou"]]}},
                   os.path.join(GOLD, "datafeed.pt"))


def state_keys(write=True):
    cfg = cfg_for(4)
    torch.manual_seed(10)
    model = ref_loader.build_reference_model(cfg, "cpu")
    sd = model.state_dict()
    params = dict(model.named_parameters())
    keys = [{"key": k, "shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", ""),
             "param": k in params} for k, v in sd.items()]
    stats = {}
    for k, p in params.items():  # init statistics, so the product's init can be checked distributionally
        stats[k] = {"mean": float(p.detach().double().mean()), "std": float(p.detach().double().std()) if p.numel() > 1 else 0.0}
    print("[state_keys] %d keys (%d params), %d elements" % (len(keys), len(params), sum(p.numel() for p in params.values())))
    if write:
        with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
            json.dump({"keys": keys, "init_stats": stats}, f)
    cfg2 = cfg_for(2)
    model2 = ref_loader.build_reference_model(cfg2, "cpu")
    keys2 = [{"key": k, "shape": list(v.shape)} for k, v in model2.state_dict().items()]
    if write:
        with open(os.path.join(GOLD, "state_dict_keys_m2.json"), "w") as f:
            json.dump({"keys": keys2}, f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    write = not a.check
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    todo = a.only.split(",") if a.only else ["keys", "loss", "step_m4", "step_m4_full", "infer_m4", "step_m2", "stage2", "variants", "shared", "stage2_u", "stage2_saca", "stage2_ssaca",
                                                "skip", "kl_p2", "fused_zd", "fused_brats", "datafeed"]
    if "keys" in todo:
        state_keys(write)
    if "datafeed" in todo:
        datafeed_case(write)
    if "loss" in todo:
        loss_cases(write)
    if "step_m4" in todo:   # config 1: B=2, M=4, default lambdas, one missing contrast, y at "iter 0"
        step_case("step_m4_b2", 4, 2, [[1, 1, 1, 1], [1, 0, 1, 1]], (0, 2), True, 8, seed=10, write=write)
    if "step_m4_full" in todo:   # all present, no y, other pair
        step_case("step_m4_b2_full", 4, 2, [[1, 1, 1, 1], [1, 1, 1, 1]], (3, 1), False, 0, seed=21, write=write, full_tensors=True)
    if "infer_m4" in todo:  # eval mode / phase test: inference sweep building block
        step_case("infer_m4_b2", 4, 2, [[1, 0, 1, 0], [0, 1, 1, 1]], (0, 2), True, 8, seed=12, training=False, write=write)
    if "step_m2" in todo:   # config 4: NCANDA 2-contrast
        step_case("step_m2_b2", 2, 2, [[1, 1], [1, 1]], (0, 1), False, 0, seed=14, write=write)
    if "stage2" in todo:    # f-2: output decoder under grad, BraTS seg loss (out_num_ch 4).  The reference's
        # fused variant (lambda_recon_y_fused > 0) raises for any mask with K != B rows (K = mask.sum()
        # rows out of reconstruct_output_si_fused vs B target rows), so only the list variant is pinned.
        step_case("stage2_m4_b2", 4, 2, [[1, 1, 0, 1], [1, 1, 1, 1]], (1, 3), False, 8, seed=16,
                  cfg_kw={"lambda_recon_y": 1.0, "out_num_ch": 4}, write=write)
    if "variants" in todo:  # f-4 / rows a9, a11, a15: mean-normalised data (softplus decoder / target / anatomy activations),
        # fuse_method mean-max-min (12-channel fused code), s_compact_method mean (average pool); y at "iter 0"
        others = dict(DEFAULT_CFG["others"])
        others["ana_dec_act"] = "softplus"
        step_case("variants_m4_b2", 4, 2, [[1, 1, 1, 0], [0, 1, 1, 1]], (2, 1), True, 8, seed=18,
                  cfg_kw={"input_output_act": "softplus", "target_output_act": "softplus", "fuse_method": "mean-max-min",
                          "s_compact_method": "mean", "others": others}, write=write)
    if "shared" in todo:    # f-4: shared_inp_dec (one SPADENew decoder for all contrasts) + mod_enc_s (the modality encoder
        # sees the anatomy code, so the cycle's second anatomy encoding has a gradient path, cf. Q7)
        others = dict(DEFAULT_CFG["others"])
        others["mod_enc_s"] = True
        step_case("shared_m4_b2", 4, 2, [[1, 1, 1, 1], [1, 1, 0, 1]], (3, 0), False, 0, seed=19,
                  cfg_kw={"shared_inp_dec": True, "others": others}, write=write)
    if "stage2_u" in todo:  # f-4: target_model_name 'U' (GANShortGenerator, the output U-Net without attention gates) under grad
        step_case("stage2_u_m4_b2", 4, 2, [[1, 1, 1, 1], [0, 1, 1, 1]], (2, 0), False, 8, seed=23,
                  cfg_kw={"lambda_recon_y": 1.0, "out_num_ch": 4, "target_model_name": "U"}, write=write)
    if "stage2_saca" in todo:   # f-4: target_model_name 'U+SA+CA' (channel attention + spatial attention on every skip connection)
        step_case("stage2_saca_m4_b2", 4, 2, [[1, 1, 0, 1], [1, 1, 1, 1]], (0, 3), False, 8, seed=33,
                  cfg_kw={"lambda_recon_y": 1.0, "out_num_ch": 4, "target_model_name": "U+SA+CA"}, write=write)
    if "stage2_ssaca" in todo:  # f-4: target_model_name 'U+SSA+CA' (the gate sees g and |g - flip(g)|, residual attention)
        step_case("stage2_ssaca_m4_b2", 4, 2, [[1, 1, 1, 1], [1, 0, 1, 1]], (1, 2), False, 8, seed=35,
                  cfg_kw={"lambda_recon_y": 1.0, "out_num_ch": 4, "target_model_name": "U+SSA+CA"}, write=write)
    if "skip" in todo:      # Q4 / Q10 under grad at step level: contrast 3 is missing in EVERY row -> its self term and every pair with it
        # are skipped, the 6 counted pairs read x_mix slots 0..5 (index lag: decoders 0 and 1 only), so the private decoder half of
        # contrast 3 is not connected to the loss at all: its parameters get grad None and Adam skips them in this iteration
        step_case("step_m4_b2_skip", 4, 2, [[1, 1, 1, 0], [1, 0, 1, 0]], (0, 2), False, 8, seed=25, write=write)
    if "kl_p2" in todo:     # lambda_kl > 0 (KL to N(0, I) on the modality codes) and p = 2 (squared reconstruction error)
        step_case("step_m4_b2_kl_p2", 4, 2, [[1, 1, 1, 1], [0, 1, 1, 1]], (1, 2), False, 0, seed=27,
                  cfg_kw={"lambda_kl": 0.1, "p": 2}, write=write)
    if "fused_zd" in todo:  # the paper's stage 2 (commented block src/config.yaml:47-52 + fix_pretrain / continue_train, src/main_missing.py:104-116)
        # on a ZeroDose-type dataset at batch 1: y_fake_fused has K = 3 rows, compute_recon_loss_y broadcasts the single target over them
        step_case("stage2_fused_zd_b1", 4, 1, [[1, 0, 1, 1]], (0, 2), False, 8, seed=29,
                  cfg_kw={"dataset_name": "ZeroDose", "contrast_list": ["T1", "T1c", "T2_FLAIR", "ASL"], "lambda_recon_y": 1.0,
                          "lambda_recon_y_fused": 2.0, "lambda_recon_x": 0.0, "lambda_recon_x_mix": 0.0, "lambda_sim_s": 0.0,
                          "lambda_sim_z": 0.0, "fix_pretrain": True, "continue_train": True}, write=write)
    if "fused_brats" in todo:   # lambda_recon_y_fused > 0 with the segmentation head: the reference only works when K == B rows come out of the
        # fusion (cross_entropy needs equal batch sizes); one present contrast per sample gives K = B = 2
        step_case("stage2_fused_brats_b2", 4, 2, [[1, 0, 0, 0], [0, 0, 1, 0]], (0, 2), False, 8, seed=31,
                  cfg_kw={"lambda_recon_y": 1.0, "lambda_recon_y_fused": 2.0, "out_num_ch": 4}, write=write)
    print("OK")


if __name__ == "__main__":
    main()
