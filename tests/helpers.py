"""Shared test helpers: rebuild the golden cases (weights, batch, eps) and compare digests."""
import json
import os

import torch

from tests.conftest import GOLDEN, load_golden
from oracle.params import synth_fill_
from oracle.rd_oracle import RDOracle, clone_state
import rd_b200.data as rd_data


def template_state(M: int = 4):
    """Zero tensors with the reference's state_dict keys / shapes (tests/golden/state_dict_keys*.json)."""
    fn = "state_dict_keys.json" if M == 4 else "state_dict_keys_m2.json"
    with open(os.path.join(GOLDEN, fn)) as f:
        keys = json.load(f)["keys"]
    st = {}
    for e in keys:
        dt = torch.int64 if e["key"].endswith("num_batches_tracked") else torch.float32
        st[e["key"]] = torch.zeros(e["shape"], dtype=dt)
    return st


def golden_state(fx, model=None):
    """The weights the fixture was made with.  Key / shape templates come from the reference's recorded state_dict
    (tests/golden/state_dict_keys*.json, out_num_ch = 1); a fixture with another head width (stage 2: out_num_ch = 4)
    takes the shapes from the product model, whose keys are pinned to the reference's by
    test_state_dict_matches_reference."""
    c = fx["cfg"]
    nondefault = (c.get("out_num_ch", 1) != 1 or c.get("fuse_method", "mean") == "mean-max-min" or c.get("shared_inp_dec", False)
                  or c.get("others", {}).get("mod_enc_s", False))
    if model is not None and nondefault:
        tmpl = {k: torch.zeros_like(v, device="cpu") for k, v in model.state_dict().items()}
        return synth_fill_(tmpl, seed=fx["param_seed"])
    return synth_fill_(template_state(fx["M"]), seed=fx["param_seed"])


def golden_inputs(fx):
    cfg = fx["cfg"]
    batch = rd_data.synthetic_batch(fx["B"], fx["M"], cfg["block_size"], cfg["input_height"], cfg["input_width"],
                                    seed=fx["seed"], missing=fx["mask_rows"], zero_border=fx["zero_border"])
    eps = rd_data.synthetic_eps(fx["B"], fx["M"], cfg["z_size"], seed=fx["seed"] + 1)
    return batch, eps


def digest_close(t, d, rtol, atol, what=""):
    """Compare tensor `t` with a stored digest (shape, float64 sum/abssum, strided sample)."""
    assert list(t.shape) == d["shape"], (what, list(t.shape), d["shape"])
    x = t.detach().to("cpu", torch.float64).reshape(-1)
    n = d["sample"].numel()
    step = max(1, x.numel() // 192) if n > 64 or x.numel() <= 64 * 3 else max(1, x.numel() // n)
    # the sampling rule must mirror oracle/make_golden.digest: step = numel // n_requested
    for n_req in (192, 64, 32):
        st = max(1, x.numel() // n_req)
        s = x[::st][:n_req]
        if s.numel() == n:
            step = st
            break
    s = x[::step][:n].to(torch.float32)
    ref = d["sample"]
    scale = max(float(ref.abs().max()), 1e-30)
    err = float((s - ref).abs().max())
    assert err <= atol + rtol * scale, "%s: sample max err %.3e (scale %.3e)" % (what, err, scale)
    assert abs(float(x.abs().sum()) - d["abssum"]) <= rtol * d["abssum"] + atol * x.numel(), \
        "%s: abssum %.6e vs %.6e" % (what, float(x.abs().sum()), d["abssum"])


def datafeed_inputs():
    """The seeded volume dict / index lists / metric inputs that oracle/make_golden.py::datafeed_case fed to the real reference
    (tests/golden/datafeed.pt holds the reference's outputs)."""
    import numpy as np
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
    subj_list = ["a", "b", "c", "a", "b", "c"]
    idx_list = [0, 77, 100, 151, 3, 148]
    gs = np.random.RandomState(11)
    tgt = gs.randint(0, 4, (5, 1, 40, 48)).astype(np.float32)
    pred = gs.randn(5, 4, 40, 48).astype(np.float32)
    pred[0, :3] = -1.0
    return contrasts, data, subj, subj_list, idx_list, tgt, pred
