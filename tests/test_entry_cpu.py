"""CPU (kernels emulated in torch, tests/emul.py): the `main_missing.py`-compatible entry point — config handling, checkpoint
directory / yaml override, train() for one iteration + validation + checkpoint, resume by key (model, optimizer incl. per-parameter
steps, scheduler), evaluate() with the device metrics path, and the slab loader against the reference dataset semantics."""
import os

import numpy as np
import pytest
import torch
import yaml

from tests import emul
import rd_b200.config as rd_config
import rd_b200.data as rd_data


@pytest.fixture()
def emulated():
    saved = emul.install()
    yield
    emul.uninstall(saved)


def _cfg(tmp_path, **kw):
    cfg = rd_config.default_config(precision="fp32", batch_size=2, epochs=1, phase="train", ckpt_root=str(tmp_path), load_yaml=True,
                                   synthetic_subjects=3, synthetic_depth=8, max_iters=1, max_eval_iters=1, cuda_graph=False,
                                   ckpt_timelabel="")
    cfg.update(kw)
    return cfg


def test_train_then_resume_then_test_phase(emulated, tmp_path):
    from rd_b200.main_missing import Run
    run = Run(_cfg(tmp_path), device=torch.device("cpu"), log=lambda *a: None)
    stat = run.train()
    ck = run.config["ckpt_path"]
    assert os.path.isfile(os.path.join(ck, "epoch000.pth.tar")) and os.path.isfile(os.path.join(ck, "model_best.pth.tar"))
    assert os.path.isfile(os.path.join(ck, "config.yaml")) and os.path.isfile(os.path.join(ck, "config.txt"))
    with open(os.path.join(ck, "stat.csv")) as f:
        rows = f.read().strip().split("\n")
    assert rows[0].startswith(",info,") and len(rows) == 3 and "epoch[ 0]" in rows[1] and "val" in rows[2]
    for k in ("recon_x", "recon_x_mix", "latent_z", "sim_s", "sim_z", "all", "ssim", "psnr", "rmse"):
        assert np.isfinite(stat[k]), k
    saved = torch.load(os.path.join(ck, "model_best.pth.tar"), weights_only=False)
    assert sorted(saved.keys()) == ["epoch", "model", "monitor_metric", "optimizer", "scheduler", "stat"]
    assert saved["monitor_metric"] == stat["recon_x_mix"]
    # batch_size 2 -> accumulation window 8 -> one iteration takes no optimizer step yet: no Adam state, gradients stay accumulated
    assert saved["optimizer"]["state"] == {} and float(run.trainer.fp.grad.abs().sum()) > 0
    # resume: same directory through ckpt_timelabel, phase test -> evaluate(test) with saved results
    label = os.path.basename(ck)
    cfg2 = _cfg(tmp_path, phase="test", ckpt_timelabel=label, lr=123.0)       # lr comes back from the saved yaml (load_yaml)
    run2 = Run(cfg2, device=torch.device("cpu"), log=lambda *a: None)
    assert run2.config["lr"] == pytest.approx(2e-4) and run2.start_epoch == 0
    for (k, a), (_, b) in zip(run.model.state_dict().items(), run2.model.state_dict().items()):
        assert torch.equal(a, b), k
    st2 = run2.evaluate(phase="test", set="test", save_res=True)
    assert np.isfinite(st2["recon_x_mix"]) and np.isfinite(st2["ssim"])
    res = torch.load(os.path.join(run2.config["ckpt_path"], "result_test", "results_all.pt"), weights_only=False)
    assert tuple(res["xi_fake_mix"].shape[1:]) == (12, 7, 160, 192) and tuple(res["s_list"].shape[1:]) == (4, 4, 160, 192)


def test_slab_loader_matches_reference_dataset_semantics(emulated):
    """SlabLoader / rd_assemble_slabs (emulated here, the CUDA kernel in test_metrics_slabs_gpu.py) against the restated
    ZeroDoseDataset.__getitem__ (oracle/metrics_oracle.assemble_sample, src/util.py:471-566): window clamp, missing contrast,
    dropoff with the reference's NumPy RNG calls, BraTS label 4 -> 3, mask_img."""
    from oracle.metrics_oracle import assemble_sample
    contrasts = ["T1", "T1c", "T2", "T2_FLAIR"]
    g = np.random.RandomState(3)
    data, subj = {}, ["a", "b", "c"]
    H, W, D = 16, 24, 155
    for s in subj:
        for c in contrasts:
            if not (s == "b" and c == "T2"):
                v = g.randn(H, W, D).astype(np.float32)
                v[:2] = 0
                data[s + "/" + c] = v
        if s != "c":
            data[s + "/seg"] = g.randint(0, 5, (H, W, D)).astype(np.float32)
    store = rd_data.VolumeStore.from_dict(data, subj, contrasts, "BraTS", "cpu")
    subj_list = ["a", "b", "c", "a", "b", "zzz", "c"]
    idx_list = [0, 77, 154, 152, 3, 5, 100]
    loader = rd_data.SlabLoader(store, subj_list, idx_list, batch_size=3, shuffle=False, dropoff=True)
    # skipped like nonechucks skips a failing sample: the unknown subject, and the two windows that leave the 155-slice volume after the
    # reference's clamp to 155 - block (slice 152 + 3 = index 155 does not exist; the reference would return a 6-slice window there)
    assert len(loader.items) == 4
    valid = [(s, i) for s, i in zip(subj_list, idx_list) if s in subj and min(max(i, 3), 152) + 4 <= D]
    np.random.seed(5)
    batches = [{k: (v.clone() if torch.is_tensor(v) else list(v)) for k, v in b.items()} for b in loader]
    assert [b["inputs"].shape[0] for b in batches] == [3, 1]
    np.random.seed(5)
    k = 0
    for b in batches:
        for r in range(b["inputs"].shape[0]):
            s, i = valid[k]
            k += 1
            present = np.array([1 if s + "/" + c in data else 0 for c in contrasts])
            drop = None
            if present.sum() > 1 and np.random.rand() > 0.8:
                drop = int(np.random.choice(np.where(present == 1)[0], 1)[0])
            ref = assemble_sample(data, s, i, contrasts, 3, "BraTS", image_size=(H, W), drop_idx=drop)
            assert np.array_equal(b["inputs"][r].numpy(), ref["inputs"].astype(np.float32)), (s, i)
            assert np.array_equal(b["targets"][r].numpy(), ref["targets"].astype(np.float32)), (s, i)
            assert b["mask"][r].tolist() == ref["mask"].tolist() and int(b["slice_idx"][r]) == ref["slice_idx"]
            assert np.array_equal(b["mask_img"][r].numpy(), ref["mask_img"].astype(np.float32))
