"""GPU: the kernels either side of the hot path through the C ABI.
  * rd_assemble_slabs (device-resident volume store -> batch) against the outputs of the REAL reference's
    ZeroDoseDataset.__getitem__ (tests/golden/datafeed.pt) and against the oracle restatement: bit-exact;
  * rd_metrics_seg against the real reference's compute_segmentation_metrics values (fixture) — exact to fp32 rounding;
  * rd_metrics_recon (PSNR / SSIM / MSE) against the NumPy restatement of the skimage algorithms (oracle/metrics_oracle.py;
    scikit-image itself is absent: unpinned) — 1e-4."""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden
from tests.helpers import datafeed_inputs
import rd_b200.data as rd_data
import rd_b200.kernels as K

pytestmark = pytest.mark.gpu


def test_slab_assembly_matches_reference_dataset():
    from oracle.metrics_oracle import assemble_sample
    fx = load_golden("datafeed.pt")
    contrasts, data, subj, subj_list, idx_list, _, _ = datafeed_inputs()
    store = rd_data.VolumeStore.from_dict(data, subj, contrasts, "BraTS", "cuda:0")
    loader = rd_data.SlabLoader(store, subj_list, idx_list, batch_size=4, shuffle=False, dropoff=True)
    assert len(loader.items) == 6
    rows = [(store.index[s], i) for s, i in zip(subj_list, idx_list)]
    drops = [r["drop"] for r in fx["rows"]]
    got = []
    for k in range(0, 6, 4):
        b = loader.assemble(rows[k:k + 4], drops[k:k + 4])
        got.append({kk: (v.cpu().clone() if torch.is_tensor(v) else v) for kk, v in b.items()})
    k = 0
    for b in got:
        for r in range(b["inputs"].shape[0]):
            ref = fx["rows"][k]
            x = b["inputs"][r].double()
            assert float(x.sum()) == pytest.approx(ref["inputs_sum"], rel=1e-12, abs=1e-9)
            assert float(x.abs().sum()) == pytest.approx(ref["inputs_abssum"], rel=1e-12)
            assert float(b["targets"][r].double().sum()) == ref["targets_sum"]
            assert [int(v) for v in b["mask"][r]] == ref["mask"] and int(b["slice_idx"][r]) == ref["slice_idx"]
            assert float(b["mask_img"][r].sum()) == ref["mask_img_sum"]
            mine = assemble_sample(data, subj_list[k], idx_list[k], contrasts, 3, "BraTS", drop_idx=None if drops[k] < 0 else drops[k])
            assert np.array_equal(b["inputs"][r].numpy(), mine["inputs"].astype(np.float32))
            assert np.array_equal(b["targets"][r].numpy(), mine["targets"].astype(np.float32))
            assert np.array_equal(b["mask_img"][r].numpy(), mine["mask_img"].astype(np.float32))
            k += 1
    # the loader's own dropoff decisions use the reference's NumPy calls: same seed -> same drops as the reference run
    np.random.seed(5)
    assert loader._drop_decisions(rows) == drops


def test_skull_strip_and_full_epoch_shapes():
    contrasts, data, subj, subj_list, idx_list, _, _ = datafeed_inputs()
    bm = (np.random.RandomState(1).rand(160, 192, 155) > 0.3).astype(np.float32)
    store = rd_data.VolumeStore.from_dict(data, subj, contrasts, "BraTS", "cuda:0", brain_mask=bm)
    from oracle.metrics_oracle import assemble_sample
    loader = rd_data.SlabLoader(store, subj_list, idx_list, batch_size=4, shuffle=True, dropoff=False)
    torch.manual_seed(0)
    seen = 0
    for b in loader:
        for r in range(b["inputs"].shape[0]):
            s, i = b["subj_id"][r], int(b["slice_idx"][r])
            mine = assemble_sample(data, s, i, contrasts, 3, "BraTS", brain_mask=bm)
            assert np.array_equal(b["inputs"][r].cpu().numpy(), mine["inputs"].astype(np.float32))
            assert np.array_equal(b["targets"][r].cpu().numpy(), mine["targets"].astype(np.float32))
            seen += 1
    assert seen == 6


def test_segmentation_metrics_match_reference_values():
    fx = load_golden("datafeed.pt")
    *_, tgt, pred = datafeed_inputs()
    for dt in (torch.float32,):
        p = torch.from_numpy(pred).permute(0, 2, 3, 1).contiguous().cuda().to(dt)
        t = torch.from_numpy(tgt).reshape(5, -1).contiguous().cuda()
        out = torch.empty(5, 2, device="cuda")
        K.metrics_seg(t, p, out)
        o = out.cpu().double()
        assert torch.allclose(o[:, 0], torch.tensor(fx["seg"]["dice"]).double(), rtol=1e-6, atol=0)
        assert torch.allclose(o[:, 1], torch.tensor(fx["seg"]["iou"]).double(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("pdt", [torch.float32, torch.bfloat16])
def test_reconstruction_metrics_match_skimage_restatement(pdt):
    from oracle.metrics_oracle import compute_reconstruction_metrics_single
    g = torch.Generator().manual_seed(4)
    N, H, W, C = 6, 160, 192, 7
    base = torch.randn(N, H, W, C, generator=g)
    base[:, :20] = -10.0                                      # the -10 background of z-score data
    tgt = base.clone()
    pred = (base + 0.3 * torch.randn(N, H, W, C, generator=g)).to(pdt)
    pred[0] = tgt[0].to(pdt)                                  # identical images (fp32 case: mse 0 -> psnr inf, ssim 1)
    t_index = torch.tensor([3, 2, 1, 0, 5, 4], dtype=torch.int32)
    out = torch.empty(N, 3, device="cuda")
    K.metrics_recon(tgt.cuda(), pred.cuda(), out, t_index=t_index.cuda())
    o = out.cpu().double()
    for n in range(N):
        m = compute_reconstruction_metrics_single(tgt[int(t_index[n]), :, :, 0].numpy(), pred[n, :, :, 0].float().numpy())
        assert abs(o[n, 0] - m["ssim"]) <= 1e-4, (n, float(o[n, 0]), m["ssim"])
        assert abs(o[n, 2] - m["rmse"]) <= 1e-5 * max(1.0, m["rmse"]), (n, float(o[n, 2]), m["rmse"])
        if np.isfinite(m["psnr"]):
            assert abs(o[n, 1] - m["psnr"]) <= 1e-4 * abs(m["psnr"]) + 1e-4, (n, float(o[n, 1]), m["psnr"])
        else:
            assert not np.isfinite(float(o[n, 1]))
    # without an index, on other channels
    K.metrics_recon(tgt.cuda(), pred.cuda(), out, t_c0=2, p_c0=2)
    o = out.cpu().double()
    m = compute_reconstruction_metrics_single(tgt[1, :, :, 2].numpy(), pred[1, :, :, 2].float().numpy())
    assert abs(o[1, 0] - m["ssim"]) <= 1e-4 and abs(o[1, 1] - m["psnr"]) <= 1e-3
