"""GPU, BASELINE.json full sizes (BraTS 4-contrast slices 28 x 160 x 192, per-GPU batch 16, bf16): properties that do not need
the oracle at that size (it would take minutes on the CPU).
  * linearity of the dominant convolution kernels at the real layer shape (sp6 gamma|beta: 256 images, 16 weight groups),
  * the missing-modality fusion gather -> scatter round trip is exact and row counts equal popcount(mask),
  * a captured training iteration on a batch with random modality dropout (config 3) produces finite losses, a finite
    gradient norm, identical parameters when the same batch is replayed from the same state (up to atomics noise), and the
    reconstruction loss goes down over a few Adam steps."""
import pytest
import torch

import rd_b200.config as rd_config
import rd_b200.data as rd_data
import rd_b200.kernels as K
import rd_b200.ops as ops
from rd_b200.trainer import Trainer, build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_conv_linearity_full_resolution():
    n, h, w, cin, cout, G = 256, 160, 192, 32, 64, 16
    g = torch.Generator(device=DEV).manual_seed(5)
    # operands exactly representable after the sum: small integers / 8
    x1 = (torch.randint(-8, 9, (n, h, w, cin), generator=g, device=DEV).float() / 8).bfloat16()
    x2 = (torch.randint(-8, 9, (n, h, w, cin), generator=g, device=DEV).float() / 8).bfloat16()
    wt = (torch.randint(-4, 5, (G, cout, 9, cin), generator=g, device=DEV).float() / 16).bfloat16()
    d = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, G, 1, 0, 0.2, 0)
    ys = []
    for x in (x1, x2, (x1.float() + x2.float()).bfloat16()):
        y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=DEV)
        K.conv2d_fwd(d, x, wt, None, y)
        ys.append(y.float())
    # fp32 accumulation of exactly representable products: only the final bf16 rounding differs
    err = (ys[2] - (ys[0] + ys[1])).abs().max().item()
    scale = ys[2].abs().max().item()
    assert err <= 2.0 ** -7 * scale + 1e-6, (err, scale)
    # wgrad linearity in dY (halo-resident kernel), fp32 output: exact up to summation order
    dy1 = (torch.randint(-4, 5, (n, h, w, cout), generator=g, device=DEV).float() / 8).bfloat16()
    dy2 = (torch.randint(-4, 5, (n, h, w, cout), generator=g, device=DEV).float() / 8).bfloat16()
    dks = []
    for dy in (dy1, dy2, (dy1.float() + dy2.float()).bfloat16()):
        dK = torch.empty(G, cout, 9, cin, device=DEV)
        K.conv2d_wgrad(d, x1, dy, dK, None)
        dks.append(dK)
    err = (dks[2] - (dks[0] + dks[1])).abs().max().item()
    assert err <= 1e-5 * dks[2].abs().max().item() + 1e-3, err


def test_fusion_round_trip_full_size():
    B, M = 16, 4
    S = torch.randn(M * B, 160, 192, 4, device=DEV).bfloat16()
    gm = torch.Generator().manual_seed(3)
    mask = (torch.rand(B, M, generator=gm) > 0.3).float().to(DEV)
    rows, idx, cnt = ops.fuse_gather(S, mask, B, M)
    k = int(cnt.item())
    assert k == int(mask.sum().item())
    order = [(b, m) for b in range(B) for m in range(M) if mask[b, m] == 1]       # row-major over (b, m), SURVEY Q3
    assert idx[:k].tolist() == [b * M + m for b, m in order]
    for r, (b, m) in enumerate(order[:: max(1, k // 7)]):
        rr = order.index((b, m))
        assert torch.equal(rows[rr], S[m * B + b])
    back = torch.empty_like(S)
    K.fuse_gather_bwd(rows, mask, back, B, M)
    sel = mask.t().reshape(-1).bool()                                              # modality-major stack
    assert torch.equal(back[sel], S[sel]) and float(back[~sel].abs().max() if (~sel).any() else 0) == 0.0


def test_dropout_training_iteration_full_size():
    B, M = 16, 4
    torch.manual_seed(10)
    cfg = rd_config.default_config(precision="bf16", batch_size=B)
    model = build_model(cfg, "cuda:0")
    tr = Trainer(model, cfg, B, use_graph=True)
    batch = rd_data.synthetic_batch(B, M, seed=77, dropoff=True)                   # config 3: random modality dropout
    eps = rd_data.synthetic_eps(B, M, cfg["z_size"], seed=78)
    assert 0 < int(batch["mask"].sum()) < B * M
    first = None
    for it in range(6):                                                            # 2 eager + capture + 3 replays
        tr.train_iteration(batch, eps, (0, 2))
        L = tr.losses_host()
        assert all(v == v and abs(v) < 1e6 for v in L.values()), L
        if first is None:
            first = L
    assert tr.grad_norm_host() == tr.grad_norm_host()                              # finite
    assert L["recon_x"] < first["recon_x"], (first, L)                             # Adam reduces the reconstruction loss
    assert len(tr.graphs) == 1 and max(tr.launches_per_graph.values()) > 300
