"""GPU: the tcgen05 / TMEM implicit-GEMM convolution (forward, dgrad, wgrad) through the C ABI against
F.conv2d on the same bf16-rounded operands (fp32 math) — the 'plain PyTorch fp32 reference' for the one
floating-point tensor-core kernel.  Tolerance: bf16 output rounding (2^-8 relative to the tensor scale)
plus fp32 accumulation-order noise.  The first test runs a tiny launch in a SUBPROCESS with a timeout so a
protocol bug shows up as a failure, not a wedged test session."""
import os
import subprocess
import sys

import pytest
import torch

from tests import emul
import rd_b200.kernels as K
from rd_b200.lib import RD_ALGO_TCGEN05, RD_ALGO_DIRECT, last_conv_algo

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CANARY = r"""
import sys, torch
sys.path.insert(0, %r)
import rd_b200.kernels as K
from tests import emul
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 8, 8, 64, generator=g).bfloat16()
w = (torch.randn(1, 32, 9, 64, generator=g) * 0.05).bfloat16()
d = K.conv_desc(2, 8, 8, 64, 32, 3, 3, 1, 1, 1, 1, 0, 0.2, 2)
y = torch.empty(2, 8, 8, 32, dtype=torch.bfloat16, device='cuda')
K.conv2d_fwd(d, x.cuda(), w.cuda(), None, y)
torch.cuda.synchronize()
yc = torch.empty(2, 8, 8, 32, dtype=torch.bfloat16)
emul.conv2d_fwd(d, x, w, None, yc)
err = (y.cpu().float() - yc.float()).abs().max().item()
print('canary max err', err, 'scale', yc.float().abs().max().item())
assert err < 0.03 * yc.float().abs().max().item() + 1e-3, err
"""


def test_tc_canary_subprocess():
    r = subprocess.run([sys.executable, "-c", _CANARY % ROOT], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, "tcgen05 canary failed:\n" + r.stdout[-2000:] + r.stderr[-4000:]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).bfloat16()


def _close(a, b, rtol, atol, what):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item()
    assert err <= atol + rtol * scale, "%s: max err %.4e scale %.4e" % (what, err, scale)


TC_CASES = [
    # n, h, w, cin, cout, k, stride, pad, groups, act    (the shape families of SURVEY §8a at reduced extent)
    (2, 8, 8, 64, 32, 3, 1, 1, 1, 0),          # canary shape
    (4, 20, 24, 128, 128, 3, 1, 1, 2, 0),      # sp3 gamma/beta/out
    (2, 20, 24, 128, 256, 3, 1, 1, 2, 0),      # fused gamma|beta, 2 N tiles
    (3, 16, 12, 32, 32, 3, 1, 1, 3, 0),        # sp6 family, groups of 1 image
    (2, 16, 12, 32, 16, 3, 1, 1, 1, 0),        # sp6 out
    (2, 10, 12, 64, 64, 3, 1, 1, 2, 1),        # sp5 + fused LeakyReLU
    (2, 16, 24, 32, 64, 4, 2, 1, 1, 0),        # anatomy enc down_2 (k4 s2)
    (4, 10, 12, 256, 256, 4, 2, 1, 4, 0),      # anatomy enc down_5
    (2, 20, 24, 512, 128, 3, 1, 1, 1, 0),      # anatomy dec up_3
    (2, 12, 16, 16, 32, 3, 2, 1, 2, 1),        # modality enc conv2 (k3 s2) + LeakyReLU
    (2, 5, 6, 128, 128, 3, 1, 1, 2, 0),        # sp1: 30-pixel images, ragged M tile
    (1, 7, 9, 24, 40, 3, 1, 1, 1, 0),          # odd sizes, N tile 48 with 40 valid, K tail
    (2, 8, 8, 64, 64, 1, 1, 0, 1, 0),          # 1x1
    (2, 8, 8, 128, 64, 2, 2, 0, 1, 0),         # attention gate W_x (k2 s2 p0)
    (2, 16, 12, 8, 32, 3, 1, 1, 2, 0),         # si_layers: 4 anatomy channels zero-padded to 8
    (2, 16, 12, 8, 32, 4, 2, 1, 1, 1),         # first encoder conv: 7 image channels zero-padded to 8
    (2, 16, 12, 64, 4, 3, 1, 1, 2, 0),         # anatomy logits: 4 output channels (scalar-store epilogue)
    (4, 16, 12, 16, 7, 1, 1, 0, 4, 0),         # decoder output 1x1: 7 output channels
    (8, 32, 48, 16, 16, 3, 2, 1, 4, 0),        # modality enc conv1 (k3 s2, 16 -> 16): dgrad as four parity-class problems (1 / 2 / 2 / 4 taps)
    (8, 40, 48, 64, 128, 4, 2, 1, 4, 0),       # anatomy enc down_3 (k4 s2): dgrad classes of 2 x 2 taps, two images per group
    (4, 20, 24, 64, 128, 3, 2, 1, 2, 1),       # modality enc conv4 (k3 s2) + LeakyReLU
    # persistent TMA kernel (stride-1 "same" convs with an exact rectangle tiling): swizzle modes, N tiles, image packing
    (8, 5, 6, 128, 128, 3, 1, 1, 2, 0),        # 4 images of 5x6 per 120-row tile
    (2, 16, 16, 48, 64, 3, 1, 1, 1, 0),        # kc = 16 (32-byte swizzle), 3 channel chunks
    (2, 16, 16, 96, 32, 3, 1, 1, 2, 1),        # kc = 32 (64-byte swizzle), fused LeakyReLU
    (2, 8, 16, 128, 256, 3, 1, 1, 1, 0),       # one 256-wide N tile (fused gamma|beta of a 128-channel block)
    (2, 8, 16, 64, 320, 3, 1, 1, 2, 0),        # two N tiles, the second one partial
    (6, 160, 192, 32, 64, 3, 1, 1, 3, 0),      # full resolution, many tiles per CTA (persistent loop, TMEM double buffer)
    (4, 10, 12, 128, 256, 3, 1, 1, 2, 0),      # sp2 gamma|beta: wgrad K tiles of 12 x 4 pixels, the last row block overhangs the image
    (4, 10, 12, 16, 128, 3, 1, 1, 2, 0),       # sp2 si_layers on the same tiling
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tc_fwd_dgrad_wgrad(case):
    n, h, w, cin, cout, k, st, pad, G, act = case
    x = _rand((n, h, w, cin), 1)
    packed = _rand((G, cout, k * k, cin), 2, 1.0 / (k * k * cin) ** 0.5)
    packedT = packed.float().permute(0, 3, 2, 1).contiguous().bfloat16()
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(3))
    d = K.conv_desc(n, h, w, cin, cout, k, k, st, pad, G, 1, act, 0.2, RD_ALGO_TCGEN05)
    y = torch.empty(n, d.oh, d.ow, cout, dtype=torch.bfloat16, device=DEV)
    K.conv2d_fwd(d, x.to(DEV), packed.to(DEV), bias.to(DEV), y)
    assert last_conv_algo(0) == RD_ALGO_TCGEN05
    yc = torch.empty(n, d.oh, d.ow, cout, dtype=torch.bfloat16)
    emul.conv2d_fwd(d, x, packed, bias, yc)
    _close(y, yc, 1.0e-2, 2e-3, "tc fwd")
    dy = _rand((n, d.oh, d.ow, cout), 4)
    d.act = 0
    if cout % 8:
        # backward of a layer whose channel count is not a multiple of 8: dY and the transposed weights are zero-padded
        # (exactly what rd_b200.ops._GroupedConv.backward does)
        cp = (cout + 7) // 8 * 8
        dyp = torch.zeros(n, d.oh, d.ow, cp, dtype=torch.bfloat16)
        dyp[..., :cout] = dy
        pT = torch.zeros(G, cin, k * k, cp, dtype=torch.bfloat16)
        pT[..., :cout] = packedT
        dy, packedT, cout = dyp, pT, cp
        d = K.conv_desc(n, h, w, cin, cout, k, k, st, pad, G, 1, 0, 0.2, RD_ALGO_TCGEN05)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=DEV)
    K.conv2d_dgrad(d, dy.to(DEV), packedT.to(DEV), dx)
    assert last_conv_algo(0) == RD_ALGO_TCGEN05
    dxc = torch.empty(n, h, w, cin, dtype=torch.bfloat16)
    emul.conv2d_dgrad(d, dy, packedT, dxc)
    _close(dx, dxc, 1.0e-2, 2e-3, "tc dgrad")
    # wgrad: AUTO picks the tcgen05 kernel where it supports the shape, else the direct one
    d.algo = 0
    dK = torch.empty(G, cout, k * k, cin, device=DEV)
    db = torch.ones(cout, device=DEV)
    K.conv2d_wgrad(d, x.to(DEV), dy.to(DEV), dK, db)
    assert last_conv_algo(0) == RD_ALGO_TCGEN05
    dKc, dbc = torch.empty(G, cout, k * k, cin), torch.ones(cout)
    emul.conv2d_wgrad(d, x, dy, dKc, dbc)
    _close(dK, dKc, 2e-3, 1e-3, "wgrad")
    _close(db, dbc, 2e-3, 2e-3, "dbias (+= through the ones column of the wgrad GEMM)")


def test_conv_tc_matches_direct_kernel_full_res():
    """One full-resolution layer (32->32 3x3 @160x192, 4 groups): tcgen05 vs the CUDA-core kernel, same operands."""
    n, h, w, c = 4, 160, 192, 32
    x, packed = _rand((n, h, w, c), 5).to(DEV), _rand((4, c, 9, c), 6, 0.06).to(DEV)
    y1 = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
    y2 = torch.empty_like(y1)
    d = K.conv_desc(n, h, w, c, c, 3, 3, 1, 1, 4, 1, 0, 0.2, RD_ALGO_TCGEN05)
    K.conv2d_fwd(d, x, packed, None, y1)
    d.algo = RD_ALGO_DIRECT
    K.conv2d_fwd(d, x, packed, None, y2)
    _close(y1, y2, 1e-2, 2e-3, "tc vs direct")


# halo-tile kernel (rd_conv_halo.cu), forced with RD_ALGO_HALO so that small shapes exercise it too
HALO_CASES = [
    # n, h, w, cin, cout, groups, act
    (2, 16, 8, 32, 64, 1, 0),        # exactly one tile per image
    (3, 40, 48, 32, 16, 3, 0),       # sp6-out family; H = 2.5 tiles (ragged last tile row)
    (2, 24, 20, 64, 128, 2, 0),      # kc = 64, ragged in both H and W
    (2, 32, 16, 128, 64, 1, 0),      # two 64-channel chunks per tile (sp4 out), 147 KB of resident weights
    (2, 16, 16, 48, 40, 1, 0),       # kc = 16 (32-byte swizzled weight boxes), N tile 48 with 40 valid channels
    (2, 16, 16, 16, 32, 2, 1),       # si_layers family (4 anatomy channels zero-padded to 16) + LeakyReLU
    (2, 16, 16, 64, 4, 2, 0),        # anatomy logits: 4 output channels (scalar-store epilogue)
    (5, 160, 192, 32, 64, 5, 0),     # full resolution: 9 tiles per CTA, group boundaries inside a CTA's tile range
    (3, 80, 96, 64, 128, 3, 0),      # sp5 gamma|beta
    (2, 24, 28, 32, 32, 2, 0),       # two tiles per stage (even tile count per row), second tile of the last pair half outside the image
    (2, 32, 48, 64, 32, 1, 0),       # two tiles per stage with 64 input channels (sp5-out family)
    (4, 48, 32, 16, 16, 4, 1),       # two tiles per stage, 16 -> 16, group change every image
    (32, 80, 96, 128, 64, 4, 0),     # 13 tiles per CTA, two chunks per tile, a ring too short for two issuers (single-issuer path at depth)
    (16, 80, 96, 64, 128, 4, 0),     # 6-7 one-tile stages per CTA: the two MMA issuers alternate tiles, group boundaries inside a CTA's range
    (48, 160, 192, 16, 32, 3, 1),    # 62 two-tile stages per CTA: one issuer per tile of a stage; wgrad with an accumulator set per issuer
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo_fwd_dgrad(case):
    from rd_b200.lib import RD_ALGO_HALO
    n, h, w, cin, cout, G, act = case
    k = 3
    x = _rand((n, h, w, cin), 11)
    packed = _rand((G, cout, k * k, cin), 12, 1.0 / (k * k * cin) ** 0.5)
    packedT = packed.float().permute(0, 3, 2, 1).contiguous().bfloat16()
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(13))
    d = K.conv_desc(n, h, w, cin, cout, k, k, 1, 1, G, 1, act, 0.2, RD_ALGO_HALO)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=DEV)
    K.conv2d_fwd(d, x.to(DEV), packed.to(DEV), bias.to(DEV), y)
    yc = torch.empty(n, h, w, cout, dtype=torch.bfloat16)
    emul.conv2d_fwd(d, x, packed, bias, yc)
    _close(y, yc, 1.0e-2, 2e-3, "halo fwd")
    if cout % 16:
        cp = (cout + 15) // 16 * 16          # rd_b200.ops pads dY / the transposed weights of narrow layers to 16
        dy = torch.zeros(n, h, w, cp, dtype=torch.bfloat16)
        dy[..., :cout] = _rand((n, h, w, cout), 14)
        pT = torch.zeros(G, cin, k * k, cp, dtype=torch.bfloat16)
        pT[..., :cout] = packedT
        packedT, cout = pT, cp
    else:
        dy = _rand((n, h, w, cout), 14)
    d = K.conv_desc(n, h, w, cin, cout, k, k, 1, 1, G, 1, 0, 0.2, RD_ALGO_HALO)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=DEV)
    K.conv2d_dgrad(d, dy.to(DEV), packedT.to(DEV), dx)
    dxc = torch.empty(n, h, w, cin, dtype=torch.bfloat16)
    emul.conv2d_dgrad(d, dy, packedT, dxc)
    _close(dx, dxc, 1.0e-2, 2e-3, "halo dgrad")
    mt = (3 * (cin // 8) + 15) // 16
    if cin in (16, 32, 64, 128) and cout in (16, 32, 64, 128) and 3 * mt * cout <= 1024 and (3 * mt * cout <= 512 or (cout >= 32 and cin <= 64)):
        # halo-resident wgrad (k_wgrad_halo), forced; bias gradient summed from the dY tiles in shared memory
        dK = torch.empty(G, cout, k * k, cin, device=DEV)
        db = torch.ones(cout, device=DEV)
        K.conv2d_wgrad(d, x.to(DEV), dy.to(DEV), dK, db)
        dKc, dbc = torch.empty(G, cout, k * k, cin), torch.ones(cout)
        emul.conv2d_wgrad(d, x, dy, dKc, dbc)
        _close(dK, dKc, 2e-3, 1e-3, "halo wgrad")
        _close(db, dbc, 2e-3, 2e-3, "halo dbias")


@pytest.mark.parametrize("case", [(6, 40, 44, 32, 64, 3), (4, 32, 24, 32, 32, 2), (4, 32, 32, 64, 16, 4), (3, 48, 16, 16, 128, 1)])
def test_wgrad_halo_dy_tile_layouts(case, monkeypatch):
    """k_wgrad_halo with the dY tile as MN-major SWIZZLED rows ([pixel][64 / 32 / 16 channels], one 4-D TMA box per channel block;
    RD_B200_WGH_DYSW=2 forces it for every row width) against the no-swizzle [row][block][column] planes (= 0): the MMAs see the same
    operands in the same K order, so dK agrees to fp32 rounding of the split-K reductions; the bias gradient is summed in a different order."""
    from rd_b200.lib import RD_ALGO_HALO
    n, h, w, cin, cout, G = case
    x, dy = _rand((n, h, w, cin), 41).to(DEV), _rand((n, h, w, cout), 42).to(DEV)
    d = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, G, 1, 0, 0.2, RD_ALGO_HALO)
    outs = []
    for mode in ("0", "2"):
        monkeypatch.setenv("RD_B200_WGH_DYSW", mode)
        dK = torch.empty(G, cout, 9, cin, device=DEV)
        db = torch.zeros(cout, device=DEV)
        K.conv2d_wgrad(d, x, dy, dK, db)
        outs.append((dK, db))
    _close(outs[1][0], outs[0][0], 2e-6, 1e-6, "dK, swizzled dY rows")      # split-K partial sums meet in red.global.add: order varies
    _close(outs[1][1], outs[0][1], 2e-3, 2e-3, "dbias, swizzled dY rows")


@pytest.mark.parametrize("case", [
    # n, h, w, C, groups, bias rows
    (2, 16, 8, 32, 1, 0),          # one tile per image; C = 32: staged TMA-store epilogue
    (4, 32, 32, 32, 4, 2),         # two tiles per stage, group / image changes inside a CTA's range, one bias row per module
    (3, 40, 44, 32, 3, 0),         # ragged in H and W (second tile of the last pair partly outside, clipped by the TMA store)
    (3, 32, 24, 64, 3, 0),         # C = 64 (sp5): two 32-channel blocks; staging does not fit next to 147 KB of weights -> direct stores
    (16, 160, 192, 32, 16, 4),     # sp6 at full size, 16 weight groups in 4 modules
])
def test_conv_halo_spade_epilogue(case):
    """rd_conv2d_fwd_spade (gamma|beta convolution with the SPADE modulation in its epilogue) against convolution + modulation in torch;
    then rd_spade_modulate_bwd_g (gamma as its own tensor) against rd_spade_modulate_bwd on the concatenated gamma|beta."""
    from rd_b200.lib import RD_ALGO_HALO
    n, h, w, Cz, G, bg = case
    cin, cout = Cz, 2 * Cz
    gen = torch.Generator().manual_seed(21)
    x = _rand((n, h, w, cin), 31)
    z = (_rand((n, h, w, Cz), 32).float() * 1.7 + 0.3).bfloat16()
    packed = _rand((G, cout, 9, cin), 33, 1.0 / (9 * cin) ** 0.5)
    bias = torch.randn((bg, cout) if bg > 1 else (cout,), generator=gen)
    d = K.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, G, 1, 0, 0.2, RD_ALGO_HALO if n * h * w < 100000 else 0, bg)
    assert K.conv2d_fwd_spade_supported(d, x.to(DEV))
    zg = z.to(DEV)
    mean, invstd = torch.empty(n * Cz, device=DEV), torch.empty(n * Cz, device=DEV)
    ws = K.norm_workspace(n, h * w, Cz, torch.device(DEV))
    K.norm_stats(zg, n, h * w, Cz, 1e-5, ws, mean, invstd, None, None, None, 0.0)
    gamma = torch.full((n, h, w, Cz), 9.0, dtype=torch.bfloat16, device=DEV)
    mix = torch.full((n, h, w, Cz), 9.0, dtype=torch.bfloat16, device=DEV)
    K.conv2d_fwd_spade(d, x.to(DEV), packed.to(DEV), bias.to(DEV), zg, mean, invstd, gamma, mix)
    gc, mc = torch.empty(n, h, w, Cz, dtype=torch.bfloat16), torch.empty(n, h, w, Cz, dtype=torch.bfloat16)
    emul.conv2d_fwd_spade(d, x, packed, bias, z, mean.cpu(), invstd.cpu(), gc, mc)
    _close(gamma, gc, 1.0e-2, 2e-3, "spade gamma")
    _close(mix, mc, 1.0e-2, 4e-3, "spade mix")
    if n * h * w > 100000:
        return
    dmix = _rand((n, h, w, Cz), 34).to(DEV)
    gb = torch.cat([gamma, torch.zeros_like(gamma)], -1).contiguous()
    outs = []
    for fn, garg in ((K.spade_modulate_bwd, gb), (K.spade_modulate_bwd_g, gamma)):
        dz = torch.empty_like(zg)
        dgb = torch.empty(n, h, w, 2 * Cz, dtype=torch.bfloat16, device=DEV)
        fn(zg, mean, invstd, garg, dmix, dz, dgb, K.spade_bwd_workspace(zg))
        outs.append((dz, dgb))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("case", [
    # n, h, w, cin, cout, k, stride, pad, groups, bias rows, algo     (module-batched launches: one bias row per module)
    (8, 16, 16, 32, 64, 3, 1, 1, 4, 2, 3),        # halo kernel (forced)
    (8, 20, 24, 128, 256, 3, 1, 1, 4, 4, 2),      # TMA kernel, 256-wide N tile
    (8, 16, 12, 16, 7, 1, 1, 0, 4, 2, 2),         # 1x1 decoder output, 7 channels (scalar-store epilogue)
    (8, 16, 24, 32, 64, 4, 2, 1, 4, 2, 2),        # stride-2 gather kernel
    (4, 12, 8, 8, 16, 3, 1, 1, 4, 2, 1),          # CUDA-core kernel
    (8, 24, 20, 16, 16, 1, 1, 0, 4, 2, 2),        # 1x1 16 -> 16 (padded decoder output): streaming CUDA-core wgrad (k_wgrad_1x1_c16)
    (6, 33, 17, 16, 16, 1, 1, 0, 3, 1, 2),        # same, ragged pixel chunks, one bias row
    (8, 16, 12, 16, 7, 1, 1, 0, 4, 2, 0),         # 1x1 16 -> 7, AUTO: streaming CUDA-core kernel (k_conv1x1_c16<7>)
    (6, 40, 48, 16, 16, 1, 1, 0, 3, 3, 0),        # 1x1 16 -> 16, AUTO: k_conv1x1_c16<16> forward and dgrad, several 256-pixel rounds
])
def test_conv_bias_rows_per_module(case):
    n, h, w, cin, cout, k, st, pad, G, R, algo = case
    x = _rand((n, h, w, cin), 31)
    packed = _rand((G, cout, k * k, cin), 32, 1.0 / (k * k * cin) ** 0.5)
    bias = torch.randn(R, cout, generator=torch.Generator().manual_seed(33))
    d = K.conv_desc(n, h, w, cin, cout, k, k, st, pad, G, 1, 0, 0.2, algo, R)
    y = torch.empty(n, d.oh, d.ow, cout, dtype=torch.bfloat16, device=DEV)
    K.conv2d_fwd(d, x.to(DEV), packed.to(DEV), bias.to(DEV), y)
    yc = torch.empty(n, d.oh, d.ow, cout, dtype=torch.bfloat16)
    emul.conv2d_fwd(d, x, packed, bias, yc)
    _close(y, yc, 1.0e-2, 2e-3, "fwd with per-module bias rows")
    if cout % 8 == 0:
        dy = _rand((n, d.oh, d.ow, cout), 34)
        d.algo = 0 if algo != 1 else 1
        dK = torch.empty(G, cout, k * k, cin, device=DEV)
        db = torch.full((R, cout), 0.5, device=DEV)
        K.conv2d_wgrad(d, x.to(DEV), dy.to(DEV), dK, db)
        dKc, dbc = torch.empty(G, cout, k * k, cin), torch.full((R, cout), 0.5)
        emul.conv2d_wgrad(d, x, dy, dKc, dbc)
        _close(dK, dKc, 2e-3, 1e-3, "wgrad")
        _close(db, dbc, 2e-3, 2e-3, "dbias rows (+=)")
        if k == 1:
            packedT = packed.float().permute(0, 3, 2, 1).contiguous().bfloat16()
            dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=DEV)
            K.conv2d_dgrad(d, dy.to(DEV), packedT.to(DEV), dx)
            dxc = torch.empty(n, h, w, cin, dtype=torch.bfloat16)
            emul.conv2d_dgrad(d, dy, packedT, dxc)
            _close(dx, dxc, 1.0e-2, 2e-3, "1x1 dgrad")
