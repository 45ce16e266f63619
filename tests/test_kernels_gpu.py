"""GPU: every CUDA kernel, called through the C ABI (rd_b200.kernels -> ctypes -> librd_b200.so), against the
plain-PyTorch statement of its contract in tests/emul.py on the same seeded inputs.
Tolerances: fp32 storage 1e-4 relative (reductions in different order), bf16 storage 2^-8 relative on the
rounded result; integer / index outputs bit-exact."""
import math

import pytest
import torch

from tests import emul
import rd_b200.kernels as K

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _close(a, b, rtol, atol, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = b.abs().max().item() if b.numel() else 0.0
    assert err <= atol + rtol * scale, "%s: max err %.4e, scale %.4e" % (what, err, scale)


def _tol(dt):
    return (2e-4, 1e-5) if dt == torch.float32 else (1.2e-2, 1e-3)


def _rand(shape, dt, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dt)


DTS = [torch.float32, torch.bfloat16]


@pytest.mark.parametrize("dt", DTS)
def test_layout_and_cast(dt):
    src = _rand((3, 28, 16, 24), torch.float32, 1)
    d_g = torch.empty(3, 16, 24, 7, dtype=dt, device=DEV)
    d_c = torch.empty(3, 16, 24, 7, dtype=dt)
    K.nchw_to_nhwc(src.to(DEV), d_g, 14, 7)
    emul.nchw_to_nhwc(src, d_c, 14, 7)
    _close(d_g, d_c, 0, 0, "nchw_to_nhwc")
    sl = src.to(DEV)[:, 7:14]
    K.nchw_to_nhwc_strided(sl, d_g, 28)
    emul.nchw_to_nhwc_strided(src[:, 7:14], d_c, 28)
    _close(d_g, d_c, 0, 0, "nchw_to_nhwc_strided")
    back_g = torch.empty(3, 7, 16, 24, device=DEV)
    back_c = torch.empty(3, 7, 16, 24)
    K.nhwc_to_nchw(d_g, back_g)
    emul.nhwc_to_nchw(d_c, back_c)
    _close(back_g, back_c, 0, 0, "nhwc_to_nchw")
    a, b = _rand((2, 5, 6, 8), dt, 2), _rand((2, 5, 6, 12), dt, 3)
    o_g = torch.empty(2, 5, 6, 20, dtype=dt, device=DEV)
    o_c = torch.empty(2, 5, 6, 20, dtype=dt)
    K.concat_channels(a.to(DEV), b.to(DEV), o_g)
    emul.concat_channels(a, b, o_c)
    _close(o_g, o_c, 0, 0, "concat")
    a2, b2 = torch.empty_like(a, device=DEV), torch.empty_like(b, device=DEV)
    K.split_channels(o_g, a2, b2, 8, 12)
    _close(a2, a, 0, 0, "split a")
    _close(b2, b, 0, 0, "split b")
    a, b = _rand((2, 5, 6, 16), dt, 4), _rand((2, 5, 6, 24), dt, 5)         # 16-byte vector path for both dtypes
    o_g = torch.empty(2, 5, 6, 40, dtype=dt, device=DEV)
    o_c = torch.empty(2, 5, 6, 40, dtype=dt)
    K.concat_channels(a.to(DEV), b.to(DEV), o_g)
    emul.concat_channels(a, b, o_c)
    _close(o_g, o_c, 0, 0, "concat (vector)")
    a2, b2 = torch.empty_like(a, device=DEV), torch.empty_like(b, device=DEV)
    K.split_channels(o_g, a2, b2, 16, 24)
    _close(a2, a, 0, 0, "split a (vector)")
    _close(b2, b, 0, 0, "split b (vector)")
    y = torch.empty_like(a, device=DEV)
    K.add(a.to(DEV), a.to(DEV), y)
    _close(y, (a.float() * 2).to(dt), 0, 0, "add")
    x7 = _rand((2, 5, 6, 7), dt, 9)
    pg, pc = torch.empty(2, 5, 6, 8, dtype=dt, device=DEV), torch.empty(2, 5, 6, 8, dtype=dt)
    K.pad_channels(x7.to(DEV), pg)
    emul.pad_channels(x7, pc)
    _close(pg, pc, 0, 0, "pad_channels")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("cond", [True, False])
def test_condconv_mix(dt, cond):
    E, O, I_, kh, kw = (3, 16, 8, 3, 3) if cond else (1, 16, 8, 3, 3)
    W = _rand((E, O, I_, kh, kw), torch.float32, 4) if cond else _rand((O, I_, kh, kw), torch.float32, 4)
    fcw, fcb = (_rand((3, 1), torch.float32, 5), _rand((3,), torch.float32, 6)) if cond else (None, None)
    for types, i_pad in (([1.0, 2.0, 4.0] if cond else [0.0], I_), ([float(t % 4 + 1) for t in range(12)] if cond else [0.0], 16)):
        G, o_total, oT_total, o_off = len(types), 24, 32, 8
        outs = []
        for dev, mod in ((DEV, K), ("cpu", emul)):
            packed = torch.zeros(G, o_total, kh * kw, i_pad, dtype=dt, device=dev)
            packedT = torch.zeros(G, i_pad, kh * kw, oT_total, dtype=dt, device=dev)
            r = torch.zeros(G, E, device=dev)
            mod.condconv_mix_fwd(W.to(dev), None if fcw is None else fcw.to(dev), None if fcb is None else fcb.to(dev), types,
                                 i_pad, o_total, oT_total, o_off, packed, packedT, r)
            dK = _rand((G, o_total, kh * kw, i_pad), torch.float32, 7).to(dev)
            dW = torch.ones_like(W, device=dev)
            dfw = torch.ones(3, 1, device=dev) if cond else None
            dfb = torch.ones(3, device=dev) if cond else None
            mod.condconv_mix_bwd(dK, W.to(dev), None if fcw is None else fcw.to(dev), None if fcb is None else fcb.to(dev), types,
                                 i_pad, o_total, o_off, dW, dfw, dfb)
            outs.append((packed, packedT, r, dW, dfw, dfb))
        rt, at = _tol(dt)
        names = ["packed", "packedT", "r", "dW", "dfc_w", "dfc_b"]
        for n, a, b in zip(names, outs[0], outs[1]):
            if a is not None:
                _close(a, b, rt if n.startswith("packed") else 3e-4, at, n)
    outs = [[None] * 6, [None] * 6]
    rt, at = _tol(dt)
    names = ["packed", "packedT", "r", "dW", "dfc_w", "dfc_b"]
    for n, a, b in zip(names, outs[0], outs[1]):
        if a is not None:
            _close(a, b, rt if n.startswith("packed") else 2e-4, at, n)


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, groups, act
    (2, 12, 10, 7, 32, 4, 2, 1, 1, 1),
    (4, 9, 11, 4, 16, 3, 1, 1, 2, 0),
    (2, 8, 8, 16, 7, 1, 1, 0, 2, 0),
    (2, 10, 12, 24, 40, 3, 2, 1, 1, 1),
    (3, 6, 6, 8, 1, 2, 2, 0, 1, 0),
]


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_direct(dt, case):
    n, h, w, cin, cout, k, st, pad, G, act = case
    x = _rand((n, h, w, cin), dt, 10)
    packed = _rand((G, cout, k * k, cin), dt, 11, 0.2)
    packedT = packed.float().permute(0, 3, 2, 1).contiguous().to(dt)
    bias = _rand((cout,), torch.float32, 12)
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        d = K.conv_desc(n, h, w, cin, cout, k, k, st, pad, G, K._dt(x), act, 0.2, 1)
        y = torch.empty(n, d.oh, d.ow, cout, dtype=dt, device=dev)
        mod.conv2d_fwd(d, x.to(dev), packed.to(dev), bias.to(dev), y)
        dy = _rand((n, d.oh, d.ow, cout), dt, 13).to(dev)
        dx = torch.empty(n, h, w, cin, dtype=dt, device=dev)
        d.act = 0
        mod.conv2d_dgrad(d, dy, packedT.to(dev), dx)
        dK = torch.empty(G, cout, k * k, cin, device=dev)
        db = torch.ones(cout, device=dev)
        mod.conv2d_wgrad(d, x.to(dev), dy, dK, db)
        res.append((y, dx, dK, db))
    rt, at = _tol(dt)
    for nme, a, b in zip(["y", "dx", "dK", "dbias"], res[0], res[1]):
        _close(a, b, rt if nme in ("y", "dx") else 3e-4, at, nme)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("G,N,HW,C", [(1, 2, (9, 7), 32), (4, 8, (5, 6), 48), (2, 2, (40, 50), 7), (6, 6, (16, 16), 4)])
def test_norm_train_fwd_bwd(dt, G, N, HW, C):
    H, W = HW
    x = (_rand((N, H, W, C), torch.float32, 20) * 2 + 3).to(dt)
    dy = _rand((N, H, W, C), dt, 21)
    wt, bs = _rand((C,), torch.float32, 22) + 1, _rand((C,), torch.float32, 23)
    ppg = (N // G) * H * W
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        mean, invstd = torch.empty(G * C, device=dev), torch.empty(G * C, device=dev)
        rm, rv = torch.zeros(C, device=dev) + 0.5, torch.ones(C, device=dev) * 2
        nbt = torch.zeros((), dtype=torch.int64, device=dev)
        ws = mod.norm_workspace(G, ppg, C, dev)
        mod.norm_stats(x.to(dev), G, ppg, C, 1e-5, ws, mean, invstd, rm, rv, nbt, 0.1)
        y = torch.empty_like(x, device=dev)
        mod.norm_apply(x.to(dev), mean, invstd, wt.to(dev), bs.to(dev), y, G, ppg, C)
        dx = torch.empty_like(x, device=dev)
        dw, db = torch.ones(C, device=dev), torch.ones(C, device=dev)
        ws2 = mod.norm_workspace(G, ppg, C, dev)
        mod.norm_bwd(x.to(dev), dy.to(dev), mean, invstd, wt.to(dev), dx, dw, db, ws2, G, ppg, C)
        res.append((mean, invstd, rm, rv, nbt.float(), y, dx, dw, db))
    rt, at = _tol(dt)
    names = ["mean", "invstd", "running_mean", "running_var", "nbt", "y", "dx", "dweight", "dbias"]
    for nme, a, b in zip(names, res[0], res[1]):
        _close(a, b, rt if nme in ("y", "dx") else 5e-4, at if nme in ("y", "dx") else 1e-5, nme)


@pytest.mark.parametrize("dt", DTS)
def test_norm_eval_and_instance(dt):
    C, G = 16, 3
    rm, rv = _rand((C,), torch.float32, 30), _rand((C,), torch.float32, 31).abs() + 0.5
    mg, ig = torch.empty(G * C, device=DEV), torch.empty(G * C, device=DEV)
    mc, ic = torch.empty(G * C), torch.empty(G * C)
    K.norm_eval_stats(rm.to(DEV), rv.to(DEV), G, 1e-5, mg, ig)
    emul.norm_eval_stats(rm, rv, G, 1e-5, mc, ic)
    _close(mg, mc, 1e-6, 0, "eval mean")
    _close(ig, ic, 1e-5, 0, "eval invstd")
    # SPADE modulation = instance norm (one group per image) + (1+gamma), beta
    N, H, W = 3, 10, 12
    z, gb, dmix = _rand((N, H, W, C), dt, 32), _rand((N, H, W, 2 * C), dt, 33), _rand((N, H, W, C), dt, 34)
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        mean, invstd = torch.empty(N * C, device=dev), torch.empty(N * C, device=dev)
        ws = mod.norm_workspace(N, H * W, C, dev)
        mod.norm_stats(z.to(dev), N, H * W, C, 1e-5, ws, mean, invstd, None, None, None, 0.0)
        mix = torch.empty_like(z, device=dev)
        mod.spade_modulate_fwd(z.to(dev), mean, invstd, gb.to(dev), mix)
        dz, dgb = torch.empty_like(z, device=dev), torch.empty_like(gb, device=dev)
        ws2 = mod.spade_bwd_workspace(z.to(dev))
        mod.spade_modulate_bwd(z.to(dev), mean, invstd, gb.to(dev), dmix.to(dev), dz, dgb, ws2)
        res.append((mix, dz, dgb))
    rt, at = _tol(dt)
    for nme, a, b in zip(["mix", "dz", "dgb"], res[0], res[1]):
        _close(a, b, rt, at * 4, nme)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("case", [(3, 10, 12, 16), (5, 37, 29, 32), (2, 160, 192, 32), (4, 20, 24, 128), (3, 5, 6, 128), (2, 9, 8, 256),
                                  (16, 80, 96, 64), (300, 4, 5, 32)])
def test_spade_bwd_single_pass_equals_two_pass(dt, case, monkeypatch):
    """k_spade_bwd_fused (one pass, per-image counter barrier, chunk kept in registers) against the two-pass form
    (RD_B200_SPADE_BWD_FUSED=0) and against the torch statement: d(gamma|beta) bit-exact (same per-element arithmetic), dz within
    the storage tolerance (the per-image sums fold the chunks in a different, fixed order).  20 launches in a row walk the 8 counter
    slots more than twice: every launch must leave its slot zeroed, and every launch must give the same bits (deterministic)."""
    n, h, w, C = case
    z, dmix = _rand((n, h, w, C), dt, 41, 1.3), _rand((n, h, w, C), dt, 42)
    gamma = _rand((n, h, w, C), dt, 43, 0.5)
    zg, dg, gg = z.to(DEV), dmix.to(DEV), gamma.to(DEV)
    mean, invstd = torch.empty(n * C, device=DEV), torch.empty(n * C, device=DEV)
    K.norm_stats(zg, n, h * w, C, 1e-5, K.norm_workspace(n, h * w, C, torch.device(DEV)), mean, invstd, None, None, None, 0.0)

    def run():
        dz = torch.full_like(zg, 7.0)
        dgb = torch.full((n, h, w, 2 * C), 7.0, dtype=dt, device=DEV)
        K.spade_modulate_bwd_g(zg, mean, invstd, gg, dg, dz, dgb, K.spade_bwd_workspace(zg))
        return dz, dgb

    monkeypatch.setenv("RD_B200_SPADE_BWD_FUSED", "0")
    dz2, dgb2 = run()
    monkeypatch.setenv("RD_B200_SPADE_BWD_FUSED", "1")
    dz1, dgb1 = run()
    assert torch.equal(dgb1, dgb2)
    rt, at = _tol(dt)
    _close(dz1, dz2, rt, at * 4, "dz fused vs two-pass")
    dzc, dgbc = torch.empty_like(z), torch.empty(n, h, w, 2 * C, dtype=dt)
    emul.spade_modulate_bwd_g(z, mean.cpu(), invstd.cpu(), gamma, dmix, dzc, dgbc, None)
    _close(dz1, dzc, rt, at * 4, "dz fused vs torch")
    _close(dgb1, dgbc, rt, at * 4, "dgb fused vs torch")
    for _ in range(20):
        dzr, dgbr = run()
        assert torch.equal(dzr, dz1) and torch.equal(dgbr, dgb1)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("case", [(5, 6, 10, 12, True), (5, 6, 10, 12, False), (160, 192, 5, 6, False), (160, 192, 80, 96, False),
                                  (20, 24, 40, 48, True), (7, 9, 13, 5, False), (3, 3, 3, 3, False), (20, 24, 40, 48, False),
                                  (9, 11, 18, 22, False), (33, 8, 66, 16, False)])
def test_bilinear(dt, case):
    h, w, oh, ow, align = case
    for c in (4, 3, 16):          # 16: the 16-byte vector kernels incl. the x2 specialisation of the backward
        x = _rand((2, h, w, c), dt, 40)
        dy = _rand((2, oh, ow, c), dt, 41)
        yg, yc = torch.empty(2, oh, ow, c, dtype=dt, device=DEV), torch.empty(2, oh, ow, c, dtype=dt)
        K.bilinear_fwd(x.to(DEV), yg, align)
        emul.bilinear_fwd(x, yc, align)
        rt, at = _tol(dt)
        _close(yg, yc, rt, at, "bilinear fwd")
        dg, dc = torch.empty_like(x, device=DEV), torch.empty_like(x)
        K.bilinear_bwd(dy.to(DEV), dg, align)
        emul.bilinear_bwd(dy, dc, align)
        _close(dg, dc, rt, at * 8, "bilinear bwd")


@pytest.mark.parametrize("case", [(20, 24, 40, 48, True, 64), (20, 24, 40, 48, False, 32), (10, 12, 20, 24, True, 256), (9, 7, 23, 19, True, 24),
                                  (9, 7, 23, 19, False, 56), (40, 48, 80, 96, False, 16), (5, 6, 17, 6, True, 8)])
def test_bilinear_rolling_rows_forward_bit_exact(case):
    """k_bilinear_fwd_roll (up-sampling, bf16: horizontal blends of an input row kept in registers and reused by the output rows that read
    it, coordinates from per-block tables) against the generic one-vector-per-thread kernel, which evaluates the same expression:
    bit-identical outputs; and against torch within the bf16 tolerance."""
    import os
    import subprocess
    h, w, oh, ow, align, c = case
    x = _rand((3, h, w, c), torch.bfloat16, 71)
    yg = torch.empty(3, oh, ow, c, dtype=torch.bfloat16, device=DEV)
    K.bilinear_fwd(x.to(DEV), yg, align)
    yc = torch.empty(3, oh, ow, c, dtype=torch.bfloat16)
    emul.bilinear_fwd(x, yc, align)
    _close(yg, yc, 1.2e-2, 1e-3, "bilinear roll vs torch")
    # the generic kernel on the same data: channels-last slices of 4 (c % 8 != 0 path is a different kernel with the same formula)
    x4 = x[..., :4].contiguous()
    y4 = torch.empty(3, oh, ow, 4, dtype=torch.bfloat16, device=DEV)
    K.bilinear_fwd(x4.to(DEV), y4, align)
    assert torch.equal(yg[..., :4].cpu(), y4.cpu())


@pytest.mark.parametrize("dt", DTS)
def test_activations_softmax(dt):
    x, dy = _rand((2, 6, 7, 12), dt, 50), _rand((2, 6, 7, 12), dt, 51)
    rt, at = _tol(dt)
    yg, yc = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.lrelu_fwd(x.to(DEV), yg, 0.2)
    emul.lrelu_fwd(x, yc, 0.2)
    _close(yg, yc, 0, 0, "lrelu")
    dg, dc = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.lrelu_bwd(dy.to(DEV), yg, dg, 0.2)
    emul.lrelu_bwd(dy, yc, dc, 0.2)
    _close(dg, dc, 0, 0, "lrelu bwd")
    # masked softmax: mask broadcast over a stack of 3 "modalities"
    s = _rand((6, 8, 9, 4), dt, 52) * 3
    mask = (torch.rand(2, 8, 9, generator=torch.Generator().manual_seed(53)) > 0.6).float()
    for m in (mask, None):
        pg, pc = torch.empty_like(s, device=DEV), torch.empty_like(s)
        K.masked_softmax_fwd(s.to(DEV), None if m is None else m.to(DEV), pg)
        emul.masked_softmax_fwd(s, m, pc)
        _close(pg, pc, rt, at, "masked softmax")
        dp = _rand(tuple(s.shape), dt, 54)
        dsg, dsc = torch.empty_like(s, device=DEV), torch.empty_like(s)
        K.masked_softmax_bwd(pg, dp.to(DEV), dsg)
        emul.masked_softmax_bwd(pc, dp, dsc)
        _close(dsg, dsc, rt * 2, at * 2, "masked softmax bwd")
    # attention-gate helpers
    a, b = _rand((2, 5, 5, 8), dt, 55), _rand((2, 5, 5, 8), dt, 56)
    for name, ins in (("add_relu_fwd", (a, b)), ("sigmoid_fwd", (a,))):
        og, oc = torch.empty_like(a, device=DEV), torch.empty_like(a)
        getattr(K, name)(*[t.to(DEV) for t in ins], og)
        getattr(emul, name)(*ins, oc)
        _close(og, oc, rt, at, name)
    al = _rand((2, 5, 5, 1), dt, 57)
    og, oc = torch.empty_like(a, device=DEV), torch.empty_like(a)
    K.mul_bcast_fwd(al.to(DEV), a.to(DEV), og)
    emul.mul_bcast_fwd(al, a, oc)
    _close(og, oc, rt, at, "mul_bcast")
    dxg, dag = torch.empty_like(a, device=DEV), torch.empty_like(al, device=DEV)
    dxc, dac = torch.empty_like(a), torch.empty_like(al)
    K.mul_bcast_bwd(al.to(DEV), a.to(DEV), b.to(DEV), dxg, dag)
    emul.mul_bcast_bwd(al, a, b, dxc, dac)
    _close(dxg, dxc, rt, at, "mul_bcast dx")
    _close(dag, dac, rt, at * 4, "mul_bcast dalpha")
    for name in ("relu_bwd", "sigmoid_bwd"):
        yv = _rand((2, 5, 5, 8), dt, 58).abs() * 0.5
        dg2, dc2 = torch.empty_like(a, device=DEV), torch.empty_like(a)
        getattr(K, name)(b.to(DEV), yv.to(DEV), dg2)
        getattr(emul, name)(b, yv, dc2)
        _close(dg2, dc2, rt, at, name)


def test_linear_and_sample():
    x, W, b = _rand((9, 3840), torch.float32, 60), _rand((32, 3840), torch.float32, 61, 0.02), _rand((32,), torch.float32, 62)
    for act in (0, 1):
        yg, yc = torch.empty(9, 32, device=DEV), torch.empty(9, 32)
        K.linear_fwd(x.to(DEV), W.to(DEV), b.to(DEV), yg, act, 0.2)
        emul.linear_fwd(x, W, b, yc, act, 0.2)
        _close(yg, yc, 2e-5, 1e-5, "linear")
    # short reduction, many outputs (zi_scaler 16 -> 3840): thread-per-output kernel
    xs, Ws, bs = _rand((12, 16), torch.float32, 68), _rand((3840, 16), torch.float32, 69, 0.2), _rand((3840,), torch.float32, 70)
    yg, yc = torch.empty(12, 3840, device=DEV), torch.empty(12, 3840)
    K.linear_fwd(xs.to(DEV), Ws.to(DEV), bs.to(DEV), yg, 0, 0.2)
    emul.linear_fwd(xs, Ws, bs, yc, 0, 0.2)
    _close(yg, yc, 2e-5, 1e-5, "linear 16 -> 3840")
    dy = _rand((9, 32), torch.float32, 63)
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        dx, dW, db = torch.empty(9, 3840, device=dev), torch.ones(32, 3840, device=dev), torch.ones(32, device=dev)
        mod.linear_bwd(x.to(dev), W.to(dev), dy.to(dev), dx, dW, db)
        res.append((dx, dW, db))
    for a, c in zip(*res):
        _close(a, c, 2e-5, 1e-5, "linear bwd")
    # many rows: tiled dW kernel (rows staged through shared memory), ragged row / column tiles
    for (rows, inf, outf) in ((70, 16, 200), (64, 3840, 32), (33, 20, 7)):
        x2, W2, dy2 = _rand((rows, inf), torch.float32, 80), _rand((outf, inf), torch.float32, 81, 0.1), _rand((rows, outf), torch.float32, 82)
        res = []
        for dev, mod in ((DEV, K), ("cpu", emul)):
            dx, dW, db = torch.empty(rows, inf, device=dev), torch.ones(outf, inf, device=dev), torch.ones(outf, device=dev)
            mod.linear_bwd(x2.to(dev), W2.to(dev), dy2.to(dev), dx, dW, db)
            res.append((dx, dW, db))
        for a, c in zip(*res):
            _close(a, c, 5e-5, 2e-5, "linear bwd (tiled dW)")
    mu, lv, eps, dz = (_rand((8, 16), torch.float32, s) for s in (64, 65, 66, 67))
    zg, zc = torch.empty(8, 16, device=DEV), torch.empty(8, 16)
    K.sample_fwd(mu.to(DEV), lv.to(DEV), eps.to(DEV), zg)
    emul.sample_fwd(mu, lv, eps, zc)
    _close(zg, zc, 1e-5, 1e-6, "sample")
    a1, a2, c1, c2 = torch.empty(8, 16, device=DEV), torch.empty(8, 16, device=DEV), torch.empty(8, 16), torch.empty(8, 16)
    K.sample_bwd(dz.to(DEV), lv.to(DEV), eps.to(DEV), a1, a2)
    emul.sample_bwd(dz, lv, eps, c1, c2)
    _close(a1, c1, 1e-5, 1e-6, "sample dmu")
    _close(a2, c2, 1e-5, 1e-6, "sample dlv")


@pytest.mark.parametrize("dt", DTS)
def test_fuse_gather_bit_exact(dt):
    """Q3: boolean gather order over all 2^M mask rows incl. empty and full masks — bit-exact rows and indices."""
    B, M = 4, 4
    si = _rand((M * B, 6, 5, 4), dt, 70)
    gm = torch.Generator().manual_seed(71)
    masks = [torch.randint(0, 2, (B, M), generator=gm).float() for _ in range(6)] + [torch.zeros(B, M), torch.ones(B, M)]
    for mask in masks:
        res = []
        for dev, mod in ((DEV, K), ("cpu", emul)):
            out = torch.zeros_like(si, device=dev)
            idx = torch.empty(B * M, dtype=torch.int32, device=dev)
            cnt = torch.zeros(1, dtype=torch.int32, device=dev)
            mod.fuse_gather_fwd(si.to(dev), mask.to(dev), out, idx, cnt, B, M)
            dsi = torch.empty_like(si, device=dev)
            mod.fuse_gather_bwd(out, mask.to(dev), dsi, B, M)
            res.append((out, idx, cnt, dsi))
        for a, b in zip(*res):
            assert torch.equal(a.cpu(), b.cpu())
        # and against torch boolean indexing, the reference's own statement
        ref = torch.stack([si[m * B:(m + 1) * B] for m in range(M)], 1)[mask == 1]
        k = int(res[0][2].item())
        assert k == ref.shape[0] == int(mask.sum())
        assert torch.equal(res[0][0][:k].cpu(), ref)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("p", [1, 2])
def test_recon_and_combine(dt, p):
    B, M = 3, 4
    gt = _rand((M * B, 8, 9, 7), torch.float32, 80)
    gm = torch.Generator().manual_seed(81)
    masks = [torch.randint(0, 2, (B, M), generator=gm).float() for _ in range(6)] + [torch.zeros(B, M), torch.ones(B, M)]
    for kind, R in ((0, M * B), (1, M * (M - 1) * B)):
        x = _rand((R, 8, 9, 7), dt, 82 + kind)
        for mask in masks:
            res = []
            for dev, mod in ((DEV, K), ("cpu", emul)):
                gi = None
                if kind == 1:
                    gi = torch.empty(R, dtype=torch.int32, device=dev)
                    mod.xmix_plan(mask.to(dev), gi, B, M)
                rl = torch.empty(R, device=dev)
                part = torch.empty(R * mod.recon_chunks(8 * 9 * 7), device=dev)
                mod.recon_rows_fwd(x.to(dev), gt.to(dev), gi, rl, part, R, p)
                loss, coef = torch.empty(1, device=dev), torch.empty(R, device=dev)
                mod.masked_combine(rl, mask.to(dev), loss, coef, B, M, kind)
                dx = torch.empty_like(x, device=dev)
                mod.recon_rows_bwd(x.to(dev), gt.to(dev), gi, coef, dx, R, p)
                res.append((gi, rl, loss, coef, dx))
            if kind == 1:
                assert torch.equal(res[0][0].cpu(), res[1][0]), "xmix plan (Q4 index lag) must be bit-exact"
            _close(res[0][1], res[1][1], 2e-4 if dt == torch.float32 else 1e-2, 1e-6, "row loss")
            _close(res[0][2], res[1][2], 2e-4 if dt == torch.float32 else 1e-2, 1e-6, "loss")
            _close(res[0][3], res[1][3], 1e-6, 1e-7, "coef")
            _close(res[0][4], res[1][4], *_tol(dt), "dx")


def test_small_losses():
    B, M, Z = 5, 4, 16
    mu, mun, lv = (_rand((M * B, Z), torch.float32, s) for s in (90, 91, 92))
    gm = torch.Generator().manual_seed(93)
    masks = [torch.randint(0, 2, (B, M), generator=gm).float() for _ in range(8)] + [torch.ones(B, M), torch.zeros(B, M)]
    for mask in masks:
        for name in ("latent_z_loss", "sim_z_loss", "kl_loss"):
            if name == "kl_loss" and mask.sum() == 0:
                continue
            res = []
            for dev, mod in ((DEV, K), ("cpu", emul)):
                loss = torch.empty(1, device=dev)
                g1, g2 = torch.empty(M * B, Z, device=dev), torch.empty(M * B, Z, device=dev)
                if name == "latent_z_loss":
                    mod.latent_z_loss(mu.to(dev), mun.to(dev), mask.to(dev), loss, g1, g2, B, M, Z)
                elif name == "sim_z_loss":
                    mod.sim_z_loss(mu.to(dev), mask.to(dev), 0.1, loss, g1, B, M, Z)
                    g2.zero_()
                else:
                    mod.kl_loss(mu.to(dev), lv.to(dev), mask.to(dev), loss, g1, g2, B, M, Z)
                res.append((loss, g1, g2))
            for a, b in zip(*res):
                _close(a, b, 2e-4, 2e-6, name)


@pytest.mark.parametrize("dt", DTS)
def test_sim_s_and_maxpool(dt):
    B, M, H, W, C = 3, 4, 32, 48, 4
    s = torch.softmax(_rand((M * B, H, W, C), torch.float32, 100) * 2, -1).to(dt)
    res = []
    gm = torch.Generator().manual_seed(101)
    masks = [torch.ones(B, M), torch.randint(0, 2, (B, M), generator=gm).float(), torch.zeros(B, M)]
    D = C * (H // 16) * (W // 16)
    for dev, mod in ((DEV, K), ("cpu", emul)):
        pooled = torch.empty(M * B, D, device=dev)
        arg = torch.empty(M * B, D, dtype=torch.int32, device=dev)
        mod.maxpool16_fwd(s.to(dev), pooled, arg)
        out = [pooled, arg]
        for mask in masks:
            for pair in ((0, 2), (3, 1)):
                loss, dp = torch.empty(1, device=dev), torch.empty(M * B, D, device=dev)
                pr = torch.tensor(pair, dtype=torch.int32).to(dev)
                mod.sim_s_loss(pooled, mask.to(dev), pr, 0.1, loss, dp, B, M, D)
                out += [loss, dp]
        ds = torch.empty_like(s, device=dev)
        mod.maxpool16_bwd(_rand((M * B, D), torch.float32, 102).to(dev), arg, ds)
        out.append(ds)
        res.append(out)
    assert torch.equal(res[0][1].cpu(), res[1][1]), "argmax must be bit-exact"
    for k, (a, b) in enumerate(zip(*res)):
        if k != 1:
            _close(a, b, 3e-4, 2e-6, "sim_s[%d]" % k)


@pytest.mark.parametrize("dt", DTS)
def test_seg_loss(dt):
    N, H, W = 2, 20, 24
    y = _rand((N, H, W, 4), dt, 110)
    tgt = torch.randint(0, 4, (N, H * W), generator=torch.Generator().manual_seed(111)).float()
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        loss, part = torch.empty(1, device=dev), torch.empty(mod.SEG_PARTIAL_FLOATS, device=dev)
        mod.seg_loss_fwd(y.to(dev), tgt.to(dev), loss, part)
        dy = torch.empty_like(y, device=dev)
        mod.seg_loss_bwd(y.to(dev), tgt.to(dev), part, torch.tensor([0.7], device=dev), dy)
        res.append((loss, dy))
    _close(res[0][0], res[1][0], 2e-4, 1e-6, "seg loss")
    _close(res[0][1], res[1][1], *_tol(dt), "seg dy")


def test_optimizer_kernels():
    torch.manual_seed(0)
    n = 200000
    segs = torch.tensor([[0, 65536], [65536, 65536], [140000, 30000]], dtype=torch.int64)
    hyper0 = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 0, 0])
    p0, g0 = torch.randn(n), torch.randn(n) * 0.01
    res = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        p, g = p0.clone().to(dev), g0.clone().to(dev)
        m, v, vm = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        hyper = hyper0.clone().to(dev)
        part, sc = torch.zeros(3, device=dev), torch.zeros(4, device=dev)
        sg = segs.to(dev)
        for it in range(3):
            mod.grad_norm(g, sg, 3, part, sc, 1.0)
            mod.grad_scale(g, sg, 3, sc)
            mod.adam_amsgrad(p, g, m, v, vm, sg, 3, hyper)
            g.mul_(1.5)
        res.append((p, g, m, v, vm, sc[:3], hyper))
    for k, (a, b) in enumerate(zip(*res)):
        _close(a, b, 1e-5, 1e-7, "optimizer[%d]" % k)


def test_clip_adam_fused_equals_three_launches():
    """rd_clip_adam_amsgrad == rd_grad_scale + rd_adam_amsgrad + zero of the segments, bit for bit (aligned and unaligned segments)."""
    torch.manual_seed(1)
    n = 300000
    segs = torch.tensor([[0, 65536], [65536, 65536], [140001, 30003], [200000, 7]], dtype=torch.int64).to(DEV)
    hyper0 = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-5, 0.0, 0, 0])
    p0, g0 = torch.randn(n), torch.randn(n) * 0.05
    res = []
    for fused in (False, True):
        p, g = p0.clone().to(DEV), g0.clone().to(DEV)
        m, v, vm = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        hyper = hyper0.clone().to(DEV)
        part, sc = torch.zeros(4, device=DEV), torch.zeros(4, device=DEV)
        for it in range(3):
            g.copy_(g0.to(DEV) * (1.0 + it))
            K.grad_norm(g, segs, 4, part, sc, 1.0)
            if fused:
                K.clip_adam_amsgrad(p, g, m, v, vm, segs, 4, hyper, sc, True)
            else:
                K.grad_scale(g, segs, 4, sc)
                K.adam_amsgrad(p, g, m, v, vm, segs, 4, hyper)
                for o, l in segs.tolist():
                    g[o:o + l] = 0
        res.append((p, g, m, v, vm, hyper))
    for k, (a, b) in enumerate(zip(*res)):
        assert torch.equal(a, b), "fused clip+adam differs in tensor %d" % k
    assert float(res[1][1][131072:140001].abs().sum()) > 0      # outside the segments: untouched


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("case", [
    # rows per block, trailing shape, c_pad, index                                (the fan-outs of MultimodalModel.decode_nhwc)
    (2, (6, 8, 4), 16, [0, 0, 0, 1, 1, 2, 3, 3]),       # anatomy codes, 4 -> 16 channels, some sources unused / repeated
    (3, (5, 6, 7), 7, [3, 1, 0, 2]),                    # x-hat rows: 7 channels, no padding, odd block size (scalar path)
    (2, (16,), 16, [1, 1, 0, 3, 2, 2, 2]),              # z rows (M*B, 16): vector copy path
    (2, (4, 4, 12), 16, [2, 0]),                        # partial vector: 12 -> 16
    (3, (40, 48, 4), 16, [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3]),      # the 16-decode fan-out at depth (k_gather_pad4_*, many blocks)
    (2, (10, 12, 4), 8, [1, 0, 1]),                     # 4 -> 8 channels
])
def test_gather_blocks(dt, case):
    block, tail, c_pad, index = case
    nsrc = max(index) + 2                                # one source block that nobody references -> zero gradient
    src = _rand((nsrc * block,) + tail, dt, 71)
    out_g = torch.full((len(index) * block,) + tail[:-1] + (c_pad,), 7.0, dtype=dt, device=DEV)
    out_c = torch.empty((len(index) * block,) + tail[:-1] + (c_pad,), dtype=dt)
    K.gather_blocks_fwd(src.to(DEV), out_g, index, block)
    emul.gather_blocks_fwd(src, out_c, index, block)
    _close(out_g, out_c, 0, 0, "gather_blocks fwd (bit exact)")
    dout = _rand(tuple(out_c.shape), dt, 72)
    ds_g = torch.full(tuple(src.shape), 3.0, dtype=dt, device=DEV)
    ds_c = torch.empty(tuple(src.shape), dtype=dt)
    K.gather_blocks_bwd(dout.to(DEV), ds_g, index, block)
    emul.gather_blocks_bwd(dout, ds_c, index, block)
    _close(ds_g, ds_c, 1e-6, 0, "gather_blocks bwd (fp32 accumulation, same order)")


@pytest.mark.parametrize("dt", DTS)
def test_mix_fwd_batched_equals_per_head(dt):
    """rd_condconv_mix_fwd_batched (one launch over a device job table; a block owns units of 16 x 16 channels x taps staged through
    shared memory, 16-byte runs into both packed layouts) against one rd_condconv_mix_fwd launch per head: identical arithmetic, so the
    packed tensors must be equal.  Heads share packed tensors at different o_off; shapes cover ragged O / I, zero-padded input channels
    (I = 4 -> 16), k = 4 (16 taps), 1 x 1, a single expert, unaligned o_off (element-store path) and 7 output channels."""
    g = torch.Generator().manual_seed(97)
    heads = [  # E, O, I, k, G, i_pad, o_total, oT_total, o_off
        (3, 32, 32, 3, 4, 32, 64, 64, 0), (3, 32, 32, 3, 4, 32, 64, 64, 32), (3, 16, 4, 3, 16, 16, 16, 16, 0),
        (1, 8, 24, 4, 1, 24, 8, 16, 0), (3, 7, 16, 1, 4, 16, 16, 16, 0), (3, 40, 72, 3, 3, 80, 48, 48, 4),
        (3, 128, 64, 3, 16, 64, 256, 256, 128)]
    plan = K.MixFwdPlan(DEV)
    refs, bufs = [], []
    for n, (E, O, I, k, G, i_pad, o_total, oT_total, o_off) in enumerate(heads):
        W = torch.randn((E, O, I, k, k) if E > 1 else (O, I, k, k), generator=g).to(DEV)
        fcw = torch.randn(3, 1, generator=g).to(DEV) if E > 1 else None
        fcb = torch.randn(3, generator=g).to(DEV) if E > 1 else None
        types = [float(1 + (t % 4)) for t in range(G)]
        pk = torch.full((G, o_total, k * k, i_pad), 5.0, dtype=dt, device=DEV)
        pkT = torch.full((G, i_pad, k * k, oT_total), 5.0, dtype=dt, device=DEV)
        pk_r, pkT_r = pk.clone(), pkT.clone()
        K.condconv_mix_fwd(W, fcw, fcb, types, i_pad, o_total, oT_total, o_off, pk_r, pkT_r, None)
        refs.append((pk_r, pkT_r))
        bufs.append((pk, pkT))
        plan.register(n, pk, pkT, None, [(W, fcw, fcb, types, i_pad, o_total, oT_total, o_off, pk, pkT, None, None)])
    plan.prepare()
    torch.cuda.synchronize()
    for n, ((pk, pkT), (pk_r, pkT_r)) in enumerate(zip(bufs, refs)):
        assert torch.equal(pk, pk_r), "packed differs for head %d" % n
        assert torch.equal(pkT, pkT_r), "packedT differs for head %d" % n


def test_mix_bwd_batched_equals_per_head():
    """rd_condconv_mix_bwd_batched (one launch, device job table) against one rd_condconv_mix_bwd launch per head."""
    g = torch.Generator().manual_seed(91)
    heads = [  # E, O, I, k, G, i_pad, o_total, o_off
        (3, 32, 32, 3, 4, 32, 64, 0), (3, 32, 32, 3, 4, 32, 64, 32), (3, 16, 4, 3, 16, 16, 16, 0),
        (1, 8, 24, 4, 1, 24, 8, 0), (3, 7, 16, 1, 4, 16, 16, 0)]
    batch = K.MixBwdBatch("cuda")
    batch.begin_iteration()
    ref, got = [], []
    for (E, O, I, k, G, i_pad, o_total, o_off) in heads:
        W = torch.randn((E, O, I, k, k) if E > 1 else (O, I, k, k), generator=g).to(DEV)
        fcw = torch.randn(3, 1, generator=g).to(DEV) if E > 1 else None
        fcb = torch.randn(3, generator=g).to(DEV) if E > 1 else None
        types = [float(1 + (t % 4)) for t in range(G)]
        dK = torch.randn(G, o_total, k * k, i_pad, generator=g).to(DEV)
        outs = []
        for _ in range(2):
            dW = torch.full_like(W, 0.5)
            dfw = torch.full_like(fcw, 0.25) if E > 1 else None
            dfb = torch.full_like(fcb, -0.25) if E > 1 else None
            outs.append((dW, dfw, dfb))
        K.condconv_mix_bwd(dK, W, fcw, fcb, types, i_pad, o_total, o_off, *outs[0])
        batch.add(dK, W, fcw, fcb, types, i_pad, o_total, o_off, *outs[1])
        ref.append(outs[0]); got.append(outs[1])
    batch.flush()
    torch.cuda.synchronize()
    for (r, q) in zip(ref, got):
        for a, b in zip(r, q):
            if a is not None:
                _close(b, a, 1e-5, 1e-5, "mix_bwd batched vs per head")


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("shape", [(16, 4, 16, 7, 9, 32), (4, 1, 16, 7, 9, 32), (8, 2, 32, 4, 9, 16)])
def test_compose_tail_fwd_bwd(dt, shape):
    """rd_compose_tail_fwd / _bwd (W_eff = W_B W_A, b_eff = W_B b_A + b_B and their chain rule) against the einsum statement."""
    G, modules, OA, OB, taps, Cin = shape
    g = torch.Generator().manual_seed(17)
    pA = torch.randn(G, OA, taps, Cin, generator=g)
    pB = torch.randn(G, OB, OA, generator=g)
    bA, bB = torch.randn(modules, OA, generator=g), torch.randn(modules, OB, generator=g)
    o_pad = 16 if dt == torch.bfloat16 else OB
    outs = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        packed = torch.empty(G, OB, taps, Cin, dtype=dt, device=dev)
        packedT = torch.full((G, Cin, taps, o_pad), 7.0, dtype=dt, device=dev)
        b_eff = torch.empty(G, OB, device=dev)
        mod.compose_tail_fwd(pA.to(dev), pB.to(dev), bA.to(dev), bB.to(dev), modules, packed, packedT, b_eff)
        outs.append((packed, packedT, b_eff))
    rt, at = _tol(dt)
    for a, b, w in zip(outs[0], outs[1], ("packed", "packedT", "b_eff")):
        _close(a, b, rt, at, "compose fwd " + w)
    dK = torch.randn(G, o_pad, taps, Cin, generator=g)
    db = torch.randn(G, o_pad, generator=g)
    outs = []
    for dev, mod in ((DEV, K), ("cpu", emul)):
        dpA = torch.empty(G, OA, taps, Cin, device=dev)
        dpB = torch.empty(G, OB, OA, device=dev)
        dbA, dbB = torch.full((modules, OA), 0.5, device=dev), torch.full((modules, OB), -0.5, device=dev)
        mod.compose_tail_bwd(dK.to(dev), db.to(dev), pA.to(dev), pB.to(dev), bA.to(dev), modules, dpA, dpB, dbA, dbB)
        outs.append((dpA, dpB, dbA, dbB))
    for a, b, w in zip(outs[0], outs[1], ("dpA", "dpB", "dbA", "dbB")):
        _close(a, b, 2e-4, 1e-4, "compose bwd " + w)


@pytest.mark.parametrize("dt", DTS)
def test_softplus_and_avgpool16(dt):
    rt, at = _tol(dt)
    x = _rand((3, 8, 8, 5), dt, 81, 8.0)          # includes |x| > 20 (the linear branch of F.softplus)
    x[0, 0, 0, 0] = 25.0
    y_g, y_c = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.softplus_fwd(x.to(DEV), y_g)
    emul.softplus_fwd(x, y_c)
    _close(y_g, y_c, rt, at, "softplus fwd")
    dy = _rand(tuple(x.shape), dt, 82)
    dx_g, dx_c = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.softplus_bwd(dy.to(DEV), x.to(DEV), dx_g)
    emul.softplus_bwd(dy, x, dx_c)
    _close(dx_g, dx_c, rt, at, "softplus bwd")
    s = _rand((2, 32, 48, 4), dt, 83)
    p_g, p_c = torch.empty(2, 4 * 2 * 3, device=DEV), torch.empty(2, 4 * 2 * 3)
    K.avgpool16_fwd(s.to(DEV), p_g)
    emul.avgpool16_fwd(s, p_c)
    _close(p_g, p_c, 1e-5, 1e-6, "avgpool16 fwd")
    dp = torch.randn(2, 24, generator=torch.Generator().manual_seed(84))
    ds_g, ds_c = torch.empty_like(s, device=DEV), torch.empty_like(s)
    K.avgpool16_bwd(dp.to(DEV), ds_g)
    emul.avgpool16_bwd(dp, ds_c)
    _close(ds_g, ds_c, rt, at, "avgpool16 bwd")


def test_abi_graph_capture_and_replay():
    """rd_graph_begin / _end / _launch: a sequence of rd_* launches captured through the C ABI replays with the same result."""
    a = torch.arange(4096, dtype=torch.float32, device=DEV)
    b = torch.ones(4096, device=DEV)
    out = torch.zeros(4096, device=DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = K.AbiGraph()
        with g.capture():
            K.add(a, b, out)          # out = a + b
            K.add(out, b, out)        # out += b
            K.add(out, out, a)        # a = 2 out  (the next replay starts from the updated a)
        assert float(out.sum()) == 0.0, "capture must not execute"
        g.launch()
        s.synchronize()
        first = out.clone()
        g.launch()
        s.synchronize()
    base = torch.arange(4096, dtype=torch.float32)
    assert torch.equal(first.cpu(), base + 2)
    assert torch.equal(out.cpu(), 2 * (base + 2) + 2)
    kernels, total = g.node_count()
    assert kernels == 3 and total >= 3
    g.destroy()


def test_abi_ddp_two_ranks():
    """rd_ddp_* over NCCL with one process per GPU (tools/ddp_abi_check.py under torchrun); needs two GPUs."""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_kernels_gpu.py -m gpu -k abi_ddp)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29544", os.path.join(root, "tools", "ddp_abi_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok_all_ranks"]


@pytest.mark.parametrize("dt", DTS)
def test_channel_attention_and_symmetry_helpers(dt):
    """rd_chan_scale_fwd / _bwd, rd_chan_bcast, rd_flip_absdiff_fwd / _bwd and rd_mul_bcast with the residual offset (output-decoder
    variants U+SA+CA / U+SSA+CA) against their torch statements."""
    rt, at = _tol(dt)
    N, H, W, C = 3, 10, 12, 40
    x, dy = _rand((N, H, W, C), dt, 201), _rand((N, H, W, C), dt, 202)
    a = torch.rand(N, C, generator=torch.Generator().manual_seed(203))
    y_g, y_c = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.chan_scale_fwd(x.to(DEV), a.to(DEV), y_g)
    emul.chan_scale_fwd(x, a, y_c)
    _close(y_g, y_c, rt, at, "chan_scale fwd")
    dx_g, da_g = torch.empty_like(x, device=DEV), torch.empty(N, C, device=DEV)
    dx_c, da_c = torch.empty_like(x), torch.empty(N, C)
    K.chan_scale_bwd(x.to(DEV), a.to(DEV), dy.to(DEV), dx_g, da_g)
    emul.chan_scale_bwd(x, a, dy, dx_c, da_c)
    _close(dx_g, dx_c, rt, at, "chan_scale dx")
    _close(da_g, da_c, 2e-4, 1e-4, "chan_scale dalpha")
    b_g, b_c = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.chan_bcast(a.to(DEV), b_g, 1.0 / (H * W))
    emul.chan_bcast(a, b_c, 1.0 / (H * W))
    _close(b_g, b_c, rt, 1e-7, "chan_bcast")
    g = _rand((N, H, W, C), dt, 204)
    g[0, 2] = g[0, H - 3]                      # exact ties: sign(0) = 0 in the backward
    o_g, o_c = torch.empty_like(g, device=DEV), torch.empty_like(g)
    K.flip_absdiff_fwd(g.to(DEV), o_g)
    emul.flip_absdiff_fwd(g, o_c)
    _close(o_g, o_c, 0, 0, "flip_absdiff fwd")
    dg_g, dg_c = torch.empty_like(g, device=DEV), torch.empty_like(g)
    K.flip_absdiff_bwd(g.to(DEV), dy.to(DEV), dg_g)
    emul.flip_absdiff_bwd(g, dy, dg_c)
    _close(dg_g, dg_c, rt, at, "flip_absdiff bwd")
    al = torch.rand(N, H, W, 1, generator=torch.Generator().manual_seed(205)).to(dt)
    m_g, m_c = torch.empty_like(x, device=DEV), torch.empty_like(x)
    K.mul_bcast_fwd(al.to(DEV), x.to(DEV), m_g, 1.0)
    emul.mul_bcast_fwd(al, x, m_c, 1.0)
    _close(m_g, m_c, rt, at, "mul_bcast(1 + alpha) fwd")
    dxm_g, dal_g = torch.empty_like(x, device=DEV), torch.empty_like(al, device=DEV)
    dxm_c, dal_c = torch.empty_like(x), torch.empty_like(al)
    K.mul_bcast_bwd(al.to(DEV), x.to(DEV), dy.to(DEV), dxm_g, dal_g, 1.0)
    emul.mul_bcast_bwd(al, x, dy, dxm_c, dal_c, 1.0)
    _close(dxm_g, dxm_c, rt, at, "mul_bcast(1 + alpha) dx")
    _close(dal_g, dal_c, 2e-2 if dt == torch.bfloat16 else 2e-4, 2e-2 if dt == torch.bfloat16 else 1e-4, "mul_bcast dalpha")


def test_modality_weights_bit_exact():
    """rd_modality_weights: [mask[:, i].sum() != 0] / #present contrasts."""
    for rows in ([[1, 1, 0, 1], [1, 0, 0, 1]], [[0, 0, 0, 0], [0, 0, 0, 0]], [[1, 1], [1, 1]], [[0, 1, 0, 0]]):
        m = torch.tensor(rows, dtype=torch.float32)
        w_g, w_c = torch.empty(m.shape[1], device=DEV), torch.empty(m.shape[1])
        K.modality_weights(m.to(DEV), w_g)
        emul.modality_weights(m, w_c)
        assert torch.equal(w_g.cpu(), w_c), (rows, w_g, w_c)


@pytest.mark.parametrize("dt", DTS)
def test_stack_modalities_bit_exact(dt):
    src = _rand((3, 28, 16, 24), torch.float32, 301)
    for cp in (7, 16):
        d_g, d_c = torch.full((12, 16, 24, cp), 5.0, dtype=dt, device=DEV), torch.empty(12, 16, 24, cp, dtype=dt)
        K.stack_modalities(src.to(DEV), d_g, 4)
        emul.stack_modalities(src, d_c, 4)
        _close(d_g, d_c, 0, 0, "stack_modalities")


@pytest.mark.parametrize("dt", DTS)
def test_add_n(dt):
    xs = [_rand((3, 5, 7, 9), dt, 400 + k) for k in range(5)]
    y_g, y_c = torch.empty_like(xs[0], device=DEV), torch.empty_like(xs[0])
    K.add_n([x.to(DEV) for x in xs], y_g)
    emul.add_n(xs, y_c)
    rt, at = _tol(dt)
    _close(y_g, y_c, rt, at, "add_n")


@pytest.mark.parametrize("dt", DTS)
def test_scatter_blocks2(dt):
    a, b = _rand((4 * 2, 5, 6, 7), dt, 500), _rand((12 * 2, 5, 6, 7), dt, 501)
    sel = [0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0]
    sblk, ka, kb = [], 0, 0
    for s_ in sel:
        if s_:
            sblk.append(kb); kb += 1
        else:
            sblk.append(ka); ka += 1
    cp = 16 if dt == torch.bfloat16 else 7
    d_g, d_c = torch.full((32, 5, 6, cp), 3.0, dtype=dt, device=DEV), torch.empty(32, 5, 6, cp, dtype=dt)
    K.scatter_blocks2(a.to(DEV), b.to(DEV), d_g, sel, sblk, 2)
    emul.scatter_blocks2(a, b, d_c, sel, sblk, 2)
    _close(d_g, d_c, 0, 0, "scatter_blocks2")
