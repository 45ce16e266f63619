"""TEST INFRASTRUCTURE — plain-PyTorch statement of every kernel contract in rd_b200/kernels.py.

Same function names and argument lists as rd_b200.kernels, but computed with torch ops in fp32 on
whatever device the tensors live on.  Two uses, both in tests only:
  * -m gpu   : each CUDA kernel is compared against the function of the same name here;
  * not gpu  : `install()` monkeypatches rd_b200.kernels with these functions so that the HOST logic
               (autograd wiring in ops.py, batching / indexing in model.py, the loop body in trainer.py)
               can be checked against the oracle on CPU.  The product never imports this file.
"""
import math

import torch
import torch.nn.functional as F

from rd_b200.lib import ConvDesc, RD_BF16, RD_F32


def _dt(t):
    return RD_F32 if t.dtype == torch.float32 else RD_BF16


def _f(t):
    return t.float()


def _wr(dst, val):
    dst.copy_(val.reshape(dst.shape).to(dst.dtype))


# ------------------------------------------------------------------------------- layout / cast
def nchw_to_nhwc(src, dst, c0, c):
    _wr(dst, src[:, c0:c0 + c].permute(0, 2, 3, 1))


def stack_modalities(src, dst, mods):
    n, ct, h, w = src.shape
    c = ct // mods
    dst.zero_()
    for m in range(mods):
        dst[m * n:(m + 1) * n, ..., :c] = src[:, m * c:(m + 1) * c].permute(0, 2, 3, 1).to(dst.dtype)


def nchw_to_nhwc_strided(src, dst, c_total):
    _wr(dst, src.permute(0, 2, 3, 1))


def nhwc_to_nchw(src, dst):
    _wr(dst, _f(src).permute(0, 3, 1, 2))


def cast(src, dst):
    _wr(dst, src)


def concat_channels(a, b, out):
    _wr(out, torch.cat([a, b], -1))


def split_channels(inp, a, b, ca, cb):
    if a is not None:
        _wr(a, inp[..., :ca])
    if b is not None:
        _wr(b, inp[..., ca:ca + cb])


def add(x, a, y):
    _wr(y, _f(x) + _f(a))


# ------------------------------------------------------------------------------- CondConv mixing
def zeros(shape, dtype, device):
    return torch.zeros(shape, dtype=dtype, device=device)


def add_n(xs, y):
    acc = _f(xs[0]).clone()
    for t in xs[1:]:
        acc = acc + _f(t)
    _wr(y, acc)


def _route(fc_w, fc_b, types, E, dev):
    t = torch.tensor(list(types), dtype=torch.float32, device=dev).reshape(-1, 1)
    if fc_w is None:
        return torch.ones(len(types), E, device=dev)
    return torch.sigmoid(t * fc_w.reshape(1, -1) + fc_b.reshape(1, -1))


def condconv_mix_fwd(W, fc_w, fc_b, types, i_pad, o_total, oT_total, o_off, packed, packedT, r_out):
    W5 = W if W.dim() == 5 else W.unsqueeze(0)
    E, O, I_, kh, kw = W5.shape
    r = _route(fc_w, fc_b, types, E, W.device)
    Kmix = torch.einsum("ge,eoihw->goihw", r, W5)                  # (G,O,I,kh,kw)
    ohwi = torch.zeros(len(types), O, kh * kw, i_pad, device=W.device)
    ohwi[..., :I_] = Kmix.permute(0, 1, 3, 4, 2).reshape(len(types), O, kh * kw, I_)
    if packed is not None:
        packed[:, o_off:o_off + O] = ohwi.to(packed.dtype)
    if packedT is not None:
        packedT[:, :, :, o_off:o_off + O] = ohwi.permute(0, 3, 2, 1).to(packedT.dtype)
    if r_out is not None:
        _wr(r_out, r)


def condconv_mix_bwd(dK, W, fc_w, fc_b, types, i_pad, o_total, o_off, dW, dfc_w, dfc_b):
    W5 = W if W.dim() == 5 else W.unsqueeze(0)
    E, O, I_, kh, kw = W5.shape
    G = len(types)
    r = _route(fc_w, fc_b, types, E, W.device)
    d = dK[:, o_off:o_off + O, :, :I_].reshape(G, O, kh, kw, I_).permute(0, 1, 4, 2, 3)     # (G,O,I,kh,kw)
    dW += torch.einsum("ge,goihw->eoihw", r, d).reshape(dW.shape)
    if fc_w is not None and dfc_w is not None:
        dr = torch.einsum("goihw,eoihw->ge", d, W5)
        t = torch.tensor(list(types), dtype=torch.float32, device=W.device).reshape(-1, 1)
        s = dr * r * (1 - r)
        dfc_w += (s * t).sum(0).reshape(dfc_w.shape)
        dfc_b += s.sum(0).reshape(dfc_b.shape)


def modality_weights(mask, w):
    present = (mask.sum(0) != 0).float()
    n = present.sum()
    w.copy_(present / n if n > 0 else present)


def compose_tail_fwd(pA, pB, bA, bB, modules, packed, packedT, b_eff):
    """rd_compose_tail_fwd: W_eff[g] = pB[g] . pA[g], b_eff[g] = pB[g] . bA[m] + bB[m]."""
    G, OB, OA = pB.shape
    Gm = G // modules
    w_eff = torch.einsum("goc,gcti->goti", pB, pA)
    mod_of = torch.arange(G, device=pA.device) // Gm
    be = torch.zeros(G, OB, device=pA.device)
    if bA is not None:
        be = be + torch.einsum("goc,gc->go", pB, bA[mod_of])
    if bB is not None:
        be = be + bB[mod_of]
    packed.copy_(w_eff.to(packed.dtype))
    packedT.zero_()
    packedT[..., :OB] = w_eff.permute(0, 3, 2, 1).to(packedT.dtype)
    b_eff.copy_(be)


def compose_tail_bwd(dK, db, pA, pB, bA, modules, dpA, dpB, dbA, dbB):
    G, OB, OA = pB.shape
    Gm = G // modules
    dKe, dbe = dK[:, :OB], db[:, :OB]
    mod_of = torch.arange(G, device=pA.device) // Gm
    dpA.copy_(torch.einsum("goc,goti->gcti", pB, dKe))
    t = torch.einsum("goti,gcti->goc", dKe, pA)
    if bA is not None:
        t = t + dbe[:, :, None] * bA[mod_of][:, None, :]
    dpB.copy_(t)
    if dbA is not None:
        dbA += torch.einsum("goc,go->gc", pB, dbe).reshape(modules, Gm, OA).sum(1)
    if dbB is not None:
        dbB += dbe.reshape(modules, Gm, OB).sum(1)


def pad_channels(inp, out):
    out.zero_()
    out[..., :inp.shape[-1]] = inp


def gather_blocks_fwd(src, dst, index, block):
    c = src.shape[-1]
    dst.zero_()
    for k, i in enumerate(index):
        dst[k * block:(k + 1) * block, ..., :c] = src[i * block:(i + 1) * block]


def scatter_blocks2(a, b, dst, sel, sblk, block):
    c = a.shape[-1]
    dst.zero_()
    for d, (s_, k) in enumerate(zip(sel, sblk)):
        src = b if s_ else a
        dst[d * block:(d + 1) * block, ..., :c] = src[k * block:(k + 1) * block]


def gather_blocks_bwd(dout, dsrc, index, block):
    c = dsrc.shape[-1]
    acc = torch.zeros(dsrc.shape, dtype=torch.float32)
    for k, i in enumerate(index):
        acc[i * block:(i + 1) * block] += dout[k * block:(k + 1) * block, ..., :c].float()
    dsrc.copy_(acc.to(dsrc.dtype))


# ------------------------------------------------------------------------------- convolution
def conv_desc(n, h, w, cin, cout, kh, kw, stride, pad, groups, dtype, act=0, slope=0.2, algo=0, bias_groups=0) -> ConvDesc:
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    return ConvDesc(n, h, w, cin, oh, ow, cout, kh, kw, stride, pad, groups, dtype, act, slope, algo, bias_groups)


def _w_oihw(packed, d, g):
    return _f(packed[g]).reshape(d.cout, d.kh, d.kw, d.cin).permute(0, 3, 1, 2)


def conv2d_fwd(d, x, packed, bias, y):
    ipg = d.n // d.groups
    for g in range(d.groups):
        xi = _f(x[g * ipg:(g + 1) * ipg]).permute(0, 3, 1, 2)
        if bias is None:
            bg = None
        elif d.bias_groups <= 1:
            bg = bias.reshape(-1)
        else:
            bg = bias.reshape(d.bias_groups, -1)[g // (d.groups // d.bias_groups)]
        o = F.conv2d(xi, _w_oihw(packed, d, g), bg, d.stride, d.pad)
        if d.act == 1:
            o = F.leaky_relu(o, d.act_slope)
        y[g * ipg:(g + 1) * ipg] = o.permute(0, 2, 3, 1).to(y.dtype)


def conv2d_dgrad(d, dy, packedT, dx):
    ipg = d.n // d.groups
    for g in range(d.groups):
        w = _f(packedT[g]).reshape(d.cin, d.kh, d.kw, d.cout).permute(3, 0, 1, 2)      # (O,I,kh,kw)
        g_in = torch.nn.grad.conv2d_input((ipg, d.cin, d.h, d.w), w, _f(dy[g * ipg:(g + 1) * ipg]).permute(0, 3, 1, 2),
                                          d.stride, d.pad)
        dx[g * ipg:(g + 1) * ipg] = g_in.permute(0, 2, 3, 1).to(dx.dtype)


def conv2d_wgrad(d, x, dy, dK, dbias):
    ipg = d.n // d.groups
    for g in range(d.groups):
        gw = torch.nn.grad.conv2d_weight(_f(x[g * ipg:(g + 1) * ipg]).permute(0, 3, 1, 2), (d.cout, d.cin, d.kh, d.kw),
                                         _f(dy[g * ipg:(g + 1) * ipg]).permute(0, 3, 1, 2), d.stride, d.pad)
        dK[g] = gw.permute(0, 2, 3, 1).reshape(d.cout, d.kh * d.kw, d.cin)
    if dbias is not None:
        if d.bias_groups > 1:
            dbias += _f(dy).reshape(d.bias_groups, -1, d.cout).sum(1).reshape(dbias.shape)
        else:
            dbias += _f(dy).sum((0, 1, 2))


# ------------------------------------------------------------------------------- normalisation
def norm_workspace(G, ppg, Cn, device):
    return torch.empty(1, device=device)


def norm_stats(x, G, ppg, Cn, eps, partial, mean, invstd, running_mean=None, running_var=None, nbt=None, momentum=0.1):
    xg = _f(x).reshape(G, ppg, Cn)
    mu = xg.mean(1)
    var = xg.var(1, unbiased=False)
    _wr(mean, mu)
    _wr(invstd, torch.rsqrt(var + eps))
    if running_mean is not None:
        for g in range(G):
            unb = var[g] * ppg / (ppg - 1) if ppg > 1 else var[g]
            running_mean.mul_(1 - momentum).add_(momentum * mu[g])
            running_var.mul_(1 - momentum).add_(momentum * unb)
        if nbt is not None:
            nbt += G


def norm_eval_stats(running_mean, running_var, G, eps, mean, invstd):
    _wr(mean, running_mean.repeat(G))
    _wr(invstd, torch.rsqrt(running_var + eps).repeat(G))


def norm_apply(x, mean, invstd, weight, bias, y, G, ppg, Cn):
    xg = _f(x).reshape(G, ppg, Cn)
    o = (xg - mean.reshape(G, 1, Cn)) * invstd.reshape(G, 1, Cn)
    if weight is not None:
        o = o * weight + bias
    _wr(y, o)


def norm_bwd(x, dy, mean, invstd, weight, dx, dweight, dbias, partial, G, ppg, Cn):
    xg, dg = _f(x).reshape(G, ppg, Cn), _f(dy).reshape(G, ppg, Cn)
    xh = (xg - mean.reshape(G, 1, Cn)) * invstd.reshape(G, 1, Cn)
    s1, s2 = dg.sum(1, keepdim=True), (dg * xh).sum(1, keepdim=True)
    w = weight if weight is not None else torch.ones(Cn, device=x.device)
    _wr(dx, w * invstd.reshape(G, 1, Cn) * (dg - s1 / ppg - xh * s2 / ppg))
    if dweight is not None:
        dweight += s2.sum((0, 1))
    if dbias is not None:
        dbias += s1.sum((0, 1))


def spade_modulate_fwd(z, mean, invstd, gb, mix):
    N, H, W, Cn = z.shape
    zh = (_f(z) - mean.reshape(N, 1, 1, Cn)) * invstd.reshape(N, 1, 1, Cn)
    g = _f(gb)
    _wr(mix, zh * (1 + g[..., :Cn]) + g[..., Cn:])


def spade_modulate_bwd(z, mean, invstd, gb, dmix, dz, dgb, partial):
    N, H, W, Cn = z.shape
    is_ = invstd.reshape(N, 1, 1, Cn)
    zh = (_f(z) - mean.reshape(N, 1, 1, Cn)) * is_
    dm = _f(dmix)
    _wr(dgb, torch.cat([dm * zh, dm], -1))
    dxh = dm * (1 + _f(gb)[..., :Cn])
    n = H * W
    s1, s2 = dxh.sum((1, 2), keepdim=True), (dxh * zh).sum((1, 2), keepdim=True)
    _wr(dz, is_ * (dxh - s1 / n - zh * s2 / n))


def spade_bwd_workspace(z):
    return torch.empty(1)


def spade_modulate_bwd_g(z, mean, invstd, gamma, dmix, dz, dgb, partial):
    spade_modulate_bwd(z, mean, invstd, torch.cat([_f(gamma), torch.zeros_like(_f(gamma))], -1), dmix, dz, dgb, partial)


SPADE_FUSE = False          # CPU host-logic tests set this to route the SPADE blocks through the fused gamma|beta convolution


def conv2d_fwd_spade_supported(d, x):
    return bool(SPADE_FUSE) and d.cout % 2 == 0


def conv2d_fwd_spade(d, x, packed, bias, z, mean, invstd, gamma, mix):
    """rd_conv2d_fwd_spade: gb = conv(x) + bias (fp32 accumulators, NOT rounded to the storage dtype), gamma = gb[..., :C],
    mix = (z - mean) * invstd * (1 + gamma) + beta."""
    gb = torch.empty(x.shape[:-1] + (d.cout,), dtype=torch.float32, device=x.device)
    conv2d_fwd(d, x, packed, bias, gb)
    N, H, W, Cn = z.shape
    zh = (_f(z) - mean.reshape(N, 1, 1, Cn)) * invstd.reshape(N, 1, 1, Cn)
    _wr(gamma, gb[..., :Cn])
    _wr(mix, zh * (1 + gb[..., :Cn]) + gb[..., Cn:])


# ------------------------------------------------------------------------------- resize / activations
def bilinear_fwd(x, y, align):
    o = F.interpolate(_f(x).permute(0, 3, 1, 2), size=(y.shape[1], y.shape[2]), mode="bilinear", align_corners=bool(align))
    _wr(y, o.permute(0, 2, 3, 1))


def bilinear_bwd(dy, dx, align):
    with torch.enable_grad():
        xin = torch.zeros(dx.shape, dtype=torch.float32, device=dx.device).permute(0, 3, 1, 2).requires_grad_(True)
        o = F.interpolate(xin, size=(dy.shape[1], dy.shape[2]), mode="bilinear", align_corners=bool(align))
        (g,) = torch.autograd.grad(o, xin, _f(dy).permute(0, 3, 1, 2))
    _wr(dx, g.permute(0, 2, 3, 1))


def lrelu_fwd(x, y, slope):
    _wr(y, F.leaky_relu(_f(x), slope))


def lrelu_bwd(dy, y, dx, slope):
    _wr(dx, torch.where(_f(y) > 0, _f(dy), _f(dy) * slope))


def masked_softmax_fwd(s, mask_img, p):
    sf = _f(s)
    if mask_img is None:
        _wr(p, torch.softmax(sf, -1))
        return
    rep = sf.shape[0] // mask_img.shape[0]
    m = mask_img.repeat(rep, 1, 1).unsqueeze(-1)
    _wr(p, torch.softmax(torch.cat([100 * m, sf], -1), -1)[..., 1:])


def masked_softmax_bwd(p, dp, ds):
    pf, df = _f(p), _f(dp)
    _wr(ds, pf * (df - (pf * df).sum(-1, keepdim=True)))


def add_relu_fwd(a, b, y):
    _wr(y, torch.relu(_f(a) + _f(b)))


def relu_bwd(dy, y, dx):
    _wr(dx, torch.where(_f(y) > 0, _f(dy), torch.zeros_like(_f(dy))))


def sigmoid_fwd(x, y):
    _wr(y, torch.sigmoid(_f(x)))


def sigmoid_bwd(dy, y, dx):
    _wr(dx, _f(dy) * _f(y) * (1 - _f(y)))


def mul_bcast_fwd(alpha, x, y, off=0.0):
    _wr(y, (off + _f(alpha)) * _f(x))


def mul_bcast_bwd(alpha, x, dy, dx, dalpha, off=0.0):
    _wr(dx, (off + _f(alpha)) * _f(dy))
    _wr(dalpha, (_f(dy) * _f(x)).sum(-1, keepdim=True))


def chan_scale_fwd(x, a, y):
    N, H, W, Cn = x.shape
    _wr(y, (1 + a.reshape(N, 1, 1, Cn)) * _f(x))


def chan_scale_bwd(x, a, dy, dx, da):
    N, H, W, Cn = x.shape
    _wr(dx, (1 + a.reshape(N, 1, 1, Cn)) * _f(dy))
    _wr(da, (_f(dy) * _f(x)).sum((1, 2)).reshape(da.shape))


def chan_bcast(v, dx, scale):
    N, H, W, Cn = dx.shape
    _wr(dx, (v.reshape(N, 1, 1, Cn) * scale).expand(N, H, W, Cn))


def flip_absdiff_fwd(g, out):
    _wr(out, (_f(g) - torch.flip(_f(g), dims=[1])).abs())


def flip_absdiff_bwd(g, dout, dg):
    d = _f(g) - torch.flip(_f(g), dims=[1])
    _wr(dg, torch.sign(d) * (_f(dout) + torch.flip(_f(dout), dims=[1])))


# ------------------------------------------------------------------------------- small dense
def linear_fwd(x, W, b, y, act=0, slope=0.2):
    o = F.linear(x, W, b)
    if act == 1:
        o = F.leaky_relu(o, slope)
    _wr(y, o)


def linear_bwd(x, W, dy, dx, dW, db):
    if dx is not None:
        _wr(dx, dy @ W)
    if dW is not None:
        dW += dy.t() @ x
    if db is not None:
        db += dy.sum(0)


def sample_fwd(mu, lv, eps, z):
    _wr(z, mu + eps * torch.exp(0.5 * lv))


def sample_bwd(dz, lv, eps, dmu, dlv):
    _wr(dmu, dz)
    _wr(dlv, dz * eps * 0.5 * torch.exp(0.5 * lv))


# ------------------------------------------------------------------------------- fusion gather
def fuse_gather_fwd(si, mask, out, idx_out, count_out, B, M):
    rows = si.reshape(M, B, -1)
    flags = (mask.reshape(B, M) == 1)
    k = 0
    o = out.reshape(B * M, -1)
    if idx_out is not None:
        idx_out.fill_(-1)
    for b in range(B):
        for m in range(M):
            if flags[b, m]:
                o[k] = rows[m, b]
                if idx_out is not None:
                    idx_out[k] = b * M + m
                k += 1
    if count_out is not None:
        count_out.fill_(k)


def fuse_gather_bwd(dout, mask, dsi, B, M):
    d = dsi.reshape(M, B, -1)
    d.zero_()
    flags = (mask.reshape(B, M) == 1)
    src = dout.reshape(B * M, -1)
    k = 0
    for b in range(B):
        for m in range(M):
            if flags[b, m]:
                d[m, b] = src[k]
                k += 1


# ------------------------------------------------------------------------------- losses
def recon_chunks(row_elems):
    return 1


def recon_rows_fwd(x, gt, gt_index, row_loss, partial, R, p):
    xr, gr = _f(x).reshape(R, -1), _f(gt).reshape(-1, x.numel() // R)
    for r in range(R):
        gi = int(gt_index[r]) if gt_index is not None else r
        if gi < 0:
            row_loss[r] = 0
            continue
        d = gr[gi] - xr[r]
        row_loss[r] = d.abs().mean() if p == 1 else (d * d).mean()


def recon_rows_bwd(x, gt, gt_index, coef, dx, R, p):
    xr, gr = _f(x).reshape(R, -1), _f(gt).reshape(-1, x.numel() // R)
    out = torch.zeros_like(xr)
    n = xr.shape[1]
    for r in range(R):
        gi = int(gt_index[r]) if gt_index is not None else r
        if gi < 0:
            continue
        d = xr[r] - gr[gi]
        out[r] = (torch.sign(d) if p == 1 else 2 * d) * coef[r] / n
    _wr(dx, out)


def xmix_plan(mask, gt_index, B, M):
    t = 0
    gt_index.fill_(-1)
    for i in range(M):
        for j in range(M):
            if i == j:
                continue
            if float((mask[:, i] * mask[:, j]).sum()) == 0:
                continue
            for b in range(B):
                gt_index[t * B + b] = j * B + b
            t += 1


def masked_combine(row_loss, mask, loss, coef, B, M, kind):
    coef.zero_()
    total, cnt = 0.0, 0
    if kind == 0:
        for i in range(M):
            ms = float(mask[:, i].sum())
            if ms == 0:
                continue
            cnt += 1
            total += float((mask[:, i] * row_loss[i * B:(i + 1) * B]).sum()) / ms
            coef[i * B:(i + 1) * B] = mask[:, i] / ms
    else:
        for i in range(M):
            for j in range(M):
                if i == j:
                    continue
                mm = mask[:, i] * mask[:, j]
                ms = float(mm.sum())
                if ms == 0:
                    continue
                total += float((mm * row_loss[cnt * B:(cnt + 1) * B]).sum()) / ms
                coef[cnt * B:(cnt + 1) * B] = mm / ms
                cnt += 1
    if cnt > 0:
        total /= cnt
        coef /= cnt
    loss.fill_(total)


def _with_grad(fn, *ins):
    with torch.enable_grad():
        xs = [t.detach().clone().requires_grad_(True) for t in ins]
        l = fn(*xs)
        if l.requires_grad:
            gs = torch.autograd.grad(l, xs, allow_unused=True)
        else:
            gs = [None] * len(xs)
    return l.detach(), [g if g is not None else torch.zeros_like(x) for g, x in zip(gs, xs)]


def _cos(x, y):
    xn = torch.sqrt((x ** 2).sum(1) + 1e-8).clamp_min(1e-8)
    yn = torch.sqrt((y ** 2).sum(1) + 1e-8).clamp_min(1e-8)
    return (x * y).sum(1) / (xn * yn)


def latent_z_loss(mu, mu_new, mask, loss, dmu, dmu_new, B, M, Z):
    def fn(a, b):
        tot, cnt = torch.zeros(()), 0
        a, b = a.reshape(M, B, Z), b.reshape(M, B, Z)
        for i in range(M):
            if mask[:, i].sum() == 0:
                continue
            cnt += 1
            tot = tot + (mask[:, i].unsqueeze(1) * (a[i] - b[i]).abs()).sum() / mask[:, i].sum()
        return tot / cnt if cnt else tot + 0 * a.sum()
    l, (ga, gb) = _with_grad(fn, mu, mu_new)
    loss.fill_(float(l)), _wr(dmu, ga), _wr(dmu_new, gb)


def sim_z_loss(z, mask, margin, loss, dz, B, M, Z):
    def fn(a):
        a = a.reshape(M, B, Z)
        tot, cnt = torch.zeros(()), 0
        for i in range(M - 1):
            zp = torch.roll(a[i], -1, 0)
            mp = torch.roll(mask[:, i], -1, 0)
            for j in range(i + 1, M):
                mm = mask[:, i] * mask[:, j] * mp
                if mm.sum() == 0:
                    continue
                cnt += 1
                tot = tot + (mm * torch.clamp_min(margin - _cos(a[i], zp) + _cos(a[i], a[j]), 0)).sum() / mm.sum()
        return tot / cnt if cnt else tot + 0 * a.sum()
    l, (g,) = _with_grad(fn, z)
    loss.fill_(float(l)), _wr(dz, g)


def kl_loss(mu, lv, mask, loss, dmu, dlv, B, M, Z):
    def fn(a, b):
        m = mask.t().reshape(-1)
        kl = 0.5 * torch.sum(torch.exp(b) + a ** 2 - 1. - b, 1)
        return (kl * m).sum() / m.sum() / M
    l, (ga, gb) = _with_grad(fn, mu, lv)
    loss.fill_(float(l)), _wr(dmu, ga), _wr(dlv, gb)


def avgpool16_fwd(s, pooled):
    x = _f(s).permute(0, 3, 1, 2)
    _wr(pooled, torch.nn.functional.avg_pool2d(x, 16).reshape(x.shape[0], -1))


def avgpool16_bwd(dpooled, ds):
    n, h, w, c = ds.shape
    g = dpooled.float().reshape(n, c, h // 16, w // 16)
    g = g.repeat_interleave(16, 2).repeat_interleave(16, 3) / 256.0
    _wr(ds, g.permute(0, 2, 3, 1))


def softplus_fwd(x, y):
    _wr(y, torch.nn.functional.softplus(_f(x)))


def softplus_bwd(dy, x, dx):
    _wr(dx, _f(dy) * torch.sigmoid(_f(x)))


def maxpool16_fwd(s, pooled, argmax):
    N, H, W, Cn = s.shape
    o, idx = F.max_pool2d(_f(s).permute(0, 3, 1, 2), (16, 16), return_indices=True)
    _wr(pooled, o.reshape(N, -1))
    _wr(argmax, idx.reshape(N, -1).to(torch.int32))


def maxpool16_bwd(dpooled, argmax, ds):
    N, H, W, Cn = ds.shape
    flat = torch.zeros(N, Cn, H * W, device=ds.device)
    PHW = (H // 16) * (W // 16)
    flat.scatter_(2, argmax.reshape(N, Cn, PHW).long(), dpooled.reshape(N, Cn, PHW))
    _wr(ds, flat.reshape(N, Cn, H, W).permute(0, 2, 3, 1))


def sim_s_loss(pooled, mask, pair, margin, loss, dpooled, B, M, D):
    i, j = int(pair[0]), int(pair[1])

    def fn(a):
        a = a.reshape(M, B, D)
        si, sj = a[i], a[j]
        sp = torch.roll(si, -1, 0)
        mm = mask[:, i] * mask[:, j] * torch.roll(mask[:, i], -1, 0)
        if mm.sum() > 0:
            return (mm * torch.clamp_min(margin - _cos(si, sj) + _cos(sp, si), 0)).sum() / mm.sum()
        return 0 * a.sum()
    l, (g,) = _with_grad(fn, pooled)
    loss.fill_(float(l)), _wr(dpooled, g)


SEG_PARTIAL_FLOATS = 256 * 11 + 11


def _seg(yv, target):
    y = yv.permute(0, 3, 1, 2)
    gt = target.reshape(y.shape[0], 1, y.shape[2], y.shape[3])
    ce = F.cross_entropy(y, gt.squeeze(1).long(), weight=torch.tensor([1., 5., 5., 5.], device=y.device))
    act = F.softmax(y, dim=1)
    dice = 0
    for c in range(1, 4):
        g = (gt[:, 0] == c).float()
        dice = dice + 1 - 2 * (act[:, c] * g).sum() / ((act[:, c] ** 2 + g ** 2).sum() + 1e-6)
    return ce + dice / 3


def seg_loss_fwd(y, target, loss, partial):
    loss.fill_(float(_seg(_f(y), target)))


def seg_loss_bwd(y, target, partial, upstream, dy):
    with torch.enable_grad():
        yy = _f(y).detach().clone().requires_grad_(True)
        (g,) = torch.autograd.grad(_seg(yy, target), yy)
    _wr(dy, g * upstream.reshape(()))


# ------------------------------------------------------------------------------- optimizer
def _segs(segments, nseg):
    return [(int(segments[k, 0]), int(segments[k, 1])) for k in range(nseg)]


def grad_norm(grad, segments, nseg, partial, scalars, max_norm):
    tot = 0.0
    for o, l in _segs(segments, nseg):
        tot += float((grad[o:o + l].double() ** 2).sum())
    total = math.sqrt(tot)
    scalars[0] = total
    scalars[1] = min(1.0, max_norm / (total + 1e-6))
    scalars[2] = 1.0 if math.isfinite(total) else 0.0


def grad_scale(grad, segments, nseg, scalars):
    c = float(scalars[1])
    if c >= 1.0:
        return
    for o, l in _segs(segments, nseg):
        grad[o:o + l] *= c


def _one_minus_betas(hyper):
    """hyper[6], hyper[7] = (1 - beta1), (1 - beta2) rounded from double (0 = derive from the fp32 betas), see rd_loss.cu adam_betas."""
    b1, b2 = float(hyper[1]), float(hyper[2])
    o1 = float(hyper[6]) if float(hyper[6]) != 0.0 else float(torch.tensor(1.0) - hyper[1])
    o2 = float(hyper[7]) if float(hyper[7]) != 0.0 else float(torch.tensor(1.0) - hyper[2])
    return o1, o2


def adam_amsgrad(param, grad, m, v, vmax, segments, nseg, hyper):
    lr, b1, b2, eps, wd, step = [float(x) for x in hyper[:6]]
    omb1, omb2 = _one_minus_betas(hyper)
    step += 1
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for o, l in _segs(segments, nseg):
        p = param[o:o + l]
        g = grad[o:o + l] + wd * p
        m[o:o + l].mul_(b1).add_(g, alpha=omb1)
        v[o:o + l].mul_(b2).addcmul_(g, g, value=omb2)
        torch.maximum(vmax[o:o + l], v[o:o + l], out=vmax[o:o + l])
        denom = vmax[o:o + l].sqrt() / math.sqrt(bc2) + eps
        p.addcdiv_(m[o:o + l], denom, value=-lr / bc1)
    hyper[5] += 1


def clip_adam_amsgrad(param, grad, m, v, vmax, segments, nseg, hyper, scalars, zero_grad=True):
    if scalars is not None:
        grad_scale(grad, segments, nseg, scalars)
    adam_amsgrad(param, grad, m, v, vmax, segments, nseg, hyper)
    if zero_grad:
        for o, l in _segs(segments, nseg):
            grad[o:o + l] = 0


def clip_adam_amsgrad_gated(param, grad, m, v, vmax, segments, seg_param, nseg, partial, param_flags, param_steps, hyper, scalars,
                            zero_grad=True):
    """torch.optim.Adam semantics: a parameter whose gradient is exactly zero everywhere (grad None in the reference) is skipped;
    every parameter has its own step counter."""
    segs = _segs(segments, nseg)
    param_flags.zero_()
    for k, (o, l) in enumerate(segs):
        if bool((grad[o:o + l] != 0).any()):
            param_flags[int(seg_param[k])] = 1
    if scalars is not None:
        grad_scale(grad, segments, nseg, scalars)
    lr, b1, b2, eps, wd = [float(x) for x in hyper[:5]]
    omb1, omb2 = _one_minus_betas(hyper)
    for k, (o, l) in enumerate(segs):
        pid = int(seg_param[k])
        if not int(param_flags[pid]):
            continue
        step = float(param_steps[pid]) + 1
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        p = param[o:o + l]
        g = grad[o:o + l] + wd * p
        m[o:o + l].mul_(b1).add_(g, alpha=omb1)
        v[o:o + l].mul_(b2).addcmul_(g, g, value=omb2)
        torch.maximum(vmax[o:o + l], v[o:o + l], out=vmax[o:o + l])
        denom = vmax[o:o + l].sqrt() / math.sqrt(bc2) + eps
        p.addcdiv_(m[o:o + l], denom, value=-lr / bc1)
        if zero_grad:
            grad[o:o + l] = 0
    param_steps += param_flags.to(param_steps.dtype)
    hyper[5] += 1


def metrics_recon(target, pred, out, t_index=None, t_c0=0, p_c0=0):
    from oracle.metrics_oracle import compute_reconstruction_metrics_single
    n = pred.shape[0]
    for k in range(n):
        tn = int(t_index[k]) if t_index is not None else k
        m = compute_reconstruction_metrics_single(target[tn, :, :, t_c0].float().numpy(), pred[k, :, :, p_c0].float().numpy())
        out[k, 0], out[k, 1], out[k, 2] = m["ssim"], m["psnr"], m["rmse"]


def metrics_seg(target, pred, out):
    from oracle.metrics_oracle import compute_segmentation_metrics_single
    n = pred.shape[0]
    h, w = pred.shape[1], pred.shape[2]
    for k in range(n):
        m = compute_segmentation_metrics_single(target[k].reshape(h, w).numpy(), pred[k].float().permute(2, 0, 1).numpy())
        out[k, 0], out[k, 1] = float(m["dice"]), float(m["iou"])


def assemble_slabs(vols, present, tvols, has_target, brain_mask, subj, slice_idx, drop, inputs, targets, mask, mask_img, block,
                   remap4, clamp_hi):
    S, M, D, H, W = vols.shape
    C = 2 * block + 1
    for b in range(inputs.shape[0]):
        s = int(subj[b])
        sl = min(max(int(slice_idx[b]), block), clamp_hi - block)
        sl = min(sl, D - 1 - block)
        for m in range(M):
            on = bool(present[s, m]) and int(drop[b]) != m
            mask[b, m] = 1.0 if on else 0.0
            win = vols[s, m, sl - block:sl + block + 1] if on else torch.zeros(C, H, W)
            if brain_mask is not None and on:
                win = win * brain_mask[sl - block:sl + block + 1]
            inputs[b, m * C:(m + 1) * C] = win
        if tvols is not None and bool(has_target[s]):
            t = tvols[s, sl].clone()
            if remap4:
                t[t == 4] = 3.0
            if brain_mask is not None:
                t = t * brain_mask[sl]
            targets[b, 0] = t
        else:
            targets[b, 0] = 0
        mask_img[b] = (inputs[b, 0] == 0).float()


_NAMES = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("ConvDesc",)]


def install():
    """Monkeypatch rd_b200.kernels with the functions above (CPU host-logic tests only)."""
    import rd_b200.kernels as K
    saved = {}
    for n in _NAMES:
        if hasattr(K, n) and n not in ("conv_desc",):
            saved[n] = getattr(K, n)
            setattr(K, n, globals()[n])
    saved["SEG_PARTIAL_FLOATS"] = K.SEG_PARTIAL_FLOATS
    return saved


def uninstall(saved):
    import rd_b200.kernels as K
    for n, v in saved.items():
        setattr(K, n, v)
