"""Autograd glue: one torch.autograd.Function per differentiable operator of the hot path.

Every forward / backward here is a sequence of rd_b200 kernel launches (through `kernels`); PyTorch
provides only the tape, the tensors and the stream.  Activations are NHWC tensors (N, H, W, C).
"Groups": the batch dimension is G * Ng; group g uses the g-th mixed CondConv kernel / the g-th set of
BatchNorm statistics — this is how the reference's per-modality Python loops (src/model.py:3135-3224)
are batched into single launches without changing the arithmetic.
"""
from typing import List, Optional, Sequence

import torch
from torch.autograd import Function

from . import kernels as K
from .lib import RD_ACT_LRELU, RD_ACT_NONE, RD_ALGO_AUTO

LRELU_SLOPE = 0.2


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def _sink(p):
    """Gradient sink: the trainer pre-assigns `p.grad` as a view into one flat fp32 buffer and sets
    `p._rd_sink`; backward kernels then accumulate straight into it (they all `+=`) and autograd gets None,
    so there is no per-parameter temporary, zero-fill or AccumulateGrad add launch."""
    if p is not None and getattr(p, "_rd_sink", False) and p.grad is not None and p.requires_grad:
        if SINK_HOOK is not None:
            SINK_HOOK(p)           # data-parallel: the gradient reducer counts the accumulations of each readiness stage (ddp.GradReducer)
        return p.grad              # frozen parameters (fix_pretrain, src/main_missing.py:104-116) have no sink: nothing is accumulated
    return None


SINK_HOOK = None


# Deferred mixing backward: when the trainer installs a kernels.MixBwdBatch here, _GroupedConv.backward queues its heads
# instead of launching one latency-bound kernel per head; the trainer flushes the queue (one launch) at the tape marker
# that follows the decoder backward and once more after the whole backward.
MIX_BATCH = None


# Up-front mixing: when the trainer installs a kernels.MixFwdPlan here, _GroupedConv.forward takes the packed weights the plan's
# one launch produced at the start of the iteration (and records layers the plan has not seen yet).
MIX_FWD = None


def flush_mix_bwd():
    if MIX_BATCH is not None:
        MIX_BATCH.flush()


class ConvHead:
    """Static (non-tensor) description of one CondConv2d / nn.Conv2d contributing output channels."""
    __slots__ = ("cond", "has_bias", "out_ch")

    def __init__(self, cond: bool, has_bias: bool, out_ch: int):
        self.cond, self.has_bias, self.out_ch = cond, has_bias, out_ch


class _PadChannels(Function):
    """Zero-pad the channel dimension (the tensor-core gathers need 16-byte channel vectors: 4 / 7 -> 8)."""

    @staticmethod
    def forward(ctx, x, c_pad):
        x = _c(x)
        out = torch.empty(x.shape[:-1] + (c_pad,), dtype=x.dtype, device=x.device)
        K.pad_channels(x, out)
        ctx.c = x.shape[-1]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        dx = torch.empty(dout.shape[:-1] + (ctx.c,), dtype=dout.dtype, device=dout.device)
        K.split_channels(dout, dx, None, ctx.c, dout.shape[-1] - ctx.c)
        return dx, None


def pad_channels(x, c_pad):
    if x.shape[-1] == c_pad:
        return x
    return _PadChannels.apply(x, int(c_pad))


def _up8(c):
    """Channel count the tensor-core kernels can gather: a multiple of 8 (16-byte vectors); counts below 16 go to 16 so
    that the TMA kernels (16-channel minimum box) cover the 4-channel anatomy codes and 7-channel image slabs too."""
    return 16 if c < 16 else (c + 7) // 8 * 8


class _GroupedConv(Function):
    """y = act(conv2d(x, mix(W, types)) + bias) for one or several heads sharing the input
    (CondConv2d.forward, reference src/model.py:2108-2117; heads > 1 fuses SPADE gamma & beta, :2444-2445).

    tensors: per MODULE, per head (W, fc_w, fc_b, bias) — fc_* / bias may be None.  `modules` > 1 batches the same layer
    of several nn.Modules into one launch (the per-modality decoder halves input_decoder_list[i], :3221-3222): the G
    weight groups split evenly over the modules, module m owns groups [m*G/modules, (m+1)*G/modules) and its own bias row.
    x may carry zero-padded channels (Cin_storage >= W's in_channels); in bf16 the backward pads dY to a
    multiple of 8 channels when the layer's output channel count is not one (4-channel logits, 7-channel images).
    """

    @staticmethod
    def forward(ctx, x, z, types, stride, pad, act, algo, heads, modules, *tensors):
        """z (optional, (N, H, W, C) with 2C = the output channels of the two heads gamma | beta): the SPADE modulation is fused into the
        convolution's epilogue (rd_conv2d_fwd_spade) and the function returns mix = IN(z) * (1 + gamma) + beta instead of gamma|beta."""
        x = _c(x)
        N, H, Wd, Cin = x.shape            # storage channels (>= logical in_channels)
        G = len(types)
        Gm = G // modules
        nh = len(heads)
        o_total = sum(h.out_ch for h in heads)
        o_pad = _up8(o_total) if x.dtype == torch.bfloat16 else o_total
        W0 = tensors[0]
        kh, kw = W0.shape[-2], W0.shape[-1]
        taps = kh * kw
        dev, dt = x.device, x.dtype
        any_bias = any(h.has_bias for h in heads)
        plan, key, hit = MIX_FWD, None, None
        if plan is not None:
            key = (tuple(t.data_ptr() if t is not None else 0 for t in tensors), tuple(float(t) for t in types), Cin, modules, str(dt), o_total)
            hit = plan.get(key)
        if hit is not None:
            packed, packedT, bias_all = hit
        else:
            packed = torch.empty((G, o_total, taps, Cin), dtype=dt, device=dev)
            packedT = (torch.empty((G, Cin, taps, o_pad), dtype=dt, device=dev) if o_pad == o_total
                       else K.zeros((G, Cin, taps, o_pad), dt, dev))
            bias_all = K.zeros((modules, o_total), torch.float32, dev) if any_bias else None
            jobs = []
            for m in range(modules):
                off = 0
                tm = types[m * Gm:(m + 1) * Gm]
                for hi, h in enumerate(heads):
                    W, fcw, fcb, b = tensors[4 * (m * nh + hi): 4 * (m * nh + hi) + 4]
                    pk, pkT = packed[m * Gm:(m + 1) * Gm], packedT[m * Gm:(m + 1) * Gm]
                    K.condconv_mix_fwd(W, fcw, fcb, tm, Cin, o_total, o_pad, off, pk, pkT, None)
                    bdst = None
                    if h.has_bias:
                        bdst = bias_all[m, off: off + h.out_ch]
                        K.cast(b, bdst)
                    jobs.append((W, fcw, fcb, tuple(float(t) for t in tm), Cin, o_total, o_pad, off, pk, pkT, b if h.has_bias else None, bdst))
                    off += h.out_ch
            if plan is not None:
                plan.register(key, packed, packedT, bias_all, jobs)
        bg = modules if modules > 1 else 0               # one bias row per module
        d = K.conv_desc(N, H, Wd, Cin, o_total, kh, kw, stride, pad, G, K._dt(x), act, LRELU_SLOPE, algo, bg)
        if z is not None:
            z = _c(z)
            Cz = z.shape[-1]
            mean = torch.empty(N * Cz, dtype=torch.float32, device=dev)
            invstd = torch.empty(N * Cz, dtype=torch.float32, device=dev)
            ws = K.norm_workspace(N, H * Wd, Cz, dev)
            K.norm_stats(z, N, H * Wd, Cz, 1e-5, ws, mean, invstd, None, None, None, 0.0)       # nn.InstanceNorm2d: per image, eps 1e-5
            gamma = torch.empty((N, d.oh, d.ow, Cz), dtype=dt, device=dev)
            mix = torch.empty((N, d.oh, d.ow, Cz), dtype=dt, device=dev)
            K.conv2d_fwd_spade(d, x, packed, bias_all, z, mean, invstd, gamma, mix)
            ctx.save_for_backward(x, packedT, None, z, gamma, mean, invstd, *tensors)
            ctx.fused = True
            ctx.meta = (types, stride, pad, act, algo, heads, modules, (N, H, Wd, Cin), o_total, o_pad, kh, kw)
            return mix
        y = torch.empty((N, d.oh, d.ow, o_total), dtype=dt, device=dev)
        K.conv2d_fwd(d, x, packed, bias_all, y)
        ctx.save_for_backward(x, packedT, y if act != RD_ACT_NONE else None, *tensors)
        ctx.fused = False
        ctx.meta = (types, stride, pad, act, algo, heads, modules, (N, H, Wd, Cin), o_total, o_pad, kh, kw)
        return y

    @staticmethod
    def backward(ctx, dy):
        types, stride, pad, act, algo, heads, modules, (N, H, Wd, Cin), o_total, o_pad, kh, kw = ctx.meta
        x, packedT, y = ctx.saved_tensors[:3]
        base = _padded_base(dy, o_pad) if (o_pad != o_total and act == RD_ACT_NONE and not ctx.fused) else None
        dy = base if base is not None else _c(dy)
        dz = None
        if ctx.fused:       # dy is d(mix): through the modulation first -> dz and d(gamma|beta), the latter continues as the conv's dy
            z, gamma, mean, invstd = ctx.saved_tensors[3:7]
            tensors = ctx.saved_tensors[7:]
            dz = torch.empty_like(z)
            dgb = torch.empty(z.shape[:-1] + (2 * z.shape[-1],), dtype=z.dtype, device=z.device)
            ws = K.spade_bwd_workspace(z)
            K.spade_modulate_bwd_g(z, mean, invstd, gamma, dy, dz, dgb, ws)
            dy = dgb
        else:
            tensors = ctx.saved_tensors[3:]
        G = len(types)
        Gm = G // modules
        nh = len(heads)
        dev = x.device
        if act == RD_ACT_LRELU:
            d_pre = torch.empty_like(dy)
            K.lrelu_bwd(dy, y, d_pre, LRELU_SLOPE)
            dy = d_pre
        if o_pad != o_total and base is None:
            dy_p = torch.empty(dy.shape[:-1] + (o_pad,), dtype=dy.dtype, device=dev)
            K.pad_channels(dy, dy_p)
            dy = dy_p
        bg = modules if modules > 1 else 0
        d = K.conv_desc(N, H, Wd, Cin, o_pad, kh, kw, stride, pad, G, K._dt(x), RD_ACT_NONE, LRELU_SLOPE, algo, bg)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            K.conv2d_dgrad(d, dy, packedT, dx)
        grads: List[Optional[torch.Tensor]] = [None] * len(tensors)
        need_w = any(ctx.needs_input_grad[9 + 4 * k] for k in range(modules * nh))
        if need_w:
            dK = torch.empty((G, o_pad, kh * kw, Cin), dtype=torch.float32, device=dev)
            any_bias = any(h.has_bias for h in heads)
            single_sink = (modules == 1 and nh == 1 and heads[0].has_bias and o_pad == o_total and _sink(tensors[3]) is not None)
            if single_sink:
                dbias_all = _sink(tensors[3])          # wgrad accumulates (+=) straight into bias.grad
            elif modules > 1:
                dbias_all = K.zeros((modules, o_pad), torch.float32, dev) if any_bias else None
            else:
                dbias_all = K.zeros((o_pad,), torch.float32, dev) if any_bias else None
            K.conv2d_wgrad(d, x, dy, dK, dbias_all)
            for m in range(modules):
                off = 0
                tm = types[m * Gm:(m + 1) * Gm]
                dKm = dK[m * Gm:(m + 1) * Gm]
                for hi, h in enumerate(heads):
                    base = 4 * (m * nh + hi)
                    W, fcw, fcb, b = tensors[base: base + 4]
                    sW, sfw, sfb = _sink(W), _sink(fcw), _sink(fcb)
                    dW = sW if sW is not None else torch.zeros_like(W)
                    dfw = (sfw if sfw is not None else torch.zeros_like(fcw)) if fcw is not None else None
                    dfb = (sfb if sfb is not None else torch.zeros_like(fcb)) if fcb is not None else None
                    bias_riding = False
                    if MIX_BATCH is not None and sW is not None and (fcw is None or (sfw is not None and sfb is not None)):
                        b_src = b_dst = None
                        if h.has_bias and not single_sink and _sink(b) is not None:
                            # the head's slice of the launch's bias-gradient row is added into bias.grad by the batched launch
                            b_src = (dbias_all[m, off: off + h.out_ch] if modules > 1 else dbias_all[off: off + h.out_ch])
                            b_dst = _sink(b)
                            bias_riding = b_src.is_contiguous() and b_dst.is_contiguous()
                        if bias_riding:
                            MIX_BATCH.add(dKm, W, fcw, fcb, tm, Cin, o_pad, off, dW, dfw, dfb, b_src, b_dst)
                        else:
                            MIX_BATCH.add(dKm, W, fcw, fcb, tm, Cin, o_pad, off, dW, dfw, dfb)
                    else:
                        K.condconv_mix_bwd(dKm, W, fcw, fcb, tm, Cin, o_pad, off, dW, dfw, dfb)
                    grads[base] = None if sW is not None else dW
                    grads[base + 1] = None if sfw is not None else dfw
                    grads[base + 2] = None if sfb is not None else dfb
                    if h.has_bias and not single_sink and not bias_riding:
                        if modules > 1:        # the wgrad kernel accumulated the module's groups into its own row
                            db = dbias_all[m, off: off + h.out_ch]
                        else:
                            db = dbias_all[off: off + h.out_ch]
                        sb = _sink(b)
                        if sb is not None:
                            K.add(sb, _c(db), sb)
                        else:
                            grads[base + 3] = db
                    off += h.out_ch
        return (dx, dz, None, None, None, None, None, None, None, *grads)


def grouped_conv(x, types: Sequence[float], stride: int, pad: int, heads: List[ConvHead], tensors: List,
                 act: int = RD_ACT_NONE, algo: int = RD_ALGO_AUTO, modules: int = 1):
    """`tensors`: for each of the `modules` nn.Modules, for each head: (W, fc_w, fc_b, bias)."""
    if x.dtype == torch.bfloat16 and x.shape[-1] != _up8(x.shape[-1]):
        x = pad_channels(x, _up8(x.shape[-1]))     # 4-channel anatomy codes, 7-channel image slabs -> 16
    if len(types) % modules or len(tensors) != 4 * len(heads) * modules:
        raise ValueError("grouped_conv: types / tensors do not split over %d modules" % modules)
    return _GroupedConv.apply(x, None, tuple(float(t) for t in types), stride, pad, act, algo, tuple(heads), int(modules), *tensors)


def spade_conv(a, z, types: Sequence[float], heads: List[ConvHead], tensors: List, modules: int = 1, eps: float = 1e-5):
    """SPADEBlockNew's middle (src/model.py:2444-2452): mix = InstanceNorm(z) * (1 + gamma(a)) + beta(a), gamma and beta one 3x3 convolution
    with 2C output channels (heads = [gamma, beta]).  Where the kernel layer can fuse the modulation into the convolution's epilogue
    (rd_conv2d_fwd_spade: bf16, weights resident in shared memory) only gamma and mix are written; elsewhere the convolution writes
    gamma|beta and spade_modulate reads it back."""
    if a.dtype == torch.bfloat16 and a.shape[-1] != _up8(a.shape[-1]):
        a = pad_channels(a, _up8(a.shape[-1]))
    N, H, Wd, Cin = a.shape
    Cz = z.shape[-1]
    o_total = sum(h.out_ch for h in heads)
    if o_total == 2 * Cz and eps == 1e-5 and tuple(z.shape[:3]) == (N, H, Wd):
        d = K.conv_desc(N, H, Wd, Cin, o_total, 3, 3, 1, 1, len(types), K._dt(a), RD_ACT_NONE, LRELU_SLOPE, RD_ALGO_AUTO,
                        modules if modules > 1 else 0)
        if K.conv2d_fwd_spade_supported(d, a):
            return _GroupedConv.apply(a, z, tuple(float(t) for t in types), 1, 1, RD_ACT_NONE, RD_ALGO_AUTO, tuple(heads), int(modules), *tensors)
    gb = grouped_conv(a, types, 1, 1, heads, tensors, modules=modules)
    return spade_modulate(z, gb, eps)


# Composition of the last two convolutions of a decoder half (reference src/model.py:2606-2612): SPADEBlockNew sp6 ends with
# `out` (3x3, C -> C') and is followed by the CondConv 1x1 (C' -> in_num_ch) with NOTHING in between, so
#     y = W_B * (W_A * x + b_A) + b_B = (W_B W_A) * x + (W_B b_A + b_B)
# is ONE 3x3 convolution with per-group weights W_eff[g] = mix(W_B)[g] mix(W_A)[g].  It replaces an N = 16 tensor-core launch
# plus three streaming 1x1 kernels per step by one N = 16 launch; the weight-space products are tiny (7 x 16 x 288 per group).
# Experimental (RD_B200_COMPOSE_OUT=1 / ops.COMPOSE_OUT): the host logic is pinned on the emulated kernel layer against the
# reference fixtures; GPU timing and bf16 parity are next round's first measurement.  Off by default.
import os as _os
COMPOSE_OUT = _os.environ.get("RD_B200_COMPOSE_OUT", "1") not in ("", "0")


class _ComposedOutConv(Function):
    """tensors: per module (W_A, fcw_A, fcb_A, b_A, W_B, fcw_B, fcb_B, b_B); A = 3x3 pad 1, B = 1x1; both CondConv.
    The weight-space products and their chain rule are rd_compose_tail_fwd / _bwd (csrc/rd_compose.cu)."""

    @staticmethod
    def forward(ctx, x, types, modules, *tensors):
        x = _c(x)
        N, H, Wd, Cin = x.shape
        G = len(types)
        Gm = G // modules
        dev, dt = x.device, x.dtype
        WA0, WB0 = tensors[0], tensors[4]
        OA, OB = WA0.shape[1], WB0.shape[1]
        kh, kw = WA0.shape[-2], WA0.shape[-1]
        taps = kh * kw
        f32 = torch.float32
        pA = torch.empty((G, OA, taps, Cin), dtype=f32, device=dev)
        pB = torch.empty((G, OB, 1, OA), dtype=f32, device=dev)
        has_bA = any(tensors[8 * m + 3] is not None for m in range(modules))
        has_bB = any(tensors[8 * m + 7] is not None for m in range(modules))
        bA = K.zeros((modules, OA), f32, dev) if has_bA else None
        bB = K.zeros((modules, OB), f32, dev) if has_bB else None
        for m in range(modules):
            WA, fwA, fbA, biasA, WB, fwB, fbB, biasB = tensors[8 * m: 8 * m + 8]
            tm = types[m * Gm:(m + 1) * Gm]
            K.condconv_mix_fwd(WA, fwA, fbA, tm, Cin, OA, OA, 0, pA[m * Gm:(m + 1) * Gm], None, None)
            K.condconv_mix_fwd(WB, fwB, fbB, tm, OA, OB, OB, 0, pB[m * Gm:(m + 1) * Gm], None, None)
            if biasA is not None:
                K.cast(biasA, bA[m])
            if biasB is not None:
                K.cast(biasB, bB[m])
        o_pad = _up8(OB) if dt == torch.bfloat16 else OB
        packed = torch.empty((G, OB, taps, Cin), dtype=dt, device=dev)
        packedT = torch.empty((G, Cin, taps, o_pad), dtype=dt, device=dev)
        b_eff = torch.empty((G, OB), dtype=f32, device=dev)                                   # one bias row per weight group
        K.compose_tail_fwd(pA, pB.view(G, OB, OA), bA, bB, modules, packed, packedT, b_eff)
        d = K.conv_desc(N, H, Wd, Cin, OB, kh, kw, 1, (kh - 1) // 2, G, K._dt(x), RD_ACT_NONE, LRELU_SLOPE, RD_ALGO_AUTO, G)
        y = torch.empty((N, d.oh, d.ow, OB), dtype=dt, device=dev)
        K.conv2d_fwd(d, x, packed, b_eff, y)
        ctx.save_for_backward(x, packedT, pA, pB, bA, *tensors)
        ctx.meta = (types, modules, (N, H, Wd, Cin), OA, OB, o_pad, kh, kw)
        return y

    @staticmethod
    def backward(ctx, dy):
        types, modules, (N, H, Wd, Cin), OA, OB, o_pad, kh, kw = ctx.meta
        x, packedT, pA, pB, bA = ctx.saved_tensors[:5]
        tensors = ctx.saved_tensors[5:]
        base = _padded_base(dy, o_pad) if o_pad != OB else None        # already the zero-padded dY (written by _SplitBlocks.backward)
        dy = base if base is not None else _c(dy)
        G = len(types)
        Gm = G // modules
        taps = kh * kw
        dev = x.device
        f32 = torch.float32
        if o_pad != OB and base is None:
            dy_p = torch.empty(dy.shape[:-1] + (o_pad,), dtype=dy.dtype, device=dev)
            K.pad_channels(dy, dy_p)
            dy = dy_p
        d = K.conv_desc(N, H, Wd, Cin, o_pad, kh, kw, 1, (kh - 1) // 2, G, K._dt(x), RD_ACT_NONE, LRELU_SLOPE, RD_ALGO_AUTO, G)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            K.conv2d_dgrad(d, dy, packedT, dx)
        dK = torch.empty((G, o_pad, taps, Cin), dtype=f32, device=dev)
        db = K.zeros((G, o_pad), f32, dev)
        K.conv2d_wgrad(d, x, dy, dK, db)
        # chain rule through W_eff = W_B W_A and b_eff = W_B b_A + b_B (weight-space products, fp32)
        dpA = torch.empty((G, OA, taps, Cin), dtype=f32, device=dev)
        dpB = torch.empty((G, OB, 1, OA), dtype=f32, device=dev)
        dbA = K.zeros((modules, OA), f32, dev) if bA is not None else None
        dbB = K.zeros((modules, OB), f32, dev) if any(tensors[8 * m + 7] is not None for m in range(modules)) else None
        K.compose_tail_bwd(dK, db, pA, pB.view(G, OB, OA), bA, modules, dpA, dpB.view(G, OB, OA), dbA, dbB)
        grads: List[Optional[torch.Tensor]] = [None] * len(tensors)
        for m in range(modules):
            WA, fwA, fbA, biasA, WB, fwB, fbB, biasB = tensors[8 * m: 8 * m + 8]
            tm = types[m * Gm:(m + 1) * Gm]
            sl = slice(m * Gm, (m + 1) * Gm)
            for k, (W, fw, fb, bias, dKp, i_pad, O, dbias) in enumerate((
                    (WA, fwA, fbA, biasA, dpA[sl], Cin, OA, dbA[m] if dbA is not None else None),
                    (WB, fwB, fbB, biasB, dpB[sl], OA, OB, dbB[m] if dbB is not None else None))):
                base = 8 * m + 4 * k
                sW, sfw, sfb = _sink(W), _sink(fw), _sink(fb)
                dW = sW if sW is not None else torch.zeros_like(W)
                dfw = (sfw if sfw is not None else torch.zeros_like(fw)) if fw is not None else None
                dfb = (sfb if sfb is not None else torch.zeros_like(fb)) if fb is not None else None
                sb = _sink(bias) if bias is not None else None
                riding = False
                if MIX_BATCH is not None and sW is not None and (fw is None or (sfw is not None and sfb is not None)):
                    if sb is not None and sb.is_contiguous():       # the bias gradient rides on the batched mixing launch
                        MIX_BATCH.add(dKp, W, fw, fb, tm, i_pad, O, 0, dW, dfw, dfb, dbias, sb)
                        riding = True
                    else:
                        MIX_BATCH.add(dKp, W, fw, fb, tm, i_pad, O, 0, dW, dfw, dfb)
                else:
                    K.condconv_mix_bwd(dKp, W, fw, fb, tm, i_pad, O, 0, dW, dfw, dfb)
                grads[base] = None if sW is not None else dW
                grads[base + 1] = None if sfw is not None else dfw
                grads[base + 2] = None if sfb is not None else dfb
                if bias is not None and not riding:
                    if sb is not None:
                        K.add(sb, dbias, sb)
                    else:
                        grads[base + 3] = dbias
        return (dx, None, None, *grads)


def composed_out_conv(x, types: Sequence[float], modules: int, tensors: List):
    """conv1x1_B(conv3x3_A(x)) of `modules` decoder halves as one grouped 3x3 convolution; tensors: per module the four tensors of
    A (W, fc_w, fc_b, bias) then the four of B."""
    if len(types) % modules or len(tensors) != 8 * modules:
        raise ValueError("composed_out_conv: types / tensors do not split over %d modules" % modules)
    return _ComposedOutConv.apply(x, tuple(float(t) for t in types), int(modules), *tensors)


class _GroupNorm(Function):
    """Train-mode BatchNorm2d applied independently to G batch groups (one reference module call per
    group, src/model.py:2151,2191), or eval-mode BatchNorm with running statistics."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, nbt, G, training, momentum, eps):
        x = _c(x)
        N, H, Wd, Cn = x.shape
        ppg = (N // G) * H * Wd
        dev = x.device
        mean = torch.empty(G * Cn, dtype=torch.float32, device=dev)
        invstd = torch.empty(G * Cn, dtype=torch.float32, device=dev)
        if training:
            ws = K.norm_workspace(G, ppg, Cn, dev)
            K.norm_stats(x, G, ppg, Cn, eps, ws, mean, invstd, running_mean, running_var, nbt, momentum)
        else:
            K.norm_eval_stats(running_mean, running_var, G, eps, mean, invstd)
        y = torch.empty_like(x)
        K.norm_apply(x, mean, invstd, weight, bias, y, G, ppg, Cn)
        ctx.save_for_backward(x, mean, invstd, weight)
        ctx.bias_ref = bias
        ctx.meta = (G, ppg, Cn, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, invstd, weight = ctx.saved_tensors
        G, ppg, Cn, training = ctx.meta
        if not training:
            raise RuntimeError("rd_b200: backward through eval-mode BatchNorm is not part of the hot path")
        dy = _c(dy)
        dx = torch.empty_like(x)
        sw, sb = _sink(weight), _sink(ctx.bias_ref)
        dw = sw if sw is not None else torch.zeros_like(weight)
        db = sb if sb is not None else torch.zeros_like(weight)
        ws = K.norm_workspace(G, ppg, Cn, x.device)
        K.norm_bwd(x, dy, mean, invstd, weight, dx, dw, db, ws, G, ppg, Cn)
        return dx, (None if sw is not None else dw), (None if sb is not None else db), None, None, None, None, None, None, None


def group_batch_norm(x, weight, bias, running_mean, running_var, nbt, G, training, momentum=0.1, eps=1e-5):
    return _GroupNorm.apply(x, weight, bias, running_mean, running_var, nbt, G, training, momentum, eps)


class _SpadeModulate(Function):
    """mix = InstanceNorm2d(z) * (1 + gamma) + beta with gb = [gamma | beta] (src/model.py:2440-2452)."""

    @staticmethod
    def forward(ctx, z, gb, eps):
        z, gb = _c(z), _c(gb)
        N, H, Wd, Cn = z.shape
        dev = z.device
        mean = torch.empty(N * Cn, dtype=torch.float32, device=dev)
        invstd = torch.empty(N * Cn, dtype=torch.float32, device=dev)
        ws = K.norm_workspace(N, H * Wd, Cn, dev)
        K.norm_stats(z, N, H * Wd, Cn, eps, ws, mean, invstd, None, None, None, 0.0)
        mix = torch.empty_like(z)
        K.spade_modulate_fwd(z, mean, invstd, gb, mix)
        ctx.save_for_backward(z, gb, mean, invstd)
        return mix

    @staticmethod
    def backward(ctx, dmix):
        z, gb, mean, invstd = ctx.saved_tensors
        dmix = _c(dmix)
        N, H, Wd, Cn = z.shape
        dz = torch.empty_like(z)
        dgb = torch.empty_like(gb)
        ws = K.spade_bwd_workspace(z)
        K.spade_modulate_bwd(z, mean, invstd, gb, dmix, dz, dgb, ws)
        return dz, dgb, None


def spade_modulate(z, gb, eps=1e-5):
    return _SpadeModulate.apply(z, gb, eps)


class _Bilinear(Function):
    @staticmethod
    def forward(ctx, x, oh, ow, align):
        x = _c(x)
        N, H, Wd, Cn = x.shape
        y = torch.empty((N, oh, ow, Cn), dtype=x.dtype, device=x.device)
        K.bilinear_fwd(x, y, align)
        ctx.meta = (tuple(x.shape), align)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, align = ctx.meta
        dy = _c(dy)
        dx = torch.empty(shape, dtype=dy.dtype, device=dy.device)
        K.bilinear_bwd(dy, dx, align)
        return dx, None, None, None


def bilinear(x, oh, ow, align_corners: bool):
    if int(oh) == x.shape[1] and int(ow) == x.shape[2]:
        return x            # same size: src == dst and l1 == 0 in both conventions, the resize is the identity (bit exact)
    return _Bilinear.apply(x, int(oh), int(ow), bool(align_corners))


class _MaskedSoftmax(Function):
    """softmax_c([100*mask_img, s])[:, 1:]  (src/model.py:3149-3153); mask_img None = plain softmax."""

    @staticmethod
    def forward(ctx, s, mask_img):
        s = _c(s)
        p = torch.empty_like(s)
        K.masked_softmax_fwd(s, _c(mask_img), p)
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        ds = torch.empty_like(p)
        K.masked_softmax_bwd(p, _c(dp), ds)
        return ds, None


def masked_softmax(s, mask_img):
    return _MaskedSoftmax.apply(s, mask_img)


class _ConcatChannels(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        out = torch.empty(a.shape[:-1] + (a.shape[-1] + b.shape[-1],), dtype=a.dtype, device=a.device)
        K.concat_channels(a, b, out)
        ctx.meta = (a.shape[-1], b.shape[-1])
        return out

    @staticmethod
    def backward(ctx, dout):
        ca, cb = ctx.meta
        dout = _c(dout)
        da = torch.empty(dout.shape[:-1] + (ca,), dtype=dout.dtype, device=dout.device) if ctx.needs_input_grad[0] else None
        db = torch.empty(dout.shape[:-1] + (cb,), dtype=dout.dtype, device=dout.device) if ctx.needs_input_grad[1] else None
        K.split_channels(dout, da, db, ca, cb)
        return da, db


def concat_channels(a, b):
    return _ConcatChannels.apply(a, b)


class _Cast(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        x = _c(x)
        ctx.src_dtype = x.dtype
        if x.dtype == dtype:
            return x
        y = torch.empty(x.shape, dtype=dtype, device=x.device)
        K.cast(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        if dy.dtype == ctx.src_dtype:
            return dy, None
        dx = torch.empty(dy.shape, dtype=ctx.src_dtype, device=dy.device)
        K.cast(dy, dx)
        return dx, None


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    return _Cast.apply(x, dtype)


class _Linear(Function):
    """nn.Linear (+ fused LeakyReLU(slope): 0.2 in the encoders, 0 = ReLU in the channel attention) in fp32
    (src/model.py:2359-2364, 2499, 1428)."""

    @staticmethod
    def forward(ctx, x, W, b, act, slope):
        x = _c(x)
        y = torch.empty((x.shape[0], W.shape[0]), dtype=torch.float32, device=x.device)
        K.linear_fwd(x, W, b, y, act, slope)
        ctx.save_for_backward(x, W, y if act != RD_ACT_NONE else None)
        ctx.act, ctx.slope = act, slope
        ctx.bias_ref = b
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = _c(dy)
        if ctx.act == RD_ACT_LRELU:
            t = torch.empty_like(dy)
            K.lrelu_bwd(dy, y, t, ctx.slope)
            dy = t
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        sW, sb = _sink(W), _sink(ctx.bias_ref)
        dW = sW if sW is not None else torch.zeros_like(W)
        db = sb if sb is not None else K.zeros((W.shape[0],), torch.float32, x.device)
        K.linear_bwd(x, W, dy, dx, dW, db)
        return dx, (None if sW is not None else dW), (None if sb is not None else db), None, None


def linear(x, W, b, act=RD_ACT_NONE, slope=LRELU_SLOPE):
    return _Linear.apply(x, W, b, act, float(slope))


class _Sample(Function):
    """z = mu + eps * exp(0.5 * log_var)  (MultimodalModel.sample, src/model.py:3159-3162)."""

    @staticmethod
    def forward(ctx, mu, lv, eps):
        mu, lv, eps = _c(mu), _c(lv), _c(eps)
        z = torch.empty_like(mu)
        K.sample_fwd(mu, lv, eps, z)
        ctx.save_for_backward(lv, eps)
        return z

    @staticmethod
    def backward(ctx, dz):
        lv, eps = ctx.saved_tensors
        dz = _c(dz)
        dmu, dlv = torch.empty_like(dz), torch.empty_like(dz)
        K.sample_bwd(dz, lv, eps, dmu, dlv)
        return dmu, dlv, None


def sample(mu, lv, eps):
    return _Sample.apply(mu, lv, eps)


class _FuseGather(Function):
    """si_cat[mask == 1] of reconstruct_output_si_fused (src/model.py:3241-3242).
    si: modality-major stack (M*B, H, W, C); returns (rows (B*M, H, W, C) [first K valid], idx, count)."""

    @staticmethod
    def forward(ctx, si, mask, B, M):
        si, mask = _c(si), _c(mask)
        out = torch.zeros_like(si)
        idx = torch.empty(B * M, dtype=torch.int32, device=si.device)
        cnt = torch.empty(1, dtype=torch.int32, device=si.device)
        K.fuse_gather_fwd(si, mask, out, idx, cnt, B, M)
        ctx.save_for_backward(mask)
        ctx.meta = (B, M)
        ctx.mark_non_differentiable(idx, cnt)
        return out, idx, cnt

    @staticmethod
    def backward(ctx, dout, _a, _b):
        (mask,) = ctx.saved_tensors
        B, M = ctx.meta
        dout = _c(dout)
        dsi = torch.empty_like(dout)
        K.fuse_gather_bwd(dout, mask, dsi, B, M)
        return dsi, None, None, None


def fuse_gather(si, mask, B, M):
    return _FuseGather.apply(si, mask, B, M)


# ------------------------------------------------------------------------------------------------ losses
class _MaskedRecon(Function):
    """compute_recon_loss_x_list (kind 0, src/model.py:3315-3325) / compute_recon_loss_x_mix_list (kind 1,
    :3327-3341 incl. the idx lag Q4) on stacked rows: x (R*B rows), gt (M*B rows)."""

    @staticmethod
    def forward(ctx, x, gt, mask, B, M, kind, p):
        x, gt, mask = _c(x), _c(gt), _c(mask)
        dev = x.device
        R = x.shape[0]
        row_elems = x.numel() // R
        gt_index = None
        if kind == 1:
            gt_index = torch.empty(R, dtype=torch.int32, device=dev)
            K.xmix_plan(mask, gt_index, B, M)
        row_loss = torch.empty(R, dtype=torch.float32, device=dev)
        partial = torch.empty(R * K.recon_chunks(row_elems), dtype=torch.float32, device=dev)
        K.recon_rows_fwd(x, gt, gt_index, row_loss, partial, R, p)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        coef = torch.empty(R, dtype=torch.float32, device=dev)
        K.masked_combine(row_loss, mask, loss, coef, B, M, kind)
        ctx.save_for_backward(x, gt, gt_index, coef)
        ctx.meta = (R, p)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        x, gt, gt_index, coef = ctx.saved_tensors
        R, p = ctx.meta
        c2 = torch.empty_like(coef)
        # coef * upstream (scalar broadcast): a 1-element row "linear" keeps this on our kernels
        K.linear_fwd(coef.reshape(R, 1), _c(dloss.reshape(1, 1).float()), None, c2.reshape(R, 1), RD_ACT_NONE, LRELU_SLOPE)
        dx = torch.empty_like(x)
        K.recon_rows_bwd(x, gt, gt_index, c2, dx, R, p)
        return dx, None, None, None, None, None, None


def masked_recon_loss(x, gt, mask, B, M, kind, p):
    return _MaskedRecon.apply(x, gt, mask, B, M, kind, p)


class _ScaledGradLoss(Function):
    """Helper for the tiny losses whose kernels return loss and gradients together: the forward kernel
    writes dL/dinput; backward scales it by the upstream scalar."""

    @staticmethod
    def forward(ctx, fn, *ins):
        ins = [_c(t) for t in ins]
        loss, grads = fn(*ins)          # launches the loss kernel: loss (1,) and d loss / d input for every input
        ctx.save_for_backward(*grads)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        outs = []
        up = _c(dloss.reshape(1, 1).float())
        for g in ctx.saved_tensors:
            o = torch.empty_like(g)
            K.linear_fwd(g.reshape(-1, 1), up, None, o.reshape(-1, 1), RD_ACT_NONE, LRELU_SLOPE)
            outs.append(o)
        return (None, *outs)


def latent_z_loss(mu, mu_new, mask, B, M, Z):
    """compute_latent_z_loss (src/model.py:3384-3394) on (M, B, Z) stacks."""
    mask = _c(mask)

    def fn(a, b):
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        da, db = torch.empty_like(a), torch.empty_like(b)
        K.latent_z_loss(a, b, mask, loss, da, db, B, M, Z)
        return loss, (da, db)
    return _ScaledGradLoss.apply(fn, mu, mu_new)


def sim_z_loss(z, mask, margin, B, M, Z):
    """compute_similarity_z_loss (src/model.py:3537-3557)."""
    mask = _c(mask)

    def fn(a):
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        da = torch.empty_like(a)
        K.sim_z_loss(a, mask, margin, loss, da, B, M, Z)
        return loss, (da,)
    return _ScaledGradLoss.apply(fn, z)


def kl_loss(mu, lv, mask, B, M, Z):
    """compute_kl_loss_list_standard (src/model.py:3343-3360)."""
    mask = _c(mask)

    def fn(a, b):
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        da, db = torch.empty_like(a), torch.empty_like(b)
        K.kl_loss(a, b, mask, loss, da, db, B, M, Z)
        return loss, (da, db)
    return _ScaledGradLoss.apply(fn, mu, lv)


class _MaxPool16(Function):
    """compute_compact_s_max (src/model.py:3448-3451) on an NHWC stack; output (N, C*H/16*W/16) fp32."""

    @staticmethod
    def forward(ctx, s):
        s = _c(s)
        N, H, Wd, Cn = s.shape
        D = Cn * (H // 16) * (Wd // 16)
        pooled = torch.empty((N, D), dtype=torch.float32, device=s.device)
        arg = torch.empty((N, D), dtype=torch.int32, device=s.device)
        K.maxpool16_fwd(s, pooled, arg)
        ctx.save_for_backward(arg)
        ctx.meta = (tuple(s.shape), s.dtype)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        (arg,) = ctx.saved_tensors
        shape, dt = ctx.meta
        ds = torch.empty(shape, dtype=dt, device=dpooled.device)
        K.maxpool16_bwd(_c(dpooled), arg, ds)
        return ds


def maxpool16(s):
    return _MaxPool16.apply(s)


class _AvgPool16(Function):
    """compute_compact_s_mean (src/model.py:3453-3456) on an NHWC stack; output (N, C*H/16*W/16) fp32."""

    @staticmethod
    def forward(ctx, s):
        s = _c(s)
        N, H, Wd, Cn = s.shape
        pooled = torch.empty((N, Cn * (H // 16) * (Wd // 16)), dtype=torch.float32, device=s.device)
        K.avgpool16_fwd(s, pooled)
        ctx.meta = (tuple(s.shape), s.dtype)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        shape, dt = ctx.meta
        ds = torch.empty(shape, dtype=dt, device=dpooled.device)
        K.avgpool16_bwd(_c(dpooled), ds)
        return ds


def avgpool16(s):
    return _AvgPool16.apply(s)


class _Softplus(Function):
    """F.softplus (src/model.py:2631 out_act, :3145 ana_dec_act, target_output_act)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        K.softplus_fwd(x, y)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        K.softplus_bwd(_c(dy), x, dx)
        return dx


def softplus(x):
    return _Softplus.apply(x)


def sim_s_loss(pooled, mask, pair_dev, margin, B, M):
    """compute_similarity_s_loss (src/model.py:3478-3513) on pooled (M*B, D) vectors; pair_dev = device int32[2]."""
    mask = _c(mask)
    D = pooled.shape[-1]

    def fn(a):
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        da = torch.empty_like(a)
        K.sim_s_loss(a, mask, pair_dev, margin, loss, da, B, M, D)
        return loss, (da,)
    return _ScaledGradLoss.apply(fn, pooled)


class _SegLoss(Function):
    """compute_segmentation_loss_y (src/model.py:3287-3297): weighted CE + soft Dice, y NHWC with 4 classes."""

    @staticmethod
    def forward(ctx, y, target):
        y, target = _c(y), _c(target)
        loss = torch.empty(1, dtype=torch.float32, device=y.device)
        partial = torch.empty(K.SEG_PARTIAL_FLOATS, dtype=torch.float32, device=y.device)
        K.seg_loss_fwd(y, target, loss, partial)
        ctx.save_for_backward(y, target, partial)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dloss):
        y, target, partial = ctx.saved_tensors
        dy = torch.empty_like(y)
        K.seg_loss_bwd(y, target, partial, _c(dloss.reshape(1).float()), dy)
        return dy, None


def seg_loss(y, target):
    return _SegLoss.apply(y, target)


class _WeightedSum(Function):
    """total = sum_k lambda_k * loss_k over 0-dim fp32 losses, as one small linear launch each way."""

    @staticmethod
    def forward(ctx, lambdas, *losses):
        v = torch.stack([l.reshape(()) for l in losses]).reshape(1, -1)   # view/stack of scalars (plumbing)
        out = torch.empty((1, 1), dtype=torch.float32, device=v.device)
        K.linear_fwd(_c(v), lambdas.reshape(1, -1), None, out, RD_ACT_NONE, LRELU_SLOPE)
        ctx.save_for_backward(lambdas)
        ctx.n = len(losses)
        return out.reshape(())

    @staticmethod
    def backward(ctx, dtotal):
        (lambdas,) = ctx.saved_tensors
        g = torch.empty((ctx.n, 1), dtype=torch.float32, device=lambdas.device)
        K.linear_fwd(lambdas.reshape(-1, 1), _c(dtotal.reshape(1, 1).float()), None, g, RD_ACT_NONE, LRELU_SLOPE)
        return (None, *[g[k, 0] for k in range(ctx.n)])


def weighted_sum(lambdas: torch.Tensor, losses: Sequence[torch.Tensor]):
    return _WeightedSum.apply(lambdas, *losses)


# ------------------------------------------------------------------------------------------------ layout + rows
class _ToNHWC(Function):
    """Logical NCHW fp32 (contiguous, or a channel slice of a contiguous NCHW tensor such as
    inputs[:, 7*i:7*(i+1)], src/main_missing.py:166-168) -> NHWC tensor (N,H,W,C) in `dtype`."""

    @staticmethod
    def forward(ctx, x, dtype):
        N, Cn, H, Wd = x.shape
        st = x.stride()
        if not (x.dtype == torch.float32 and st[3] == 1 and st[2] == Wd and st[1] == H * Wd and st[0] % (H * Wd) == 0
                and st[0] >= Cn * H * Wd):
            x = x.float().contiguous()
            st = x.stride()
        y = torch.empty((N, H, Wd, Cn), dtype=dtype, device=x.device)
        K.nchw_to_nhwc_strided(x, y, st[0] // (H * Wd))
        ctx.src_shape = (N, Cn, H, Wd)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty(ctx.src_shape, dtype=torch.float32, device=dy.device)
        K.nhwc_to_nchw(dy, dx)
        return dx, None


class _ToNCHW(Function):
    """NHWC (N,H,W,C) any dtype -> contiguous NCHW fp32 (used before flatten / nn.Linear, src/model.py:2396)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        N, H, Wd, Cn = x.shape
        y = torch.empty((N, Cn, H, Wd), dtype=torch.float32, device=x.device)
        K.nhwc_to_nchw(x, y)
        ctx.meta = (x.dtype,)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        N, Cn, H, Wd = dy.shape
        dx = torch.empty((N, H, Wd, Cn), dtype=ctx.meta[0], device=dy.device)
        K.nchw_to_nhwc_strided(dy, dx, Cn)
        return dx


def to_nhwc(x, dtype):
    """Accepts a logical NCHW tensor.  Zero-copy when it already is a permuted NHWC tensor."""
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous():
        return xp if xp.dtype == dtype else cast(xp, dtype)
    return _ToNHWC.apply(x, dtype)


def to_nchw_f32(x):
    return _ToNCHW.apply(x)


class _GatherBlocks(Function):
    """out block k = src block index[k]; blocks are `block` consecutive rows (images) of the leading
    dimension; optional zero padding of the channel (last) dimension to `c_pad`.  One launch each way: the
    backward sums the fan-out of s_i / z_j over the (i, j) decodes (src/model.py:3187-3224) in fp32."""

    @staticmethod
    def forward(ctx, src, index, block, c_pad):
        src = _c(src)
        nb = len(index)
        cp = c_pad if c_pad else src.shape[-1]
        out = torch.empty((nb * block,) + tuple(src.shape[1:-1]) + (cp,), dtype=src.dtype, device=src.device)
        K.gather_blocks_fwd(src, out, index, block)
        ctx.meta = (tuple(index), block, tuple(src.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        index, block, shape = ctx.meta
        dout = _c(dout)
        dsrc = torch.empty(shape, dtype=dout.dtype, device=dout.device)
        K.gather_blocks_bwd(dout, dsrc, index, block)
        return dsrc, None, None, None


def gather_blocks(src, index, block, c_pad=None):
    index = tuple(int(i) for i in index)
    if len(index) > 32:
        raise ValueError("gather_blocks: at most 32 blocks per launch")
    if src.dim() < 2:
        raise ValueError("gather_blocks: need (rows, ..., channels)")
    return _GatherBlocks.apply(src, index, int(block), int(c_pad) if c_pad else 0)


class _StackRows(Function):
    """Concatenate tensors along dim 0 with our copy kernel; backward hands out views."""

    @staticmethod
    def forward(ctx, *parts):
        parts = [_c(p) for p in parts]
        n = sum(p.shape[0] for p in parts)
        out = torch.empty((n,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype, device=parts[0].device)
        o = 0
        sizes = []
        for p in parts:
            K.cast(p, out[o:o + p.shape[0]])
            sizes.append(p.shape[0])
            o += p.shape[0]
        ctx.sizes = sizes
        return out

    @staticmethod
    def backward(ctx, dout):
        outs, o = [], 0
        for n in ctx.sizes:
            outs.append(dout[o:o + n])
            o += n
        return tuple(outs)


def stack_rows(parts):
    parts = list(parts)
    if len(parts) == 1:
        return parts[0]
    return _StackRows.apply(*parts)


# ------------------------------------------------------------------------------------------------ attention gate
class _AddRelu(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        y = torch.empty_like(a)
        K.add_relu_fwd(a, b, y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        d = torch.empty_like(y)
        K.relu_bwd(_c(dy), y, d)
        return d, d


def add_relu(a, b):
    return _AddRelu.apply(a, b)


class _Sigmoid(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        K.sigmoid_fwd(x, y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        d = torch.empty_like(y)
        K.sigmoid_bwd(_c(dy), y, d)
        return d


def sigmoid(x):
    return _Sigmoid.apply(x)


class _MulBcast(Function):
    """(off + alpha) (N,H,W,1) * x (N,H,W,C)  (src/model.py:1326; off = 1: the residual form of :1414)."""

    @staticmethod
    def forward(ctx, alpha, x, off):
        alpha, x = _c(alpha), _c(x)
        y = torch.empty_like(x)
        K.mul_bcast_fwd(alpha, x, y, off)
        ctx.save_for_backward(alpha, x)
        ctx.off = off
        return y

    @staticmethod
    def backward(ctx, dy):
        alpha, x = ctx.saved_tensors
        dx, da = torch.empty_like(x), torch.empty_like(alpha)
        K.mul_bcast_bwd(alpha, x, _c(dy), dx, da, ctx.off)
        return da, dx, None


def mul_bcast(alpha, x, off: float = 0.0):
    return _MulBcast.apply(alpha, x, float(off))


class _GlobalMean(Function):
    """torch.mean(x, (2, 3)) of an NHWC tensor -> fp32 (N, C)  (ChannelAttentionLayer, src/model.py:1426)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        N, H, Wd, Cn = x.shape
        mean = torch.empty(N * Cn, dtype=torch.float32, device=x.device)
        invstd = torch.empty(N * Cn, dtype=torch.float32, device=x.device)
        ws = K.norm_workspace(N, H * Wd, Cn, x.device)
        K.norm_stats(x, N, H * Wd, Cn, 1e-5, ws, mean, invstd, None, None, None, 0.0)
        ctx.meta = (tuple(x.shape), x.dtype)
        return mean.view(N, Cn)

    @staticmethod
    def backward(ctx, dmean):
        shape, dt = ctx.meta
        dx = torch.empty(shape, dtype=dt, device=dmean.device)
        K.chan_bcast(_c(dmean), dx, 1.0 / (shape[1] * shape[2]))
        return dx


def global_mean(x):
    return _GlobalMean.apply(x)


class _ChanScale(Function):
    """(1 + alpha[n, c]) * x  (ChannelAttentionLayer, src/model.py:1431-1432); alpha fp32 (N, C)."""

    @staticmethod
    def forward(ctx, x, alpha):
        x, alpha = _c(x), _c(alpha)
        y = torch.empty_like(x)
        K.chan_scale_fwd(x, alpha, y)
        ctx.save_for_backward(x, alpha)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, alpha = ctx.saved_tensors
        dx, da = torch.empty_like(x), torch.empty_like(alpha)
        K.chan_scale_bwd(x, alpha, _c(dy), dx, da)
        return dx, da


def chan_scale(x, alpha):
    return _ChanScale.apply(x, alpha)


class _FlipAbsDiff(Function):
    """|g - flip(g, H)|  (SymmetryGateResidualSpatialAttentionLayer, src/model.py:1408-1409)."""

    @staticmethod
    def forward(ctx, g):
        g = _c(g)
        out = torch.empty_like(g)
        K.flip_absdiff_fwd(g, out)
        ctx.save_for_backward(g)
        return out

    @staticmethod
    def backward(ctx, dout):
        (g,) = ctx.saved_tensors
        dg = torch.empty_like(g)
        K.flip_absdiff_bwd(g, _c(dout), dg)
        return dg


def flip_absdiff(g):
    return _FlipAbsDiff.apply(g)


class _Add(Function):
    """a + b on the rd_add kernel (the sum of the channel- and spatial-attention skip paths, src/model.py:1119)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        y = torch.empty_like(a)
        K.add(a, b, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _Add.apply(a, b)


class _Fanout(Function):
    """n aliases of x for n consumers; the backward sums their gradients in ONE rd_add_n launch.  (Autograd's own accumulation of a
    multi-consumer tensor's gradient is a chain of at::add launches — eager PyTorch kernels inside the captured iteration.)"""

    @staticmethod
    def forward(ctx, x, n):
        ctx.set_materialize_grads(False)          # an alias nobody consumed contributes None, not a zero tensor to fill and add
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *grads):
        gs = [_c(g) for g in grads if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        out = torch.empty_like(gs[0])
        K.add_n(gs[:8], out)
        rest = gs[8:]
        while rest:                                   # more than 8 consumers: fold the remainder in
            K.add_n([out] + rest[:7], out)
            rest = rest[7:]
        return out, None


def fanout(x, n: int):
    """n aliases of x, one per consumer (see _Fanout); without grad (or n == 1) x itself n times."""
    if n <= 1 or not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * max(n, 1)
    return _Fanout.apply(x, int(n))


class _SplitBlocks(Function):
    """(a, b) = the row blocks `idx_a` / `idx_b` of x (a partition of its blocks: the self- and the cross-reconstructions of the one
    16-decode pass).  Forward: two block gathers.  Backward: ONE pass that scatters both gradients into a tensor whose channels are
    zero-padded to the tensor-core vector (7 -> 16) and returns its first c channels as a strided view — the decoder's last convolution
    recognises the padded buffer behind the view and reads it as its dY without another copy (see _padded_base)."""

    @staticmethod
    def forward(ctx, x, idx_a, idx_b, block):
        x = _c(x)
        outs = []
        for idx in (idx_a, idx_b):
            o = torch.empty((len(idx) * block,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            K.gather_blocks_fwd(x, o, idx, block)
            outs.append(o)
        ctx.meta = (tuple(idx_a), tuple(idx_b), block, tuple(x.shape))
        return tuple(outs)

    @staticmethod
    def backward(ctx, da, db):
        idx_a, idx_b, block, shape = ctx.meta
        nb = shape[0] // block
        sel, sblk = [0] * nb, [0] * nb
        for k, d in enumerate(idx_a):
            sel[d], sblk[d] = 0, k
        for k, d in enumerate(idx_b):
            sel[d], sblk[d] = 1, k
        ref = da if da is not None else db
        c = shape[-1]
        cp = _up8(c) if ref.dtype == torch.bfloat16 else c
        zeros_like = lambda n: K.zeros((n * block,) + tuple(shape[1:]), ref.dtype, ref.device)
        da = _c(da) if da is not None else zeros_like(len(idx_a))
        db = _c(db) if db is not None else zeros_like(len(idx_b))
        out = torch.empty(tuple(shape[:-1]) + (cp,), dtype=ref.dtype, device=ref.device)
        K.scatter_blocks2(da, db, out, sel, sblk, block)
        return (out if cp == c else out[..., :c]), None, None, None


def split_blocks(x, idx_a, idx_b, block):
    idx_a, idx_b = [int(i) for i in idx_a], [int(i) for i in idx_b]
    nb = x.shape[0] // block
    if sorted(idx_a + idx_b) != list(range(nb)) or nb > 32 or _os.environ.get("RD_B200_NO_SPLIT_BLOCKS") is not None:
        return gather_blocks(x, idx_a, block), gather_blocks(x, idx_b, block)          # not a partition: two independent gathers
    return _SplitBlocks.apply(x, tuple(idx_a), tuple(idx_b), int(block))


def _padded_base(dy, c_pad):
    """dy (.., c) that is the leading-channel view of a contiguous (.., c_pad) buffer whose padding is zero (_SplitBlocks.backward):
    return that buffer, else None."""
    if dy.dim() < 2 or dy.shape[-1] >= c_pad or dy.storage_offset() != 0 or dy.stride(-1) != 1:
        return None
    shape = tuple(dy.shape[:-1]) + (c_pad,)
    want, acc = [], 1
    for d in reversed(shape):
        want.append(acc)
        acc *= d
    want = tuple(reversed(want))
    if tuple(dy.stride()) != want or dy.untyped_storage().nbytes() < acc * dy.element_size():
        return None
    return dy.as_strided(shape, want)
