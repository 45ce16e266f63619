"""Tensor-level wrappers over the C ABI (one Python function per rd_* entry point).

These functions only marshal: they take CUDA tensors that the caller allocated (PyTorch owns all
device memory), check contiguity / dtype, and pass raw pointers + sizes + the current CUDA stream to
librd_b200.so.  No arithmetic happens here and nothing falls back to PyTorch ops.
Activations are NHWC-contiguous tensors of shape (N, H, W, C), fp32 or bf16.
"""
import ctypes as C

import torch

from . import lib as _lib
from .lib import ConvDesc, RD_BF16, RD_F32

_TYPES_ARR = C.c_float * 16


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return RD_F32
    if t.dtype == torch.bfloat16:
        return RD_BF16
    raise TypeError("rd_b200: unsupported dtype %s" % t.dtype)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.RdError("rd_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise ValueError("rd_b200: tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def _ctx_stream(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.RdError("rd_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return _lib.get_ctx(idx), C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def _types(types):
    arr = _TYPES_ARR()
    for i, v in enumerate(types):
        arr[i] = float(v)
    return C.cast(arr, C.c_void_p), arr


# ------------------------------------------------------------------------------- layout / cast
def nchw_to_nhwc(src, dst, c0, c):
    n, ct, h, w = src.shape
    ctx, st = _ctx_stream(src)
    _lib.call("rd_nchw_to_nhwc", ctx, _p(src), _p(dst), n, ct, c0, c, h, w, _dt(dst), st)


def stack_modalities(src, dst, mods):
    """src (n, mods * c, h, w) fp32 -> dst (mods * n, h, w, c_pad >= c; the padding zero): every contrast of the batch in one launch."""
    n, ct, h, w = src.shape
    ctx, st = _ctx_stream(src)
    _lib.call("rd_stack_modalities", ctx, _p(src), _p(dst), n, mods, ct // mods, dst.shape[-1], h, w, _dt(dst), st)


def nchw_to_nhwc_strided(src, dst, c_total):
    """src: (N, C, H, W) fp32 view whose images are c_total*H*W apart (a channel slice of a contiguous NCHW
    tensor, or c_total == C for a contiguous one); dst (N, H, W, C)."""
    n, c, h, w = src.shape
    if not src.is_cuda:
        raise _lib.RdError("rd_b200 kernels need CUDA tensors; there is no CPU path")
    ctx, st = _ctx_stream(src)
    _lib.call("rd_nchw_to_nhwc", ctx, C.c_void_p(src.data_ptr()), _p(dst), n, c_total, 0, c, h, w, _dt(dst), st)


def nhwc_to_nchw(src, dst):
    n, h, w, c = src.shape
    ctx, st = _ctx_stream(src)
    _lib.call("rd_nhwc_to_nchw", ctx, _p(src), _p(dst), n, c, h, w, _dt(src), st)


def cast(src, dst):
    ctx, st = _ctx_stream(src)
    _lib.call("rd_cast", ctx, _p(src), _dt(src), _p(dst), _dt(dst), src.numel(), st)


def concat_channels(a, b, out):
    ctx, st = _ctx_stream(a)
    pixels = a.numel() // a.shape[-1]
    _lib.call("rd_concat_channels", ctx, _p(a), _p(b), _p(out), pixels, a.shape[-1], b.shape[-1], _dt(a), st)


def split_channels(inp, a, b, ca, cb):
    ctx, st = _ctx_stream(inp)
    pixels = inp.numel() // inp.shape[-1]
    _lib.call("rd_split_channels", ctx, _p(inp), _p(a), _p(b), pixels, ca, cb, _dt(inp), st)


def add(x, a, y):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_add", ctx, _p(x), _p(a), _p(y), x.numel(), _dt(x), st)


def zeros(shape, dtype, device):
    """A zero tensor whose fill is a memset node (rd_zero), not an at::fill kernel launch."""
    t = torch.empty(shape, dtype=dtype, device=device)
    if t.numel():
        ctx, st = _ctx_stream(t)
        _lib.call("rd_zero", ctx, _p(t), t.numel() * t.element_size(), st)
    return t


def add_n(xs, y):
    """y = sum(xs), 1..8 tensors of y's shape / dtype, one launch."""
    ctx, st = _ctx_stream(y)
    arr = (C.c_void_p * len(xs))(*[_p(t) for t in xs])
    _lib.call("rd_add_n", ctx, C.cast(arr, C.c_void_p), len(xs), _p(y), y.numel(), _dt(y), st)


def scatter_blocks2(a, b, dst, sel, sblk, block):
    """dst block d = block sblk[d] of a (sel[d] == 0) or b (sel[d] == 1), channels zero-padded to dst's."""
    ctx, st = _ctx_stream(dst)
    bp = block
    for d in dst.shape[1:-1]:
        bp *= d
    sa, sb = _idx_array(sel), _idx_array(sblk)
    _lib.call("rd_scatter_blocks2", ctx, _p(a), _p(b), _p(dst), C.cast(sa, C.c_void_p), C.cast(sb, C.c_void_p), len(sel), bp,
              a.shape[-1], dst.shape[-1], _dt(dst), st)


def _idx_array(index):
    return (C.c_int32 * len(index))(*[int(i) for i in index])


def gather_blocks_fwd(src, dst, index, block):
    """dst block k = src block index[k] (block = `block` leading rows); dst may carry zero-padded channels."""
    ctx, st = _ctx_stream(src)
    c, c_pad = src.shape[-1], dst.shape[-1]
    bp = block * (src[0].numel() // c)
    _lib.call("rd_gather_blocks_fwd", ctx, _p(src), _p(dst), _idx_array(index), len(index), bp, c, c_pad, _dt(src), st)


def gather_blocks_bwd(dout, dsrc, index, block):
    ctx, st = _ctx_stream(dout)
    c, c_pad = dsrc.shape[-1], dout.shape[-1]
    bp = block * (dsrc[0].numel() // c)
    _lib.call("rd_gather_blocks_bwd", ctx, _p(dout), _p(dsrc), _idx_array(index), len(index), dsrc.shape[0] // block, bp, c, c_pad,
              _dt(dout), st)


# ------------------------------------------------------------------------------- CondConv mixing
def _wdims(W):
    if W.dim() == 4:
        return (1,) + tuple(W.shape)
    return tuple(W.shape)


def condconv_mix_fwd(W, fc_w, fc_b, types, i_pad, o_total, oT_total, o_off, packed, packedT, r_out):
    """W (E,O,I,kh,kw) or (O,I,kh,kw) fp32; packed [G,o_total,kh*kw,i_pad], packedT [G,i_pad,kh*kw,oT_total]."""
    E, O, I_, kh, kw = _wdims(W)
    ctx, st = _ctx_stream(W)
    tp, keep = _types(types)
    dt = _dt(packed if packed is not None else packedT)
    _lib.call("rd_condconv_mix_fwd", ctx, _p(W), _p(fc_w), _p(fc_b), tp, len(types), E, O, I_, i_pad, kh, kw, o_total,
              oT_total, o_off, _p(packed), _p(packedT), _p(r_out), dt, st)
    del keep


def condconv_mix_bwd(dK, W, fc_w, fc_b, types, i_pad, o_total, o_off, dW, dfc_w, dfc_b):
    E, O, I_, kh, kw = _wdims(W)
    ctx, st = _ctx_stream(W)
    tp, keep = _types(types)
    _lib.call("rd_condconv_mix_bwd", ctx, _p(dK), _p(W), _p(fc_w), _p(fc_b), tp, len(types), E, O, I_, i_pad, kh, kw,
              o_total, o_off, _p(dW), _p(dfc_w), _p(dfc_b), st)
    del keep


class MixBwdBatch:
    """Deferred CondConv mixing backward: heads are queued during backward and mixed in ONE launch per `flush`
    (rd_condconv_mix_bwd_batched).  Every flush of an iteration has its own pinned host table + device table (slot),
    so a captured iteration replays with the tables it was captured with; in eager mode a slot is only rewritten after
    the copy issued from it in the previous iteration has completed."""

    MAX_JOBS = 256

    def __init__(self, device):
        self.device = torch.device(device)
        self.jobs, self.keep, self.slots, self.slot = [], [], [], 0

    def begin_iteration(self):
        self.slot = 0
        self.jobs, self.keep = [], []

    def _slot(self):
        while len(self.slots) <= self.slot:
            nbytes = C.sizeof(_lib.MixJob) * self.MAX_JOBS
            host = torch.empty(nbytes, dtype=torch.uint8)
            if self.device.type == "cuda":
                host = host.pin_memory()
            dev = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            table = (_lib.MixJob * self.MAX_JOBS).from_address(host.data_ptr())
            self.slots.append({"host": host, "dev": dev, "table": table, "event": None})
        sl = self.slots[self.slot]
        self.slot += 1
        return sl

    def add(self, dK, W, fc_w, fc_b, types, i_pad, o_total, o_off, dW, dfc_w, dfc_b, bias_src=None, bias_dst=None):
        """bias_src / bias_dst (optional, contiguous fp32): bias_dst += bias_src in the same launch."""
        E, O, I_, kh, kw = _wdims(W)
        if any(j[8].data_ptr() == dW.data_ptr() for j in self.jobs):
            self.flush()        # a layer that ran twice with gradients (modality encoder: real + cycle pass): its dW += must not race
        self.jobs.append((dK, W, fc_w, fc_b, tuple(float(t) for t in types), i_pad, o_total, o_off, dW, dfc_w, dfc_b, E, O, I_, kh * kw,
                          bias_src, bias_dst))
        self.keep.append(dK)
        if bias_src is not None:
            self.keep.append(bias_src)
        if len(self.jobs) >= self.MAX_JOBS:
            self.flush()

    def flush(self):
        if not self.jobs:
            return
        lib = _lib.load()
        sl = self._slot()
        capturing = self.device.type == "cuda" and torch.cuda.is_current_stream_capturing()
        if sl["event"] is not None and not capturing:
            sl["event"].synchronize()
        table = sl["table"]
        nb = 0
        for j, (dK, W, fc_w, fc_b, types, i_pad, o_total, o_off, dW, dfc_w, dfc_b, E, O, I_, taps, b_src, b_dst) in enumerate(self.jobs):
            t = table[j]
            t.dK, t.W = dK.data_ptr(), W.data_ptr()
            t.fc_w = fc_w.data_ptr() if fc_w is not None else None
            t.fc_b = fc_b.data_ptr() if fc_b is not None else None
            t.dW = dW.data_ptr()
            t.dfc_w = dfc_w.data_ptr() if dfc_w is not None else None
            t.dfc_b = dfc_b.data_ptr() if dfc_b is not None else None
            for g in range(16):
                t.types[g] = types[g] if g < len(types) else 0.0
            t.G, t.E, t.O, t.I, t.i_pad, t.taps, t.o_total, t.o_off = len(types), E, O, I_, i_pad, taps, o_total, o_off
            blocks = int(lib.rd_mix_job_blocks(O, I_, taps))
            t.block_begin, t.blocks = nb, blocks
            t.bias_src = b_src.data_ptr() if b_src is not None else None
            t.bias_dst = b_dst.data_ptr() if b_src is not None else None
            t.bias_n = b_src.numel() if b_src is not None else 0
            nb += blocks
        n = len(self.jobs)
        nbytes = C.sizeof(_lib.MixJob) * n
        sl["dev"][:nbytes].copy_(sl["host"][:nbytes], non_blocking=True)
        if self.device.type == "cuda" and not capturing:
            ev = torch.cuda.Event()
            ev.record()
            sl["event"] = ev
        ctx, st = _ctx_stream(sl["dev"])
        max_g = max(len(j[4]) for j in self.jobs)
        _lib.call("rd_condconv_mix_bwd_batched", ctx, _p(sl["dev"]), n, nb, max_g, st)
        self.jobs = []
        self.keep = []


class MixFwdPlan:
    """Every CondConv expert mixing (+ bias packing) of an iteration in ONE launch (rd_condconv_mix_fwd_batched).

    The experts only change at the optimizer step, so the packed weights of all layers can be produced up front.  The first
    iteration runs the per-head launches and records them (`register`: persistent packed / packedT / bias buffers and the job
    list of the fused launch); from the next iteration on `prepare()` — called by the trainer before the forward pass — fills
    all recorded buffers with one kernel and `_GroupedConv.forward` takes them from `get()`.  Layers that run twice per
    iteration (the cycle re-encoding) are mixed once.  The device job table is static (parameter storage and the buffers
    never move), so a captured iteration replays it as is."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.entries, self.jobs, self.valid = {}, [], set()
        self.dirty, self.dev_table, self.nblocks, self.dtype = False, None, 0, None

    def get(self, key):
        return self.entries.get(key) if key in self.valid else None

    def register(self, key, packed, packedT, bias_all, jobs):
        """jobs: (W, fc_w, fc_b, types, i_pad, o_total, oT_total, o_off, packed view, packedT view, bias src, bias dst)"""
        self.entries[key] = (packed, packedT, bias_all)
        self.jobs.extend(jobs)
        self.valid.add(key)                  # the per-head launches of this iteration have just filled it
        self.dirty = True
        self.dtype = _dt(packed)

    def _build(self):
        lib = _lib.load()
        n = len(self.jobs)
        table = (_lib.MixFJob * n)()
        nb = 0
        for j, (W, fc_w, fc_b, types, i_pad, o_total, oT_total, o_off, pk, pkT, bsrc, bdst) in enumerate(self.jobs):
            E, O, I_, kh, kw = _wdims(W)
            t = table[j]
            t.W = W.data_ptr()
            t.fc_w = fc_w.data_ptr() if fc_w is not None else None
            t.fc_b = fc_b.data_ptr() if fc_b is not None else None
            t.packed, t.packedT = pk.data_ptr(), pkT.data_ptr()
            t.bias_src = bsrc.data_ptr() if bsrc is not None else None
            t.bias_dst = bdst.data_ptr() if bdst is not None else None
            t.bias_n = bsrc.numel() if bsrc is not None else 0
            for g in range(16):
                t.types[g] = float(types[g]) if g < len(types) else 0.0
            t.G, t.E, t.O, t.I, t.i_pad, t.taps = len(types), E, O, I_, i_pad, kh * kw
            t.o_total, t.oT_total, t.o_off = o_total, oT_total, o_off
            blocks = int(lib.rd_mixf_job_blocks(O, i_pad, kh * kw))
            t.block_begin, t.blocks = nb, blocks
            nb += blocks
        raw = bytes(table)
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        self.dev_table = host.to(self.device)          # synchronous upload, outside any capture (see prepare)
        self.nblocks = nb
        self.dirty = False

    def prepare(self):
        """One launch that (re)fills every recorded buffer from the current parameters."""
        self.valid = set()
        if not self.jobs:
            return
        if self.dirty:
            if self.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
                return                                  # new layers appeared right before a capture: this iteration mixes per head
            self._build()
        ctx, st = _ctx_stream(self.dev_table)
        _lib.call("rd_condconv_mix_fwd_batched", ctx, _p(self.dev_table), len(self.jobs), self.nblocks, self.dtype, st)
        self.valid = set(self.entries.keys())


def modality_weights(mask, w):
    """w[i] = [mask[:, i].sum() != 0] / #present contrasts (device side)."""
    B, M = mask.shape
    ctx, st = _ctx_stream(mask)
    _lib.call("rd_modality_weights", ctx, _p(mask), _p(w), B, M, st)


def compose_tail_fwd(pA, pB, bA, bB, modules, packed, packedT, b_eff):
    """pA (G, OA, taps, Cin) fp32, pB (G, OB, OA) fp32, bA (modules, OA) / bB (modules, OB) fp32 or None -> packed (G, OB, taps, Cin),
    packedT (G, Cin, taps, o_pad) in the convolution dtype, b_eff (G, OB) fp32."""
    G, OA, taps, Cin = pA.shape
    OB = pB.shape[1]
    ctx, st = _ctx_stream(pA)
    _lib.call("rd_compose_tail_fwd", ctx, _p(pA), _p(pB), _p(bA), _p(bB), G, modules, OA, OB, taps, Cin, packedT.shape[-1], _dt(packed),
              _p(packed), _p(packedT), _p(b_eff), st)


def compose_tail_bwd(dK, db, pA, pB, bA, modules, dpA, dpB, dbA, dbB):
    """dK (G, o_pad, taps, Cin), db (G, o_pad) fp32 -> dpA, dpB overwritten; dbA (modules, OA) / dbB (modules, OB) accumulated (None = skip)."""
    G, OA, taps, Cin = pA.shape
    OB = pB.shape[1]
    ctx, st = _ctx_stream(pA)
    _lib.call("rd_compose_tail_bwd", ctx, _p(dK), _p(db), _p(pA), _p(pB), _p(bA), G, modules, OA, OB, taps, Cin, dK.shape[1],
              _p(dpA), _p(dpB), _p(dbA), _p(dbB), st)


def pad_channels(inp, out):
    ctx, st = _ctx_stream(inp)
    _lib.call("rd_pad_channels", ctx, _p(inp), _p(out), inp.numel() // inp.shape[-1], inp.shape[-1], out.shape[-1],
              _dt(inp), st)


# ------------------------------------------------------------------------------- convolution
def conv_desc(n, h, w, cin, cout, kh, kw, stride, pad, groups, dtype, act=0, slope=0.2, algo=0, bias_groups=0) -> ConvDesc:
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    return ConvDesc(n, h, w, cin, oh, ow, cout, kh, kw, stride, pad, groups, dtype, act, slope, algo, bias_groups)


def conv2d_fwd(d: ConvDesc, x, packed, bias, y):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_conv2d_fwd", ctx, C.cast(C.byref(d), C.c_void_p), _p(x), _p(packed), _p(bias), _p(y), st)


def conv2d_fwd_spade_supported(d: ConvDesc, x) -> bool:
    """True when the gamma|beta convolution `d` (cout = 2C) can run with the SPADE modulation fused into its epilogue."""
    ctx, _ = _ctx_stream(x)
    return bool(_lib.load().rd_conv2d_fwd_spade_supported(ctx, C.cast(C.byref(d), C.c_void_p)))


def conv2d_fwd_spade(d: ConvDesc, x, packed, bias, z, mean, invstd, gamma, mix):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_conv2d_fwd_spade", ctx, C.cast(C.byref(d), C.c_void_p), _p(x), _p(packed), _p(bias), _p(z), _p(mean), _p(invstd),
              _p(gamma), _p(mix), st)


def wgrad_tma_plan(d: ConvDesc, sm_count: int = 148):
    """rd_wgrad_tma_plan (host only, no GPU needed): the launch split of the TMA weight-gradient kernel for `d`, or None when the shape
    does not run on it.  Keys: xb_total, xb_per_cta, xsplits, tiles_per_group, chunk_tiles, chunks_per_group, ctas, ctas_two_waves."""
    out = (C.c_int * 8)()
    if not _lib.load().rd_wgrad_tma_plan(C.cast(C.byref(d), C.c_void_p), int(sm_count), C.cast(out, C.c_void_p)):
        return None
    keys = ("xb_total", "xb_per_cta", "xsplits", "tiles_per_group", "chunk_tiles", "chunks_per_group", "ctas", "ctas_two_waves")
    return dict(zip(keys, [int(v) for v in out]))


def conv2d_dgrad(d: ConvDesc, dy, packedT, dx):
    ctx, st = _ctx_stream(dy)
    _lib.call("rd_conv2d_dgrad", ctx, C.cast(C.byref(d), C.c_void_p), _p(dy), _p(packedT), _p(dx), st)


def conv2d_wgrad(d: ConvDesc, x, dy, dK, dbias):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_conv2d_wgrad", ctx, C.cast(C.byref(d), C.c_void_p), _p(x), _p(dy), _p(dK), _p(dbias), st)


# ------------------------------------------------------------------------------- normalisation
def norm_workspace(G, ppg, Cn, device):
    chunks = _lib.norm_partial_chunks(ppg, Cn)
    return torch.empty(G * chunks * 2 * Cn + G * 2 * Cn, dtype=torch.float32, device=device)


def spade_bwd_workspace(z):
    """Workspace of spade_modulate_bwd(_g) for z (N, H, W, C): sized for the single-pass kernel's smaller chunks."""
    n, h, w, c = z.shape
    return torch.empty(_lib.spade_bwd_workspace(n, h * w, c, _dt(z)), dtype=torch.float32, device=z.device)


def norm_stats(x, G, ppg, Cn, eps, partial, mean, invstd, running_mean=None, running_var=None, nbt=None, momentum=0.1):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_norm_stats", ctx, _p(x), G, ppg, Cn, _dt(x), eps, _p(partial), _p(mean), _p(invstd),
              _p(running_mean), _p(running_var), _p(nbt), momentum, st)


def norm_eval_stats(running_mean, running_var, G, eps, mean, invstd):
    ctx, st = _ctx_stream(running_mean)
    _lib.call("rd_norm_eval_stats", ctx, _p(running_mean), _p(running_var), G, running_mean.numel(), eps, _p(mean),
              _p(invstd), st)


def norm_apply(x, mean, invstd, weight, bias, y, G, ppg, Cn):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_norm_apply", ctx, _p(x), _p(mean), _p(invstd), _p(weight), _p(bias), _p(y), G, ppg, Cn, _dt(x), st)


def norm_bwd(x, dy, mean, invstd, weight, dx, dweight, dbias, partial, G, ppg, Cn):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_norm_bwd", ctx, _p(x), _p(dy), _p(mean), _p(invstd), _p(weight), _p(dx), _p(dweight), _p(dbias),
              _p(partial), G, ppg, Cn, _dt(x), st)


def spade_modulate_fwd(z, mean, invstd, gb, mix):
    n, h, w, c = z.shape
    ctx, st = _ctx_stream(z)
    _lib.call("rd_spade_modulate_fwd", ctx, _p(z), _p(mean), _p(invstd), _p(gb), _p(mix), n, h * w, c, _dt(z), st)


def spade_modulate_bwd(z, mean, invstd, gb, dmix, dz, dgb, partial):
    n, h, w, c = z.shape
    ctx, st = _ctx_stream(z)
    _lib.call("rd_spade_modulate_bwd", ctx, _p(z), _p(mean), _p(invstd), _p(gb), _p(dmix), _p(dz), _p(dgb), _p(partial),
              n, h * w, c, _dt(z), st)


def spade_modulate_bwd_g(z, mean, invstd, gamma, dmix, dz, dgb, partial):
    """spade_modulate_bwd with gamma as its own (N, H, W, C) tensor (the one conv2d_fwd_spade saved)."""
    n, h, w, c = z.shape
    ctx, st = _ctx_stream(z)
    _lib.call("rd_spade_modulate_bwd_g", ctx, _p(z), _p(mean), _p(invstd), _p(gamma), _p(dmix), _p(dz), _p(dgb), _p(partial),
              n, h * w, c, _dt(z), st)


# ------------------------------------------------------------------------------- resize / activations
def bilinear_fwd(x, y, align):
    n, h, w, c = x.shape
    ctx, st = _ctx_stream(x)
    _lib.call("rd_bilinear_fwd", ctx, _p(x), _p(y), n, h, w, c, y.shape[1], y.shape[2], int(align), _dt(x), st)


def bilinear_bwd(dy, dx, align):
    n, h, w, c = dx.shape
    ctx, st = _ctx_stream(dy)
    _lib.call("rd_bilinear_bwd", ctx, _p(dy), _p(dx), n, h, w, c, dy.shape[1], dy.shape[2], int(align), _dt(dy), st)


def lrelu_fwd(x, y, slope):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_lrelu_fwd", ctx, _p(x), _p(y), x.numel(), slope, _dt(x), st)


def lrelu_bwd(dy, y, dx, slope):
    ctx, st = _ctx_stream(dy)
    _lib.call("rd_lrelu_bwd", ctx, _p(dy), _p(y), _p(dx), dy.numel(), slope, _dt(dy), st)


def masked_softmax_fwd(s, mask_img, p):
    """mask_img (B, H, W) fp32 is broadcast over the leading stack: mask index = pixel % (B*H*W)."""
    ctx, st = _ctx_stream(s)
    mp = mask_img.numel() if mask_img is not None else 0
    _lib.call("rd_masked_softmax_fwd", ctx, _p(s), _p(mask_img), mp, _p(p), s.numel() // s.shape[-1], s.shape[-1], _dt(s), st)


def add_relu_fwd(a, b, y):
    ctx, st = _ctx_stream(a)
    _lib.call("rd_add_relu_fwd", ctx, _p(a), _p(b), _p(y), a.numel(), _dt(a), st)


def relu_bwd(dy, y, dx):
    ctx, st = _ctx_stream(dy)
    _lib.call("rd_relu_bwd", ctx, _p(dy), _p(y), _p(dx), dy.numel(), _dt(dy), st)


def sigmoid_fwd(x, y):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_sigmoid_fwd", ctx, _p(x), _p(y), x.numel(), _dt(x), st)


def sigmoid_bwd(dy, y, dx):
    ctx, st = _ctx_stream(dy)
    _lib.call("rd_sigmoid_bwd", ctx, _p(dy), _p(y), _p(dx), dy.numel(), _dt(dy), st)


def mul_bcast_fwd(alpha, x, y, off=0.0):
    """y = (off + alpha[p]) * x[p, c]."""
    ctx, st = _ctx_stream(x)
    _lib.call("rd_mul_bcast_fwd", ctx, _p(alpha), _p(x), _p(y), x.numel() // x.shape[-1], x.shape[-1], float(off), _dt(x), st)


def mul_bcast_bwd(alpha, x, dy, dx, dalpha, off=0.0):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_mul_bcast_bwd", ctx, _p(alpha), _p(x), _p(dy), _p(dx), _p(dalpha), x.numel() // x.shape[-1], x.shape[-1],
              float(off), _dt(x), st)


def chan_scale_fwd(x, a, y):
    """y[n, p, c] = (1 + a[n, c]) * x; a fp32 (N, C)."""
    n, h, w, c = x.shape
    ctx, st = _ctx_stream(x)
    _lib.call("rd_chan_scale_fwd", ctx, _p(x), _p(a), _p(y), n, h * w, c, _dt(x), st)


def chan_scale_bwd(x, a, dy, dx, da):
    n, h, w, c = x.shape
    ctx, st = _ctx_stream(x)
    _lib.call("rd_chan_scale_bwd", ctx, _p(x), _p(a), _p(dy), _p(dx), _p(da), n, h * w, c, _dt(x), st)


def chan_bcast(v, dx, scale):
    """dx[n, p, c] = v[n, c] * scale."""
    n, h, w, c = dx.shape
    ctx, st = _ctx_stream(dx)
    _lib.call("rd_chan_bcast", ctx, _p(v), _p(dx), n, h * w, c, float(scale), _dt(dx), st)


def flip_absdiff_fwd(g, out):
    n, h, w, c = g.shape
    ctx, st = _ctx_stream(g)
    _lib.call("rd_flip_absdiff_fwd", ctx, _p(g), _p(out), n, h, w, c, _dt(g), st)


def flip_absdiff_bwd(g, dout, dg):
    n, h, w, c = g.shape
    ctx, st = _ctx_stream(g)
    _lib.call("rd_flip_absdiff_bwd", ctx, _p(g), _p(dout), _p(dg), n, h, w, c, _dt(g), st)


def masked_softmax_bwd(p, dp, ds):
    ctx, st = _ctx_stream(p)
    _lib.call("rd_masked_softmax_bwd", ctx, _p(p), _p(dp), _p(ds), p.numel() // p.shape[-1], p.shape[-1], _dt(p), st)


# ------------------------------------------------------------------------------- small dense
def linear_fwd(x, W, b, y, act=0, slope=0.2):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_linear_fwd", ctx, _p(x), _p(W), _p(b), _p(y), x.shape[0], W.shape[1], W.shape[0], act, slope, st)


def linear_bwd(x, W, dy, dx, dW, db):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_linear_bwd", ctx, _p(x), _p(W), _p(dy), _p(dx), _p(dW), _p(db), x.shape[0], W.shape[1], W.shape[0], st)


def sample_fwd(mu, lv, eps, z):
    ctx, st = _ctx_stream(mu)
    _lib.call("rd_sample_fwd", ctx, _p(mu), _p(lv), _p(eps), _p(z), mu.numel(), st)


def sample_bwd(dz, lv, eps, dmu, dlv):
    ctx, st = _ctx_stream(dz)
    _lib.call("rd_sample_bwd", ctx, _p(dz), _p(lv), _p(eps), _p(dmu), _p(dlv), dz.numel(), st)


# ------------------------------------------------------------------------------- fusion gather
def fuse_gather_fwd(si, mask, out, idx_out, count_out, B, M):
    ctx, st = _ctx_stream(si)
    row = si.numel() // (B * M)
    _lib.call("rd_fuse_gather_fwd", ctx, _p(si), _p(mask), _p(out), _p(idx_out), _p(count_out), B, M, row, _dt(si), st)


def fuse_gather_bwd(dout, mask, dsi, B, M):
    ctx, st = _ctx_stream(dout)
    row = dsi.numel() // (B * M)
    _lib.call("rd_fuse_gather_bwd", ctx, _p(dout), _p(mask), _p(dsi), B, M, row, _dt(dout), st)


# ------------------------------------------------------------------------------- losses
def recon_rows_fwd(x, gt, gt_index, row_loss, partial, R, p):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_recon_rows_fwd", ctx, _p(x), _p(gt), _dt(gt), _p(gt_index), _p(row_loss), _p(partial), R,
              x.numel() // R, p, _dt(x), st)


def recon_rows_bwd(x, gt, gt_index, coef, dx, R, p):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_recon_rows_bwd", ctx, _p(x), _p(gt), _dt(gt), _p(gt_index), _p(coef), _p(dx), R, x.numel() // R, p,
              _dt(x), st)


def recon_chunks(row_elems):
    return (row_elems + 8191) // 8192


def xmix_plan(mask, gt_index, B, M):
    ctx, st = _ctx_stream(mask)
    _lib.call("rd_xmix_plan", ctx, _p(mask), _p(gt_index), B, M, st)


def masked_combine(row_loss, mask, loss, coef, B, M, kind):
    ctx, st = _ctx_stream(mask)
    _lib.call("rd_masked_combine", ctx, _p(row_loss), _p(mask), _p(loss), _p(coef), B, M, kind, st)


def latent_z_loss(mu, mu_new, mask, loss, dmu, dmu_new, B, M, Z):
    ctx, st = _ctx_stream(mu)
    _lib.call("rd_latent_z_loss", ctx, _p(mu), _p(mu_new), _p(mask), _p(loss), _p(dmu), _p(dmu_new), B, M, Z, st)


def sim_z_loss(z, mask, margin, loss, dz, B, M, Z):
    ctx, st = _ctx_stream(z)
    _lib.call("rd_sim_z_loss", ctx, _p(z), _p(mask), margin, _p(loss), _p(dz), B, M, Z, st)


def kl_loss(mu, lv, mask, loss, dmu, dlv, B, M, Z):
    ctx, st = _ctx_stream(mu)
    _lib.call("rd_kl_loss", ctx, _p(mu), _p(lv), _p(mask), _p(loss), _p(dmu), _p(dlv), B, M, Z, st)


def maxpool16_fwd(s, pooled, argmax):
    n, h, w, c = s.shape
    ctx, st = _ctx_stream(s)
    _lib.call("rd_maxpool16_fwd", ctx, _p(s), _p(pooled), _p(argmax), n, h, w, c, _dt(s), st)


def maxpool16_bwd(dpooled, argmax, ds):
    n, h, w, c = ds.shape
    ctx, st = _ctx_stream(ds)
    _lib.call("rd_maxpool16_bwd", ctx, _p(dpooled), _p(argmax), _p(ds), n, h, w, c, _dt(ds), st)


def avgpool16_fwd(s, pooled):
    n, h, w, c = s.shape
    ctx, st = _ctx_stream(s)
    _lib.call("rd_avgpool16_fwd", ctx, _p(s), _p(pooled), n, h, w, c, _dt(s), st)


def avgpool16_bwd(dpooled, ds):
    n, h, w, c = ds.shape
    ctx, st = _ctx_stream(ds)
    _lib.call("rd_avgpool16_bwd", ctx, _p(dpooled), _p(ds), n, h, w, c, _dt(ds), st)


def softplus_fwd(x, y):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_softplus_fwd", ctx, _p(x), _p(y), x.numel(), _dt(x), st)


def softplus_bwd(dy, x, dx):
    ctx, st = _ctx_stream(x)
    _lib.call("rd_softplus_bwd", ctx, _p(dy), _p(x), _p(dx), x.numel(), _dt(x), st)


def sim_s_loss(pooled, mask, pair, margin, loss, dpooled, B, M, D):
    ctx, st = _ctx_stream(pooled)
    _lib.call("rd_sim_s_loss", ctx, _p(pooled), _p(mask), _p(pair), margin, _p(loss), _p(dpooled), B, M, D, st)


def seg_loss_fwd(y, target, loss, partial):
    n, h, w, c = y.shape
    ctx, st = _ctx_stream(y)
    _lib.call("rd_seg_loss_fwd", ctx, _p(y), _p(target), _p(loss), _p(partial), n, h * w, _dt(y), st)


def seg_loss_bwd(y, target, partial, upstream, dy):
    n, h, w, c = y.shape
    ctx, st = _ctx_stream(y)
    _lib.call("rd_seg_loss_bwd", ctx, _p(y), _p(target), _p(partial), _p(upstream), _p(dy), n, h * w, _dt(y), st)


SEG_PARTIAL_FLOATS = 256 * 11 + 11


# ------------------------------------------------------------------------------- optimizer
def grad_norm(grad, segments, nseg, partial, scalars, max_norm):
    ctx, st = _ctx_stream(grad)
    _lib.call("rd_grad_norm", ctx, _p(grad), _p(segments), nseg, _p(partial), _p(scalars), max_norm, st)


def grad_scale(grad, segments, nseg, scalars):
    ctx, st = _ctx_stream(grad)
    _lib.call("rd_grad_scale", ctx, _p(grad), _p(segments), nseg, _p(scalars), st)


def adam_amsgrad(param, grad, m, v, vmax, segments, nseg, hyper):
    ctx, st = _ctx_stream(param)
    _lib.call("rd_adam_amsgrad", ctx, _p(param), _p(grad), _p(m), _p(v), _p(vmax), _p(segments), nseg, _p(hyper), st)


def clip_adam_amsgrad(param, grad, m, v, vmax, segments, nseg, hyper, scalars, zero_grad=True):
    """grad_scale + adam_amsgrad + zero_grad of the active segments in one pass (main_missing.py:272-284)."""
    ctx, st = _ctx_stream(param)
    _lib.call("rd_clip_adam_amsgrad", ctx, _p(param), _p(grad), _p(m), _p(v), _p(vmax), _p(segments), nseg, _p(hyper),
              _p(scalars) if scalars is not None else None, 1 if zero_grad else 0, st)


def clip_adam_amsgrad_gated(param, grad, m, v, vmax, segments, seg_param, nseg, partial, param_flags, param_steps, hyper, scalars,
                            zero_grad=True):
    """clip_adam_amsgrad with torch's per-parameter "grad is None -> skip" rule (a parameter whose gradient segments are all exactly
    zero is not touched) and per-parameter step counters; `partial` = the per-segment squared sums of the grad_norm call before it."""
    ctx, st = _ctx_stream(param)
    _lib.call("rd_clip_adam_amsgrad_gated", ctx, _p(param), _p(grad), _p(m), _p(v), _p(vmax), _p(segments), _p(seg_param), nseg,
              _p(partial), _p(param_flags), _p(param_steps), param_steps.numel(), _p(hyper),
              _p(scalars) if scalars is not None else None, 1 if zero_grad else 0, st)


# ------------------------------------------------------------------------------- evaluation metrics / slab assembly
def metrics_recon(target, pred, out, t_index=None, t_c0=0, p_c0=0):
    """compute_reconstruction_metrics (src/util.py:935-978) of pred[n, :, :, p_c0] against target[t_index[n] or n, :, :, t_c0]
    (NHWC tensors); out (N, 3) fp32 = ssim, psnr, mse per image."""
    n, h, w, cp = pred.shape
    ct = target.shape[-1]
    ctx, st = _ctx_stream(pred)
    tiles = int(_lib.load().rd_metrics_recon_tiles(h, w))
    stats = torch.empty(n * 3, dtype=torch.float32, device=pred.device)
    partial = torch.empty(n * tiles * 2, dtype=torch.float64, device=pred.device)
    _lib.call("rd_metrics_recon", ctx, _p(target), _dt(target), ct, t_c0, _p(t_index), _p(pred), _dt(pred), cp, p_c0, n, h, w,
              _p(stats), _p(partial), _p(out), st)


def metrics_seg(target, pred, out):
    """compute_segmentation_metrics (src/util.py:946-954, 980-992): target (N, H*W) fp32 labels, pred NHWC logits; out (N, 2) = dice, iou."""
    n = pred.shape[0]
    cp = pred.shape[-1]
    ctx, st = _ctx_stream(pred)
    _lib.call("rd_metrics_seg", ctx, _p(target), _p(pred), _dt(pred), cp, n, pred.numel() // (n * cp), _p(out), st)


def assemble_slabs(vols, present, tvols, has_target, brain_mask, subj, slice_idx, drop, inputs, targets, mask, mask_img, block,
                   remap4, clamp_hi):
    """ZeroDoseDataset.__getitem__ (src/util.py:471-566) for a batch from device-resident volumes; see include/rd_b200.h."""
    S, M, D, H, W = vols.shape
    B = inputs.shape[0]
    ctx, st = _ctx_stream(vols)
    _lib.call("rd_assemble_slabs", ctx, _p(vols), _p(present), _p(tvols), _p(has_target), _p(brain_mask), _p(subj), _p(slice_idx),
              _p(drop), _p(inputs), _p(targets), _p(mask), _p(mask_img), B, M, block, D, H, W, 1 if remap4 else 0, clamp_hi, st)


# ------------------------------------------------------------------------------- runtime services (rd_runtime.cu)
class AbiGraph:
    """A CUDA graph captured and replayed through the C ABI (rd_graph_*): the non-Python counterpart of the torch.cuda.CUDAGraph
    the trainer uses.  Capture everything launched on the current stream inside `with g.capture():`."""

    def __init__(self, device_index: int = None):
        self.idx = torch.cuda.current_device() if device_index is None else device_index
        self.ctx = _lib.get_ctx(self.idx)
        self.handle = C.c_void_p()

    def capture(self):
        import contextlib

        @contextlib.contextmanager
        def cm():
            st = C.c_void_p(torch.cuda.current_stream(self.idx).cuda_stream)
            _lib.call("rd_graph_begin", self.ctx, st)
            try:
                yield
            finally:
                _lib.call("rd_graph_end", self.ctx, st, C.cast(C.byref(self.handle), C.c_void_p))
        return cm()

    def launch(self):
        _lib.call("rd_graph_launch", self.ctx, self.handle, C.c_void_p(torch.cuda.current_stream(self.idx).cuda_stream))

    def node_count(self):
        k, t = C.c_int64(0), C.c_int64(0)
        _lib.call("rd_graph_node_count", self.ctx, self.handle, C.cast(C.byref(k), C.c_void_p), C.cast(C.byref(t), C.c_void_p))
        return int(k.value), int(t.value)

    def destroy(self):
        if self.handle:
            _lib.call("rd_graph_destroy", self.ctx, self.handle)
            self.handle = C.c_void_p()


def ddp_available(device_index: int = 0):
    v = C.c_int(0)
    ok = _lib.load().rd_ddp_available(_lib.get_ctx(device_index), C.cast(C.byref(v), C.c_void_p))
    return bool(ok), int(v.value)


def ddp_unique_id(device_index: int = 0) -> bytes:
    buf = C.create_string_buffer(128)
    _lib.call("rd_ddp_unique_id", _lib.get_ctx(device_index), C.cast(buf, C.c_void_p))
    return buf.raw


def ddp_init(world: int, rank: int, uid: bytes, device_index: int = 0):
    buf = C.create_string_buffer(bytes(uid), 128)
    _lib.call("rd_ddp_init", _lib.get_ctx(device_index), int(world), int(rank), C.cast(buf, C.c_void_p))


def ddp_bucket_allreduce(grad: torch.Tensor, average: bool = True):
    """In-place NCCL all-reduce (average or sum) of a contiguous fp32 range of the flat gradient buffer, on the current stream."""
    if grad.dtype != torch.float32:
        raise TypeError("ddp_bucket_allreduce: fp32 gradients")
    ctx, st = _ctx_stream(grad)
    _lib.call("rd_ddp_bucket_allreduce", ctx, _p(grad), grad.numel(), 1 if average else 0, st)


def ddp_broadcast(t: torch.Tensor, root: int = 0):
    ctx, st = _ctx_stream(t)
    _lib.call("rd_ddp_broadcast", ctx, _p(t), t.numel() * t.element_size(), int(root), st)


def ddp_finalize(device_index: int = 0):
    _lib.call("rd_ddp_finalize", _lib.get_ctx(device_index))
