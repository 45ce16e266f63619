"""ctypes binding of the C ABI declared in include/rd_b200.h.

The shared library `librd_b200.so` is built in-tree by `csrc/build.sh` (nvcc, sm_100a).  There is
no CPU or PyTorch fallback: if the library is missing, or no sm_100 device is present when a
context is requested, the import / call raises (the product path must fail loudly).
"""
import ctypes as C
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RD_B200_LIB_PATH") or os.path.join(_HERE, "librd_b200.so")      # the override is for A/B timing of two builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rd_b200.h")

RD_F32, RD_BF16 = 0, 1
RD_ALGO_AUTO, RD_ALGO_DIRECT, RD_ALGO_TCGEN05, RD_ALGO_HALO = 0, 1, 2, 3
RD_ACT_NONE, RD_ACT_LRELU = 0, 1


class RdError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int),
                ("oh", C.c_int), ("ow", C.c_int), ("cout", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("groups", C.c_int), ("dtype", C.c_int), ("act", C.c_int), ("act_slope", C.c_float),
                ("algo", C.c_int), ("bias_groups", C.c_int)]


class MixJob(C.Structure):
    """include/rd_b200.h: rd_mix_job"""
    _fields_ = [("dK", C.c_void_p), ("W", C.c_void_p), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p),
                ("dW", C.c_void_p), ("dfc_w", C.c_void_p), ("dfc_b", C.c_void_p),
                ("types", C.c_float * 16),
                ("G", C.c_int32), ("E", C.c_int32), ("O", C.c_int32), ("I", C.c_int32), ("i_pad", C.c_int32),
                ("taps", C.c_int32), ("o_total", C.c_int32), ("o_off", C.c_int32),
                ("block_begin", C.c_int32), ("blocks", C.c_int32),
                ("bias_src", C.c_void_p), ("bias_dst", C.c_void_p), ("bias_n", C.c_int32), ("_pad", C.c_int32)]


class MixFJob(C.Structure):
    """include/rd_b200.h: rd_mixf_job"""
    _fields_ = [("W", C.c_void_p), ("fc_w", C.c_void_p), ("fc_b", C.c_void_p),
                ("packed", C.c_void_p), ("packedT", C.c_void_p), ("bias_src", C.c_void_p), ("bias_dst", C.c_void_p),
                ("types", C.c_float * 16),
                ("G", C.c_int32), ("E", C.c_int32), ("O", C.c_int32), ("I", C.c_int32), ("i_pad", C.c_int32),
                ("taps", C.c_int32), ("o_total", C.c_int32), ("oT_total", C.c_int32), ("o_off", C.c_int32), ("bias_n", C.c_int32),
                ("block_begin", C.c_int32), ("blocks", C.c_int32)]


_lib = None
_lock = threading.Lock()
_ctx = {}

P, I, L, F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes after the leading rd_ctx* (restype is always int unless listed in _SPECIAL)
_SIGS = {
    "rd_nchw_to_nhwc": [P, P, I, I, I, I, I, I, I, P],
    "rd_stack_modalities": [P, P, I, I, I, I, I, I, I, P],
    "rd_nhwc_to_nchw": [P, P, I, I, I, I, I, P],
    "rd_cast": [P, I, P, I, L, P],
    "rd_concat_channels": [P, P, P, L, I, I, I, P],
    "rd_split_channels": [P, P, P, L, I, I, I, P],
    "rd_add": [P, P, P, L, I, P],
    "rd_add_n": [P, I, P, L, I, P],
    "rd_gather_blocks_fwd": [P, P, P, I, L, I, I, I, P],
    "rd_scatter_blocks2": [P, P, P, P, P, I, L, I, I, I, P],
    "rd_gather_blocks_bwd": [P, P, P, I, I, L, I, I, I, P],
    "rd_condconv_mix_fwd": [P, P, P, P, I, I, I, I, I, I, I, I, I, I, P, P, P, I, P],
    "rd_condconv_mix_bwd": [P, P, P, P, P, I, I, I, I, I, I, I, I, I, P, P, P, P],
    "rd_pad_channels": [P, P, L, I, I, I, P],
    "rd_condconv_mix_bwd_batched": [P, I, I, I, P],
    "rd_zero": [P, L, P],
    "rd_graph_begin": [P],
    "rd_graph_end": [P, P],
    "rd_graph_launch": [P, P],
    "rd_graph_node_count": [P, P, P],
    "rd_graph_destroy": [P],
    "rd_ddp_available": [P],
    "rd_ddp_unique_id": [P],
    "rd_ddp_init": [I, I, P],
    "rd_ddp_bucket_allreduce": [P, L, I, P],
    "rd_ddp_broadcast": [P, L, I, P],
    "rd_ddp_finalize": [],
    "rd_modality_weights": [P, P, I, I, P],
    "rd_compose_tail_fwd": [P, P, P, P, I, I, I, I, I, I, I, I, P, P, P, P],
    "rd_compose_tail_bwd": [P, P, P, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    "rd_condconv_mix_fwd_batched": [P, I, I, I, P],
    "rd_conv2d_fwd": [P, P, P, P, P, P],
    "rd_conv2d_fwd_spade_supported": [P],
    "rd_conv2d_fwd_spade": [P, P, P, P, P, P, P, P, P, P],
    "rd_conv2d_dgrad": [P, P, P, P, P],
    "rd_conv2d_wgrad": [P, P, P, P, P, P],
    "rd_norm_stats": [P, I, L, I, I, F, P, P, P, P, P, P, F, P],
    "rd_norm_eval_stats": [P, P, I, I, F, P, P, P],
    "rd_norm_apply": [P, P, P, P, P, P, I, L, I, I, P],
    "rd_norm_bwd": [P, P, P, P, P, P, P, P, P, I, L, I, I, P],
    "rd_spade_modulate_fwd": [P, P, P, P, P, I, L, I, I, P],
    "rd_spade_modulate_bwd": [P, P, P, P, P, P, P, P, I, L, I, I, P],
    "rd_spade_modulate_bwd_g": [P, P, P, P, P, P, P, P, I, L, I, I, P],
    "rd_bilinear_fwd": [P, P, I, I, I, I, I, I, I, I, P],
    "rd_bilinear_bwd": [P, P, I, I, I, I, I, I, I, I, P],
    "rd_lrelu_fwd": [P, P, L, F, I, P],
    "rd_lrelu_bwd": [P, P, P, L, F, I, P],
    "rd_masked_softmax_fwd": [P, P, L, P, L, I, I, P],
    "rd_softplus_fwd": [P, P, L, I, P],
    "rd_softplus_bwd": [P, P, P, L, I, P],
    "rd_avgpool16_fwd": [P, P, I, I, I, I, I, P],
    "rd_avgpool16_bwd": [P, P, I, I, I, I, I, P],
    "rd_add_relu_fwd": [P, P, P, L, I, P],
    "rd_relu_bwd": [P, P, P, L, I, P],
    "rd_sigmoid_fwd": [P, P, L, I, P],
    "rd_sigmoid_bwd": [P, P, P, L, I, P],
    "rd_mul_bcast_fwd": [P, P, P, L, I, F, I, P],
    "rd_mul_bcast_bwd": [P, P, P, P, P, L, I, F, I, P],
    "rd_chan_scale_fwd": [P, P, P, I, L, I, I, P],
    "rd_chan_scale_bwd": [P, P, P, P, P, I, L, I, I, P],
    "rd_chan_bcast": [P, P, I, L, I, F, I, P],
    "rd_flip_absdiff_fwd": [P, P, I, I, I, I, I, P],
    "rd_flip_absdiff_bwd": [P, P, P, I, I, I, I, I, P],
    "rd_masked_softmax_bwd": [P, P, P, L, I, I, P],
    "rd_linear_fwd": [P, P, P, P, I, I, I, I, F, P],
    "rd_linear_bwd": [P, P, P, P, P, P, I, I, I, P],
    "rd_sample_fwd": [P, P, P, P, L, P],
    "rd_sample_bwd": [P, P, P, P, P, L, P],
    "rd_fuse_gather_fwd": [P, P, P, P, P, I, I, L, I, P],
    "rd_fuse_gather_bwd": [P, P, P, I, I, L, I, P],
    "rd_recon_rows_fwd": [P, P, I, P, P, P, I, L, I, I, P],
    "rd_recon_rows_bwd": [P, P, I, P, P, P, I, L, I, I, P],
    "rd_xmix_plan": [P, P, I, I, P],
    "rd_masked_combine": [P, P, P, P, I, I, I, P],
    "rd_latent_z_loss": [P, P, P, P, P, P, I, I, I, P],
    "rd_sim_z_loss": [P, P, F, P, P, I, I, I, P],
    "rd_kl_loss": [P, P, P, P, P, P, I, I, I, P],
    "rd_maxpool16_fwd": [P, P, P, I, I, I, I, I, P],
    "rd_maxpool16_bwd": [P, P, P, I, I, I, I, I, P],
    "rd_sim_s_loss": [P, P, P, F, P, P, I, I, I, P],
    "rd_seg_loss_fwd": [P, P, P, P, I, L, I, P],
    "rd_seg_loss_bwd": [P, P, P, P, P, I, L, I, P],
    "rd_grad_norm": [P, P, I, P, P, F, P],
    "rd_grad_scale": [P, P, I, P, P],
    "rd_adam_amsgrad": [P, P, P, P, P, P, I, P, P],
    "rd_clip_adam_amsgrad": [P, P, P, P, P, P, I, P, P, I, P],
    "rd_metrics_recon": [P, I, I, I, P, P, I, I, I, I, I, I, P, P, P, P],
    "rd_metrics_seg": [P, P, I, I, I, L, P, P],
    "rd_assemble_slabs": [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, P],
    "rd_clip_adam_amsgrad_gated": [P, P, P, P, P, P, P, I, P, P, P, I, P, P, I, P],
}


def header_symbols():
    """Every function name declared in include/rd_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load librd_b200.so (once).  Raises RdError when it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RdError("rd_b200: %s not found — build it with representation-disentanglement_b200/csrc/build.sh "
                          "(or python -c 'import __graft_entry__ as g; g.build()').  There is no fallback path." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.rd_abi_version.restype = I
        lib.rd_ctx_create.argtypes = [C.POINTER(P), I]
        lib.rd_ctx_create.restype = I
        lib.rd_ctx_destroy.argtypes = [P]
        lib.rd_last_error.argtypes = [P]
        lib.rd_last_error.restype = C.c_char_p
        lib.rd_launch_count.argtypes = [P]
        lib.rd_launch_count.restype = L
        lib.rd_last_conv_algo.argtypes = [P]
        lib.rd_last_conv_algo.restype = I
        lib.rd_mix_job_blocks.argtypes = [I, I, I]
        lib.rd_mix_job_blocks.restype = I
        lib.rd_mixf_job_blocks.argtypes = [I, I, I]
        lib.rd_mixf_job_blocks.restype = I
        lib.rd_metrics_recon_tiles.argtypes = [I, I]
        lib.rd_metrics_recon_tiles.restype = I
        lib.rd_norm_partial_chunks.argtypes = [L, I]
        lib.rd_norm_partial_chunks.restype = I
        lib.rd_spade_bwd_workspace.argtypes = [I, L, I, I]
        lib.rd_spade_bwd_workspace.restype = L
        lib.rd_wgrad_tma_plan.argtypes = [P, I, P]
        lib.rd_wgrad_tma_plan.restype = I
        for name, sig in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = [P] + sig
            fn.restype = I
        _lib = lib
        return lib


def get_ctx(device_index: int):
    """Per-device rd_ctx (created on first use).  Raises without an sm_100 CUDA device."""
    lib = load()
    with _lock:
        if device_index in _ctx:
            return _ctx[device_index]
        h = P()
        rc = lib.rd_ctx_create(C.byref(h), int(device_index))
        if rc != 0:
            msg = lib.rd_last_error(h).decode() if h else "no CUDA device"
            raise RdError("rd_ctx_create(device=%d) failed (%d): %s — rd_b200 runs on B200 (sm_100a) only, "
                          "there is no CPU fallback" % (device_index, rc, msg))
        _ctx[device_index] = h
        return h


def call(name: str, ctx, *args):
    lib = load()
    rc = getattr(lib, name)(ctx, *args)
    if rc != 0:
        raise RdError("%s failed (%d): %s" % (name, rc, lib.rd_last_error(ctx).decode()))


def launch_count(device_index: int = 0) -> int:
    return int(load().rd_launch_count(get_ctx(device_index)))


def last_conv_algo(device_index: int = 0) -> int:
    return int(load().rd_last_conv_algo(get_ctx(device_index)))


def spade_bwd_workspace(n: int, hw: int, C: int, dtype: int) -> int:
    """Floats of workspace rd_spade_modulate_bwd(_g) needs (host function, no GPU)."""
    return int(load().rd_spade_bwd_workspace(int(n), int(hw), int(C), int(dtype)))


def norm_partial_chunks(ppg: int, C: int) -> int:
    """Number of partial-sum chunks rd_norm_stats / rd_norm_bwd / rd_spade_modulate_bwd use (host function, no GPU)."""
    return int(load().rd_norm_partial_chunks(int(ppg), int(C)))
