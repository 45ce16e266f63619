"""Data-parallel gradient averaging: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Semantics (SURVEY.md §8e — the reference has no DP, so this is "what DDP would do to the reference"):
rank r runs the reference step on its local shard (local BatchNorm statistics, local masked means, local
roll-by-one negatives); parameter gradients are AVERAGED over ranks before clip + Adam; BatchNorm running
buffers stay rank-local.  Parameters that receive no gradient are excluded from the buckets
statically (the segment table of trainer.FlatParams), so every rank reduces the same byte ranges.

The flat fp32 gradient buffer is reduced in a few contiguous buckets (sum, then one scale kernel by 1/N).
Overlap with backward: gradients bypass autograd's AccumulateGrad (the kernels add straight into the flat
buffer), so readiness is signalled by a MARKER node in the tape instead of per-parameter hooks: the trainer
passes the decoder inputs (S, z) through `ready_marker`; autograd runs the marker's backward exactly when
every decode kernel has finished, i.e. when the gradients of `input_decoder_list` (half of all gradient
bytes) are final.  Their buckets are all-reduced asynchronously on NCCL's stream from that callback while the
encoder backward (the other half of the step's backward time) still runs; `finish` reduces the remaining
buckets, waits for all of them and scales by 1/N.  The NCCL calls are stream-ordered and CUDA-graph
capturable, so with graphs enabled the whole iteration — collectives included — is ONE captured graph.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

from . import kernels as K


def plan_buckets(segments: List[Tuple[int, int]], bucket_elems: int) -> List[Tuple[int, int]]:
    """Merge the active (offset, length) segments into contiguous [start, end) ranges of at most
    ~bucket_elems elements.  Gaps (inactive parameters) split buckets so they are never communicated."""
    buckets, cur_s, cur_e = [], None, None
    for off, ln in segments:
        if cur_s is None:
            cur_s, cur_e = off, off + ln
        elif off <= cur_e + 3 and (off + ln - cur_s) <= bucket_elems:   # +3: 16-byte alignment padding
            cur_e = off + ln
        else:
            buckets.append((cur_s, cur_e))
            cur_s, cur_e = off, off + ln
    if cur_s is not None:
        buckets.append((cur_s, cur_e))
    return buckets


class _ReadyMarker(torch.autograd.Function):
    """Identity on its inputs; its backward runs `callback()` once the gradients of ALL outputs have arrived,
    i.e. after every consumer of the marked tensors has run its backward."""

    @staticmethod
    def forward(ctx, callback, *tensors):
        ctx.callback = callback
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        ctx.callback()
        return (None, *grads)


def ready_marker(callback, *tensors):
    """Pass `tensors` through a tape node whose backward fires `callback` (see the module docstring)."""
    if callback is None or not any(t.requires_grad for t in tensors):
        return tensors
    return _ReadyMarker.apply(callback, *tensors)


class GradReducer:
    def __init__(self, fp, world_size: int, bucket_mb: float = 25.0, group=None, early_range=None):
        """early_range: (start, end) element range of the flat buffer whose gradients are final when the
        trainer's ready marker fires (the input decoders); buckets are split at its borders."""
        self.world = world_size
        self.group = group
        segs = [(int(a), int(b)) for a, b in fp.segments.cpu().tolist()]
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.early_range = early_range
        if early_range is None:
            self.early, self.late = [], plan_buckets(segs, cap)
        else:
            lo, hi = early_range
            inside = [(a, n) for a, n in segs if a >= lo and a + n <= hi]
            outside = [(a, n) for a, n in segs if not (a >= lo and a + n <= hi)]
            self.early, self.late = plan_buckets(inside, cap), plan_buckets(outside, cap)
        self.buckets = self.early + self.late
        self.scale = torch.tensor([0.0, 1.0 / world_size, 1.0, 0.0], dtype=torch.float32).to(fp.flat.device)
        self.bytes_per_step = sum(e - s for s, e in self.buckets) * 4
        self._fp = fp
        self._works = []
        self._early_done = False
        self.enabled_marker = True
        # NCCL averages inside the collective (ReduceOp.AVG: no extra pass over the 80 MB of gradients); gloo has no AVG: SUM, then one
        # scaling kernel.  Either way every rank ends with the same bits.
        self.use_avg = fp.flat.is_cuda and dist.is_initialized() and dist.get_backend(group) == "nccl"

    def reset(self):
        self._works = []
        self._early_done = False

    def broadcast_state(self, fp, extra=()):
        """Make every rank start from rank 0's state: parameters, Adam moments, per-parameter step counters and any `extra` tensors
        (the optimizer's hyper buffer, BatchNorm buffers).  Without it rank consistency would rest on every caller seeding identically
        and a checkpoint loaded on one rank only would make the ranks diverge silently."""
        if self.world == 1:
            return
        with torch.no_grad():
            for t in (fp.flat, fp.m, fp.v, fp.vmax, fp.param_steps, *extra):
                if t is not None:
                    dist.broadcast(t, src=0, group=self.group)

    def _launch(self, fp, buckets):
        for s, e in reversed(buckets):
            op = dist.ReduceOp.AVG if self.use_avg else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(fp.grad[s:e], op=op, group=self.group, async_op=True))

    def early_ready(self):
        """Marker callback: the early range is final — start its all-reduce now, overlapped with the rest of backward."""
        if self.world == 1 or self._early_done or not self.enabled_marker:
            return
        self._early_done = True
        self._launch(self._fp, self.early)

    def finish(self, fp):
        """Average fp.grad over ranks (call after backward, before clip)."""
        if self.world == 1:
            return
        if not self._early_done:
            self._launch(fp, self.early)
        self._launch(fp, self.late)
        for w in self._works:
            w.wait()
        self._works = []
        self._early_done = False
        if not self.use_avg:
            K.grad_scale(fp.grad, fp.segments, fp.nseg, self.scale)
