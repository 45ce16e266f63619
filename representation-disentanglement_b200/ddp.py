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
    def __init__(self, fp, world_size: int, bucket_mb: float = 25.0, group=None, early_range=None, stages=None):
        """stages: [(name, [(start, end), ...]), ...] element ranges of the flat buffer in the ORDER their gradients become final
        during backward; the trainer's tape markers call `stage_ready(k)` and the stage's buckets are all-reduced from there,
        overlapped with the rest of the backward.  early_range: shorthand for one stage (the input decoders).

        Safety of the early launches does not rest on autograd's node ordering: every accumulation into a parameter's gradient
        goes through ops._sink, which reports the parameter here (`note_sink`).  The number of reports per stage and iteration is
        learned in the first (eager) iteration; from then on a marker may launch its stage only if exactly that many accumulations
        have been issued at the moment it fires — otherwise the stage simply waits for `finish`."""
        self.world = world_size
        self.group = group
        segs = [(int(a), int(b)) for a, b in fp.segments.cpu().tolist()]
        cap = int(bucket_mb * 1024 * 1024 / 4)
        if stages is None:
            stages = [("early", [tuple(early_range)])] if early_range is not None else []
        self.early_range = early_range
        self.stage_names = [n for n, _ in stages]
        self.stage_buckets, taken = [], set()
        for _, ranges in stages:
            inside = [(a, n) for a, n in segs if any(a >= lo and a + n <= hi for lo, hi in ranges) and (a, n) not in taken]
            taken.update(inside)
            self.stage_buckets.append(plan_buckets(inside, cap))
        self.late = plan_buckets([sg for sg in segs if sg not in taken], cap)
        self.early = [b for sb in self.stage_buckets for b in sb]
        self.buckets = self.early + self.late
        # parameter -> stage (by identity of the Parameter object that ops._sink sees)
        self._stage_of = {}
        for p, o in zip(fp.params, fp.offsets):
            for k, (_, ranges) in enumerate(stages):
                if any(o >= lo and o + p.numel() <= hi for lo, hi in ranges):
                    self._stage_of[id(p)] = k
                    break
        ns = len(stages)
        self._count = [0] * ns
        self._expected = [None] * ns          # learned in the first iteration; -1 = inconsistent, never launch early
        self._launched = [False] * ns
        self.scale = torch.tensor([0.0, 1.0 / world_size, 1.0, 0.0], dtype=torch.float32).to(fp.flat.device)
        self.bytes_per_step = sum(e - s for s, e in self.buckets) * 4
        self._fp = fp
        self._works = []
        self.enabled_marker = True
        self.early_launches = 0               # stages launched from a marker in the last iteration (diagnostic)
        # NCCL averages inside the collective (ReduceOp.AVG: no extra pass over the 80 MB of gradients); gloo has no AVG: SUM, then one
        # scaling kernel.  Either way every rank ends with the same bits.
        self.use_avg = fp.flat.is_cuda and dist.is_initialized() and dist.get_backend(group) == "nccl"

    def reset(self):
        self._works = []
        self._count = [0] * len(self._count)
        self._launched = [False] * len(self._launched)

    def note_sink(self, p):
        """ops._sink hook: a backward kernel is about to accumulate into p.grad."""
        k = self._stage_of.get(id(p))
        if k is not None:
            self._count[k] += 1

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

    def stage_can_launch(self, k: int) -> bool:
        return (self.world > 1 and self.enabled_marker and k < len(self._launched) and not self._launched[k]
                and self._expected[k] is not None and self._expected[k] >= 0 and self._count[k] == self._expected[k])

    def stage_ready(self, k: int):
        """Marker callback: stage k's gradients are final (verified by the accumulation count) — start its all-reduce now."""
        if not self.stage_can_launch(k):
            return False
        self._launched[k] = True
        self.early_launches += 1
        self._launch(self._fp, self.stage_buckets[k])
        return True

    def early_ready(self):
        """The single-stage interface of round 1 (stage 0 = the input decoders)."""
        return self.stage_ready(0)

    def finish(self, fp):
        """Average fp.grad over ranks (call after backward, before clip)."""
        if self.world == 1:
            return
        for k in range(len(self._launched)):
            if self._expected[k] is None:
                self._expected[k] = self._count[k]                 # first iteration: learn the stage's accumulation count
            elif self._expected[k] != self._count[k]:
                self._expected[k] = -1                             # the tape changed between iterations: no early launch for this stage
            if not self._launched[k]:
                self._launch(fp, self.stage_buckets[k])
        self._launch(fp, self.late)
        for w in self._works:
            w.wait()
        self._works = []
        self._count = [0] * len(self._count)
        self._launched = [False] * len(self._launched)
        if not self.use_avg:
            K.grad_scale(fp.grad, fp.segments, fp.nseg, self.scale)
