"""Data-parallel gradient averaging: one process per GPU, NCCL over NVLink 5 / NVSwitch.

Semantics (SURVEY.md §8e — the reference has no DP, so this is "what DDP would do to the reference"):
rank r runs the reference step on its local shard (local BatchNorm statistics, local masked means, local
roll-by-one negatives); parameter gradients are AVERAGED over ranks before clip + Adam; BatchNorm running
buffers stay rank-local.  Parameters that receive no gradient are excluded from the buckets
statically (the segment table of trainer.FlatParams), so every rank reduces the same byte ranges.

The flat fp32 gradient buffer is reduced in a few contiguous buckets (sum, then one scale kernel by 1/N).
Buckets are issued asynchronously on NCCL's stream in reverse parameter order — decoders first, which is the
order in which backward finishes them.
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


class GradReducer:
    def __init__(self, fp, world_size: int, bucket_mb: float = 25.0, group=None):
        self.world = world_size
        self.group = group
        segs = [(int(a), int(b)) for a, b in fp.segments.cpu().tolist()]
        self.buckets = plan_buckets(segs, int(bucket_mb * 1024 * 1024 / 4))
        self.scale = torch.tensor([0.0, 1.0 / world_size, 1.0, 0.0], dtype=torch.float32).to(fp.flat.device)
        self.bytes_per_step = sum(e - s for s, e in self.buckets) * 4

    def finish(self, fp):
        """Average fp.grad over ranks (call after backward, before clip)."""
        if self.world == 1:
            return
        works = []
        for s, e in reversed(self.buckets):
            works.append(dist.all_reduce(fp.grad[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()
        K.grad_scale(fp.grad, fp.segments, fp.nseg, self.scale)
