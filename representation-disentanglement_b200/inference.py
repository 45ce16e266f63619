"""Missing-modality inference sweep (BASELINE config 5; reference `evaluate(set='test_dropoff')`, src/main_missing.py:349, with
`TestDropoffDataset`'s drop lists, src/util.py:580-584 — here over ALL 15 non-empty subsets of 4 contrasts).

For every subset the reference zeroes the absent contrasts (src/util.py:610-613), anatomy-encodes the present ones
(`compute_anatomy_encoding`, src/model.py:3135-3157, eval mode), takes the fused rows in (b, m) order
(`reconstruct_output_si_fused`, :3239-3258, Q3: one output row per present (slice, contrast) pair) and runs the output decoder on
them.  Every (subset, present contrast) pair is computed here exactly like that — 32 encodings and 32 decoded rows per slice, nothing
shared between subsets — but as ONE batched pass: the rows of all subsets are stacked contrast-major (4 weight groups of 8 subsets x B
slices), so the encoder, the masked softmax and the output decoder each run once over 32 B rows instead of 15 times over 1..4 B rows
(eval-mode BatchNorm makes every row independent of its batch).  All launches of a sweep are captured in one CUDA graph.

`dedup=True` (not the default, reported separately by bench.py) exploits that a row's result depends only on its contrast and on
whether contrast 0 is present in the subset (the brain mask of the softmax is taken from contrast 0's zeroed input): 7 distinct rows
per slice are computed and the 32 outputs are gathered from them."""
from typing import List

import torch

from . import kernels as K
from . import ops
from .model import MultimodalModel


class SweepRunner:
    def __init__(self, model: MultimodalModel, batch_size: int, use_graph: bool = True, subsets: List[int] = None, dedup: bool = False):
        self.model, self.B, self.M = model, batch_size, model.modality_num
        self.C = model.in_num_ch
        self.H, self.W = model.input_size
        self.dev = model.device
        self.subsets = list(subsets) if subsets is not None else list(range(1, 1 << self.M))
        if not model.shared_ana_enc:
            raise NotImplementedError("rd_b200 SweepRunner: the batched sweep groups rows by contrast over ONE shared anatomy encoder")
        self.inputs = torch.zeros(batch_size, self.M * self.C, self.H, self.W, device=self.dev)
        # [mask_img | ones]: (inputs[:, 0] == 0) of the full input, and what it becomes when contrast 0 is dropped (all zero -> all ones)
        self.masks = torch.ones(2 * batch_size, self.H, self.W, 1, device=self.dev)
        # rows, contrast-major: (contrast m, subset k containing m) -> block of B rows
        self.row_blocks = [(m, k) for m in range(self.M) for k, s in enumerate(self.subsets) if (s >> m) & 1]
        self.rows_per_slice = len(self.row_blocks)
        self.dedup = bool(dedup)
        if self.dedup:      # distinct (contrast, contrast-0-present) pairs, in first-use order
            self.compute_blocks, self.block_of = [], []
            for m, k in self.row_blocks:
                key = (m, bool(self.subsets[k] & 1))
                if key not in self.compute_blocks:
                    self.compute_blocks.append(key)
                self.block_of.append(self.compute_blocks.index(key))
        else:
            self.compute_blocks = [(m, bool(self.subsets[k] & 1)) for m, k in self.row_blocks]
            self.block_of = list(range(len(self.row_blocks)))
        per = [sum(1 for m, _ in self.compute_blocks if m == q) for q in range(self.M)]
        if len(set(per)) != 1:
            # the grouped kernels want equally sized weight groups: pad the smaller groups by repeating their last block
            top = max(per)
            padded = []
            for q in range(self.M):
                blk = [b for b in self.compute_blocks if b[0] == q]
                padded += blk + [blk[-1]] * (top - len(blk))
            remap = {}
            for i, b in enumerate(padded):
                remap.setdefault(b, i)
            self.block_of = [remap[self.compute_blocks[i]] for i in self.block_of]
            self.compute_blocks = padded
        self.out = None
        self.use_graph = use_graph and self.dev.type == "cuda"
        self.graph = None
        self.side = torch.cuda.Stream(device=self.dev) if self.use_graph else None
        self.plan = K.MixFwdPlan(self.dev) if self.dev.type == "cuda" else None
        self.calls = 0
        self.launches = None
        model.eval()

    def load(self, inputs: torch.Tensor, mask_img: torch.Tensor):
        self.inputs.copy_(inputs.to(torch.float32), non_blocking=True)
        self.masks[:self.B, :, :, 0].copy_(mask_img.to(torch.float32), non_blocking=True)

    def rows_of(self, subset_index: int) -> List[int]:
        """Row blocks (of B rows each) of subset `subset_index` in `sweep()`'s output, in contrast order."""
        return [i for i, (m, k) in enumerate(self.row_blocks) if k == subset_index]

    def subset_output(self, subset_index: int) -> torch.Tensor:
        """The reference's output for one subset: (r * B, H, W, out_ch) rows in (b, m) order (src/model.py:3241-3242)."""
        blocks = self.rows_of(subset_index)
        B = self.B
        parts = [self.out[i * B:(i + 1) * B] for i in blocks]                     # per contrast: (B, ...)
        return torch.stack(parts, 1).reshape((len(blocks) * B,) + tuple(self.out.shape[1:]))

    def _body(self):
        model, B, M = self.model, self.B, self.M
        cd = model.cdtype
        if self.plan is not None:
            self.plan.prepare()
            ops.MIX_FWD = self.plan
        try:
            X4 = torch.empty((M * B, self.H, self.W, self.C), dtype=cd, device=self.dev)
            K.stack_modalities(self.inputs, X4, M)
            X = ops.gather_blocks(X4, [m for m, _ in self.compute_blocks], B)                  # every (subset, contrast) pair its own rows
            types = model._types_all
            feats = model.anatomy_encoder_enc_list[0].nhwc(X, types)
            logits = model.anatomy_encoder_dec.nhwc(feats, types)
            if model.others.get("ana_dec_act") == "softplus":
                S = ops.softplus(logits)
            elif model.others.get("softmax_remove_mask", False):
                mk = ops.gather_blocks(self.masks, [0 if with0 else 1 for _, with0 in self.compute_blocks], B)
                S = ops.masked_softmax(logits, mk.reshape(mk.shape[0], self.H, self.W))
            else:
                S = ops.masked_softmax(logits, None)
            y, _ = model.output_decoder.nhwc(model.fuse_rows(S))
            if self.dedup or len(self.block_of) != len(self.compute_blocks):
                y = ops.gather_blocks(y, self.block_of, B)
            self.out = y
        finally:
            ops.MIX_FWD = None

    def sweep(self):
        """One pass over all subsets for the resident batch; returns the (rows_per_slice * B, H, W, out_ch) outputs, row blocks in
        `row_blocks` order (contrast-major); `subset_output(k)` gives one subset's rows in the reference's order."""
        with torch.no_grad():
            if not self.use_graph:
                self._body()
                return self.out
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                self.calls += 1
                if self.calls <= 2:
                    self._body()
                else:
                    if self.graph is None:
                        from . import lib as _lib
                        torch.cuda.synchronize()
                        before = _lib.launch_count(self.dev.index or 0)
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.side):
                            self._body()
                        torch.cuda.synchronize()
                        self.launches = _lib.launch_count(self.dev.index or 0) - before
                        self.graph = g
                    self.graph.replay()
            cur.wait_stream(self.side)
        return self.out
