"""Missing-modality inference sweep (BASELINE config 5; reference `evaluate(set='test_dropoff')`, src/main_missing.py:349, with
`TestDropoffDataset`'s drop lists, src/util.py:580-584 — here over ALL 15 non-empty subsets of 4 contrasts).

For every subset: zero the absent contrasts (src/util.py:610-613), anatomy-encode the present ones (`compute_anatomy_encoding`,
src/model.py:3135-3157, eval mode), gather the fused rows in (b, m) order (`reconstruct_output_si_fused`, :3239-3258, Q3) and run the
output decoder on them.  Each subset is computed independently, exactly as the reference's loop would (no reuse of a contrast's code
across subsets), all launches of a sweep captured in one CUDA graph."""
from typing import List

import torch

from . import kernels as K
from . import ops
from .model import MultimodalModel


class SweepRunner:
    def __init__(self, model: MultimodalModel, batch_size: int, use_graph: bool = True, subsets: List[int] = None):
        self.model, self.B, self.M = model, batch_size, model.modality_num
        self.C = model.in_num_ch
        self.H, self.W = model.input_size
        self.dev = model.device
        self.subsets = list(subsets) if subsets is not None else list(range(1, 1 << self.M))
        self.inputs = torch.zeros(batch_size, self.M * self.C, self.H, self.W, device=self.dev)
        self.mask_img = torch.zeros(batch_size, self.H, self.W, device=self.dev)        # (inputs[:, 0] == 0) of the full input
        self.ones_img = torch.ones(batch_size, self.H, self.W, device=self.dev)         # contrast 0 absent: inputs[:, 0] is all zero
        self.rows_per_slice = sum(bin(s).count("1") for s in self.subsets)
        self.out = [None] * len(self.subsets)
        self.use_graph = use_graph and self.dev.type == "cuda"
        self.graph = None
        self.side = torch.cuda.Stream(device=self.dev) if self.use_graph else None
        self.plan = K.MixFwdPlan(self.dev) if self.dev.type == "cuda" else None
        self.calls = 0
        self.launches = None
        model.eval()

    def load(self, inputs: torch.Tensor, mask_img: torch.Tensor):
        self.inputs.copy_(inputs.to(torch.float32), non_blocking=True)
        self.mask_img.copy_(mask_img.to(torch.float32), non_blocking=True)

    def _body(self):
        model, B, C = self.model, self.B, self.C
        cd = model.cdtype
        if self.plan is not None:
            self.plan.prepare()
            ops.MIX_FWD = self.plan
        try:
            for k, sub in enumerate(self.subsets):
                present = [m for m in range(self.M) if (sub >> m) & 1]
                r = len(present)
                X = torch.empty((r * B, self.H, self.W, C), dtype=cd, device=self.dev)
                for q, m in enumerate(present):
                    K.nchw_to_nhwc(self.inputs, X[q * B:(q + 1) * B], m * C, C)
                types = [model._types_all[m] for m in present]
                if model.shared_ana_enc:
                    feats = model.anatomy_encoder_enc_list[0].nhwc(X, types)
                else:
                    per = [model.anatomy_encoder_enc_list[m].nhwc(X[q * B:(q + 1) * B], [model._types_all[m]]) for q, m in enumerate(present)]
                    feats = [ops.stack_rows([p[j] for p in per]) for j in range(5)]
                logits = model.anatomy_encoder_dec.nhwc(feats, types)
                if model.others.get("ana_dec_act") == "softplus":
                    S = ops.softplus(logits)
                else:
                    mi = self.mask_img if 0 in present else self.ones_img
                    S = ops.masked_softmax(logits, mi if model.others.get("softmax_remove_mask", False) else None)
                ones = torch.ones(B, r, device=self.dev)
                rows, _, _ = ops.fuse_gather(S, ones, B, r)               # (b, m) row-major order of si_cat[mask == 1]
                y, _ = model.output_decoder.nhwc(model.fuse_rows(rows))
                self.out[k] = y
        finally:
            ops.MIX_FWD = None

    def sweep(self):
        """One pass over all subsets for the resident batch; returns the list of outputs (one (r*B, H, W, out_ch) tensor per subset)."""
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
