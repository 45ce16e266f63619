"""The reference training / evaluation loop body on the rd_b200 kernels.

`Trainer.train_iteration` is the body of `train()` in the reference `src/main_missing.py:165-284`:
encode -> decode -> (cycle) -> losses -> weighted sum -> backward -> clip_grad_norm_(1.0) ->
every `16 // batch_size` iterations Adam(amsgrad, wd 1e-5).step() + zero_grad (gradient accumulation
rule of :282-284).  Differences are mechanical only:
  * modality-major NHWC stacks instead of Python lists (the list API of MultimodalModel stays available),
  * no host synchronisation inside the step (masked skip logic on the device, the (i, j) pair and eps
    are drawn on the host *before* the step and handed over in small device buffers) so that the
    whole iteration can be captured in ONE CUDA graph and replayed,
  * parameters, gradients and Adam state live in flat fp32 buffers; backward kernels accumulate directly
    into the gradient buffer, clip + Adam are three launches over all active parameters,
  * data parallel: one process per GPU, gradients averaged with bucketed NCCL all-reduce overlapped with
    backward (see ddp.py).
"""
import math
from typing import Dict, List, Optional

import numpy as np
import os

import torch

from . import kernels as K
from . import ops
from .model import MultimodalModel

LOSS_KEYS = ["recon_y", "recon_y_fused", "recon_x", "recon_x_mix", "kl", "latent_z", "sim_s", "sim_z", "all"]
SEG_ELEMS = 1 << 16


def build_model(config: dict, device) -> MultimodalModel:
    """Constructor call of src/main_missing.py:87-95."""
    others = dict(config["others"])
    if "precision" in config:
        others["precision"] = config["precision"]
    return MultimodalModel(
        input_size=(config["input_height"], config["input_width"]), modality_num=len(config["contrast_list"]),
        in_num_ch=2 * config["block_size"] + 1, out_num_ch=config["out_num_ch"], s_num_ch=config["s_num_ch"],
        z_size=config["z_size"], is_cond=config["is_cond"], is_discrim_s=config.get("is_discrim_s", False),
        is_distri_z=config["is_distri_z"], s_compact_method=config["s_compact_method"],
        s_sim_method=config["s_sim_method"], z_sim_method=config["z_sim_method"],
        shared_ana_enc=config["shared_ana_enc"], shared_mod_enc=config["shared_mod_enc"],
        shared_inp_dec=config["shared_inp_dec"], device=device, input_output_act=config["input_output_act"],
        target_output_act=config["target_output_act"], target_model_name=config["target_model_name"],
        fuse_method=config["fuse_method"], others=others)


FROZEN_BY_FIX_PRETRAIN = ("anatomy_encoder_enc_list.", "anatomy_encoder_dec.", "modality_encoder_list.", "input_decoder_list.")


def apply_fix_pretrain(model: torch.nn.Module, config: dict) -> bool:
    """src/main_missing.py:104-116: with `fix_pretrain` and `continue_train` the stage-1 networks are frozen
    (`requires_grad = False`); only the output decoder trains."""
    if not (config.get("fix_pretrain") and config.get("continue_train")):
        return False
    for name, p in model.named_parameters():
        if name.startswith(FROZEN_BY_FIX_PRETRAIN):
            p.requires_grad = False
    return True


def default_active(model: torch.nn.Module, config: dict) -> List[bool]:
    """Which parameters receive gradients in the reference step (torch leaves `.grad = None` for the rest
    and Adam skips them): not the unused `ModalityEncoderNew.convs`, not the BatchNorm of the last
    anatomy-decoder block, and the output decoder only when one of the y-losses is on."""
    y_on = config["lambda_recon_y"] > 0 or config["lambda_recon_y_fused"] > 0
    act = []
    for name, p in model.named_parameters():
        a = True
        if ".convs." in name and name.startswith("modality_encoder_list"):
            a = False
        if name.startswith("anatomy_encoder_dec.output.bn."):
            a = False
        if name.startswith("output_decoder."):
            a = y_on and not name.startswith("output_decoder.output.bn.")
        act.append(a)
    return act


class FlatParams:
    """All parameters in one flat fp32 buffer (+ flat grad / Adam state).  `active` parameters are those
    that receive gradients (torch.optim.Adam skips parameters whose grad is None)."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters()]
        self.names = [n for n, _ in model.named_parameters()]
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4            # keep every parameter 16-byte aligned
        self.total = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + p.numel()].view_as(p)
                p.grad = self.grad[o:o + p.numel()].view_as(p)
                p._rd_sink = True
        self.m = self.v = self.vmax = None
        self.segments = None
        self.nseg = 0
        self.active_mask = None

    def set_active(self, active: List[bool]):
        """Build the (offset, length) segment table over parameters that receive gradients."""
        segs, seg_param = [], []
        for k, (p, o, a) in enumerate(zip(self.params, self.offsets, active)):
            if not a or not p.requires_grad:
                continue
            n, s = p.numel(), 0
            while s < n:
                l = min(SEG_ELEMS, n - s)
                segs.append((o + s, l))
                seg_param.append(k)
                s += l
        self.active_mask = list(active)
        self.nseg = len(segs)
        dev = self.flat.device
        self.segments = torch.tensor(segs, dtype=torch.int64).reshape(-1, 2).to(dev)
        self.seg_param = torch.tensor(seg_param, dtype=torch.int32).to(dev)
        if getattr(self, "param_steps", None) is None:
            self.param_steps = torch.zeros(len(self.params), dtype=torch.float32, device=dev)     # torch.optim.Adam keeps `step` per parameter
        self.param_flags = torch.zeros(len(self.params), dtype=torch.int32, device=dev)
        if self.m is None:
            self.m = torch.zeros_like(self.flat)
            self.v = torch.zeros_like(self.flat)
            self.vmax = torch.zeros_like(self.flat)
        self.partial = torch.zeros(max(self.nseg, 1), dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=dev)

    def detect_active(self) -> List[bool]:
        """After one backward into a zeroed grad buffer: a parameter is active iff something was accumulated
        into it.  (Host-side, once, outside any captured region.)"""
        nz = []
        for p, o in zip(self.params, self.offsets):
            nz.append(bool((self.grad[o:o + p.numel()] != 0).any().item()))
        return nz


def export_adam_state(fp: "FlatParams", hyper: torch.Tensor) -> dict:
    """The flat Adam state in `torch.optim.Adam(amsgrad=True).state_dict()` format (parameter order = model.parameters()):
    per-parameter step / exp_avg / exp_avg_sq / max_exp_avg_sq for every parameter that has been stepped (torch creates the state
    lazily, parameters whose grad was always None have none), one param group with lr / betas / eps / weight_decay."""
    steps = fp.param_steps.tolist()
    h = hyper.tolist()
    state = {}
    for k, (p, o) in enumerate(zip(fp.params, fp.offsets)):
        if steps[k] <= 0:
            continue
        n = p.numel()
        state[k] = {"step": torch.tensor(float(steps[k])),
                    "exp_avg": fp.m[o:o + n].view_as(p).detach().clone().cpu(),
                    "exp_avg_sq": fp.v[o:o + n].view_as(p).detach().clone().cpu(),
                    "max_exp_avg_sq": fp.vmax[o:o + n].view_as(p).detach().clone().cpu()}
    group = {"lr": h[0], "betas": (h[1], h[2]), "eps": h[3], "weight_decay": h[4], "amsgrad": True, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "params": list(range(len(fp.params)))}
    return {"state": state, "param_groups": [group]}


def import_adam_state(fp: "FlatParams", hyper: torch.Tensor, sd: dict):
    """Inverse of export_adam_state; accepts the `optimizer` entry of a reference checkpoint (src/main_missing.py:330-335)."""
    g = sd["param_groups"][0]
    order = list(g["params"])
    if len(order) != len(fp.params):
        raise ValueError("optimizer state has %d parameters, the model %d" % (len(order), len(fp.params)))
    steps = torch.zeros(len(fp.params), dtype=torch.float32)
    with torch.no_grad():
        fp.m.zero_()
        fp.v.zero_()
        fp.vmax.zero_()
        for k, pid in enumerate(order):
            st = sd["state"].get(pid)
            if st is None:
                continue
            p, o = fp.params[k], fp.offsets[k]
            n = p.numel()
            steps[k] = float(st["step"])
            fp.m[o:o + n].copy_(st["exp_avg"].reshape(-1))
            fp.v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            if "max_exp_avg_sq" in st:
                fp.vmax[o:o + n].copy_(st["max_exp_avg_sq"].reshape(-1))
        fp.param_steps.copy_(steps)
        b = g.get("betas", (0.9, 0.999))
        hyper.copy_(torch.tensor([g["lr"], b[0], b[1], g.get("eps", 1e-8), g.get("weight_decay", 1e-5), float(steps.max()),
                                  1 - b[0], 1 - b[1]], dtype=torch.float32))


class Trainer:
    def __init__(self, model: MultimodalModel, config: dict, batch_size: int, use_graph: bool = False,
                 ddp=None, active: Optional[List[bool]] = None):
        self.model, self.cfg, self.B = model, config, batch_size
        self.M = model.modality_num
        self.C = model.in_num_ch
        self.H, self.W = model.input_size
        self.dev = model.device
        self.use_graph = use_graph
        self.ddp = ddp
        self.fp = FlatParams(model)
        self.fp.set_active(active if active is not None else default_active(model, config))
        self.accum_every = max(1, 16 // batch_size) if batch_size <= 16 else None
        if self.accum_every is None:
            raise ValueError("batch_size > 16 makes the reference's `16 // batch_size` accumulation rule divide by zero (Q11)")
        self.iter = 0                 # iterations since construction (graph warm-up bookkeeping)
        self.epoch_iter = 0           # the reference's `iter` of `enumerate(trainDataLoader)`: restarts every epoch (start_epoch)
        if self.M * self.M > 16:
            raise ValueError("rd_b200 Trainer: %d contrasts need %d (i, j) decodes per step; the grouped kernels take at most 16 weight "
                             "groups per launch (M <= 4, the reference's data sets have 2-4 contrasts)" % (self.M, self.M * self.M))
        self.graph_warmup = 2 * self.accum_every    # eager iterations before capture (allocator / lazy-init warm-up)
        dev = self.dev
        B, M = self.B, self.M
        # static device buffers (graph inputs)
        self.inputs = torch.zeros(B, M * self.C, self.H, self.W, device=dev)
        self.targets = torch.zeros(B, 1, self.H, self.W, device=dev)
        self.mask = torch.ones(B, M, device=dev)
        self.mask_img = torch.zeros(B, self.H, self.W, device=dev)
        self.eps = torch.zeros(M * B, model.z_size, device=dev)
        self.pair = torch.zeros(2, dtype=torch.int32, device=dev)
        self._stage, self._staged = None, False          # prefetch(): staging copies of the six buffers above
        self.hyper = torch.tensor([config["lr"], 0.9, 0.999, 1e-8, 1e-5, 0.0, 1 - 0.9, 1 - 0.999], dtype=torch.float32).to(dev)   # [6], [7]: 1 - beta rounded from double like torch
        lam = [config["lambda_recon_y"], config["lambda_recon_y_fused"], config["lambda_recon_x"],
               config["lambda_recon_x_mix"], config["lambda_kl"], config["lambda_latent_z"], config["lambda_sim_s"],
               config["lambda_sim_z"]]
        self.lambdas_host = lam
        self.lambdas = torch.tensor(lam, dtype=torch.float32).to(dev)
        self.lambdas_eval = torch.tensor(lam[:4] + [0.0] + lam[5:], dtype=torch.float32).to(dev)
        if config["lambda_recon_y_fused"] > 0:
            self.use_graph = False      # the fused output has K = mask.sum() rows (data dependent, src/model.py:3241-3242): eager only
        self.loss_vec = torch.zeros(len(LOSS_KEYS), device=dev)
        self.graphs = {}
        self.side = torch.cuda.Stream(device=dev) if (use_graph and dev.type == "cuda") else None
        self.last = None
        self.mix_batch = None
        self.mix_plan = None
        self.use_mix_plan = os.environ.get("RD_B200_NO_MIX_PLAN") is None
        # capture the NCCL all-reduces inside the iteration graph (RD_B200_DDP_IN_GRAPH=0: two graphs around eager
        # collectives); measured at N = 2: 32.6 ms / step in-graph vs 34.8 ms eager
        import os as _os
        self.ddp_in_graph = _os.environ.get("RD_B200_DDP_IN_GRAPH", "1") == "1"
        self.launches_per_graph = {}
        self._pinned = None
        self._stats_stream, self._stats_pending, self._stats_keep = None, False, None

    # ------------------------------------------------------------------ host -> device feed
    def load_batch(self, batch: dict, eps=None, pair=None, draw_eps: bool = True):
        """Copy one reference-layout batch dict (src/util.py:566) into the static device buffers.
        eps: list of M (B, Z) CPU tensors or None (drawn like MultimodalModel.sample, CPU torch.normal);
        pair: (i, j) or None (np.random.choice, src/model.py:3485)."""
        B, M = self.B, self.M
        self._check_batch(batch, B)
        self.inputs.copy_(batch["inputs"].to(torch.float32), non_blocking=True)
        self.targets.copy_(batch["targets"].to(torch.float32), non_blocking=True)
        self.mask.copy_(batch["mask"].to(torch.float32), non_blocking=True)
        self.mask_img.copy_(batch["mask_img"].to(torch.float32), non_blocking=True)
        if eps is not None:
            self.eps.copy_(torch.cat([e.reshape(B, -1) for e in eps], 0), non_blocking=True)
        elif draw_eps:      # evaluation (phase 'test': z = z_mean) never calls sample(): the generator must not advance there
            self.eps.copy_(self._draw_eps(torch.empty(M * B, self.model.z_size)), non_blocking=True)
        if pair is None:
            pair = self.model.draw_pair(M) if M > 1 else (0, 0)
        self.pair.copy_(torch.tensor([int(pair[0]), int(pair[1])], dtype=torch.int32), non_blocking=True)

    def _draw_eps(self, out: torch.Tensor) -> torch.Tensor:
        """The reparameterisation noise on the CPU default generator in the reference's call order (MultimodalModel.sample,
        src/model.py:3159-3162, Q8): one torch.normal(0, 1, (B, Z)) per contrast for the encoding, and — when the cycle term is on — M more
        draws for the re-encoding of x_fake with phase='train' (src/main_missing.py:231) whose samples are never used (only z_mean enters
        the latent loss) but advance the generator."""
        B, M, Z = self.B, self.M, self.model.z_size
        for m in range(M):
            out[m * B:(m + 1) * B].copy_(torch.normal(0, 1, size=(B, Z)))
        if self.cfg["lambda_latent_z"] > 0 and self.model.training:
            for m in range(M):
                torch.normal(0, 1, size=(B, Z))
        return out

    def start_epoch(self):
        """Call at the top of every epoch: the reference's accumulation rule uses the per-epoch iteration index (main_missing.py:155,
        282); gradients left over from an incomplete window stay accumulated, as in the reference."""
        self.epoch_iter = 0

    def _check_batch(self, batch: dict, rows: int):
        """The static device buffers hold exactly `rows` slices: anything else must not be broadcast into them by copy_ (a final batch
        of one row would silently be trained on `rows` times).  The reference's DataLoaders do not drop the last batch
        (src/util.py:706): pass a smaller final batch to train_iteration(batch) directly — it then runs eagerly on its own rows."""
        n = int(batch["inputs"].shape[0])
        want = (rows, self.M * self.C, self.H, self.W)
        if tuple(batch["inputs"].shape) != want:
            raise ValueError("rd_b200 Trainer: batch['inputs'] has shape %s, the trainer was built for %s (per-GPU batch %d); "
                             "a smaller last batch goes through train_iteration(batch), prefetch() needs full batches"
                             % (tuple(batch["inputs"].shape), want, rows))
        for k, shp in (("targets", (n, 1, self.H, self.W)), ("mask", (n, self.M)), ("mask_img", (n, self.H, self.W))):
            if tuple(batch[k].shape) != shp:
                raise ValueError("rd_b200 Trainer: batch[%r] has shape %s, expected %s" % (k, tuple(batch[k].shape), shp))

    def prefetch(self, batch: dict, eps=None, pair=None):
        """Stage the NEXT iteration's batch (the DataLoader-prefetch step of a training loop): the host -> device copies run on a
        copy stream into staging buffers while the current iteration computes; the next `train_iteration()` called without a batch
        swaps them into the static buffers with device-side copies.  Host tensors should be pinned."""
        if self.dev.type != "cuda":
            return self.load_batch(batch, eps, pair)
        B, M = self.B, self.M
        self._check_batch(batch, B)
        if self._stage is None:
            self._stage = {k: torch.empty_like(getattr(self, k)) for k in ("inputs", "targets", "mask", "mask_img", "eps", "pair")}
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._stage_ready = torch.cuda.Event()
            self._swap_done = None
            # small host-generated tensors (eps, pair) and batches that arrive unpinned / not fp32 go through persistent PINNED host
            # buffers, two of each used alternately: a pageable source would make cudaMemcpyAsync stage the data synchronously inside
            # the driver (the host then waits for whatever the copy stream is waiting on) and allocate a temporary per step
            self._pin = [{k: torch.empty(getattr(self, k).shape, dtype=getattr(self, k).dtype).pin_memory()
                          for k in ("inputs", "targets", "mask", "mask_img", "eps", "pair")} for _ in range(2)]
            self._pin_evt = [None, None]
            self._pin_idx = 0
        slot = self._pin_idx
        self._pin_idx ^= 1
        if self._pin_evt[slot] is not None:
            self._pin_evt[slot].synchronize()             # the copies issued from this slot two prefetches ago have completed
        pin = self._pin[slot]
        if eps is None:
            self._draw_eps(pin["eps"])
        else:
            for m, e in enumerate(eps):
                pin["eps"][m * B:(m + 1) * B].copy_(e.reshape(B, -1))
        if pair is None:
            pair = self.model.draw_pair(M) if M > 1 else (0, 0)
        pin["pair"][0], pin["pair"][1] = int(pair[0]), int(pair[1])
        src = {"eps": pin["eps"], "pair": pin["pair"]}
        for k in ("inputs", "targets", "mask", "mask_img"):
            t = batch[k]
            if t.dtype == torch.float32 and t.is_pinned() and t.is_contiguous():
                src[k] = t                                # the DataLoader's pinned batch: copied from where it is
            else:
                pin[k].copy_(t)                           # host-side cast / gather into the pinned staging buffer
                src[k] = pin[k]
        st = self._stage
        if self._swap_done is not None:
            self._copy_stream.wait_event(self._swap_done)       # the previous swap must have read the staging buffers
        with torch.cuda.stream(self._copy_stream):
            for k in ("inputs", "targets", "mask", "mask_img", "eps", "pair"):
                st[k].copy_(src[k], non_blocking=True)
            self._stage_ready.record()
            ev = torch.cuda.Event()
            ev.record()
            self._pin_evt[slot] = ev
        self._staged = True

    def _swap_in_staged(self):
        """On the stream the iteration runs on: wait for the staged batch, copy it into the static buffers (device side)."""
        torch.cuda.current_stream().wait_event(self._stage_ready)
        for k, v in self._stage.items():
            getattr(self, k).copy_(v, non_blocking=True)
        self._swap_done = torch.cuda.Event()
        self._swap_done.record()
        self._staged = False

    # ------------------------------------------------------------------ forward + losses on stacks
    def _stack_inputs(self):
        B, M, C = self.B, self.M, self.C
        cd = self.model.cdtype
        # bf16: the encoders' first convolutions gather 16-channel vectors — the slabs are written zero-padded (7 -> 16) right here
        # (not when the modality encoder concatenates the anatomy code behind the image channels, mod_enc_s)
        pad = cd == torch.bfloat16 and self.model.modality_encoder_list[0].s_num_ch == 0
        X = torch.empty((M * B, self.H, self.W, ops._up8(C) if pad else C), dtype=cd, device=self.dev)
        K.stack_modalities(self.inputs, X, M)
        if cd == torch.float32:
            return X, X
        Xf = torch.empty((M * B, self.H, self.W, C), dtype=torch.float32, device=self.dev)
        K.stack_modalities(self.inputs, Xf, M)
        return X, Xf

    def forward_losses(self, with_y: bool = False, keep: bool = False, eval_total: bool = False):
        """eval_total: the `loss` of the reference's evaluate() (src/main_missing.py:463-466) — the KL term is computed and reported but,
        unlike in train(), never added to the total."""
        cfg, model = self.cfg, self.model
        B, M = self.B, self.M
        p = cfg["p"]
        training = model.training
        X, Xgt = self._stack_inputs()
        S = model.anatomy_encoding_nhwc(X, self.mask_img)
        use_s = model.modality_encoder_list[0].s_num_ch != 0
        z, mu, lv = model.modality_encoding_nhwc(X, S if use_s else None, "train" if training else "test", self.eps)
        self_combos = [(i, i) for i in range(M)]
        mix_combos = [(i, j) for i in range(M) for j in range(M) if i != j]
        # all M*M decodes in ONE pass (the reference's two loops, src/model.py:3187-3224, are independent per (i, j) and
        # the SPADE decoder has only per-sample InstanceNorm): every decoder module, and so every expert mixing and
        # convolution launch, runs once per step instead of once for the self- and once for the cross-decodes
        all_combos = [(i, j) for i in range(M) for j in range(M)]
        # tensors with several consumers hand each of them its own alias: the gradients are then summed by one rd_add_n launch
        # instead of autograd's chain of at::add kernels
        S, S_sim, S_y, S_yf = ops.fanout(S, 4)
        z, z_sim = ops.fanout(z, 2)
        Sd, zd = S, z
        if torch.is_grad_enabled():
            # tape marker: its backward fires when every decode kernel has run its backward.  There the queued expert-
            # mixing backward of the decoders is flushed (one launch), which makes the gradients of input_decoder_list
            # final -> with DDP their buckets are all-reduced while the encoder backward still runs
            from .ddp import ready_marker
            Sd, zd = ready_marker(self._decoder_grads_ready, S, z)
        Xall = model.decode_nhwc(Sd, zd, all_combos)
        Xself, Xmix = ops.split_blocks(Xall, [all_combos.index(c) for c in self_combos], [all_combos.index(c) for c in mix_combos], B)
        Xself_loss, Xself_cyc, Xself_ana = ops.fanout(Xself, 3)
        y_list = y_fused = None
        if with_y or cfg["lambda_recon_y"] > 0:
            y_list, _ = model.output_decoder.nhwc(model.fuse_rows(S_y), M)
        if with_y or cfg["lambda_recon_y_fused"] > 0:
            rows, _, cnt = ops.fuse_gather(S_yf, self.mask, B, M)
            y_fused, _ = model.output_decoder.nhwc(model.fuse_rows(rows[:int(cnt.item())]))   # K is data dependent (not graph-captured)
        zero = torch.zeros((), device=self.dev)
        L: Dict[str, torch.Tensor] = {k: zero for k in LOSS_KEYS}
        brats = cfg["dataset_name"] == "BraTS"
        if cfg["lambda_recon_y"] > 0:
            if brats:
                # mean over the contrasts present somewhere in the batch (src/model.py:3299-3313 skips the others); the skip is a
                # device-side weight of 0 — no host read of the mask, so this path is part of the captured iteration
                tgt = self.targets.reshape(B, -1)
                w = torch.empty(M, dtype=torch.float32, device=self.dev)
                K.modality_weights(self.mask, w)
                L["recon_y"] = ops.weighted_sum(w, [ops.seg_loss(y_list[i * B:(i + 1) * B], tgt) for i in range(M)])
            else:
                G = ops.gather_blocks(ops.to_nhwc(self.targets, torch.float32), [0] * M, B)
                L["recon_y"] = ops.masked_recon_loss(y_list, G, self.mask, B, M, 0, p)
        if cfg["lambda_recon_y_fused"] > 0:
            L["recon_y_fused"] = self._recon_y_fused_loss(y_fused, brats, p)
        if cfg["lambda_recon_x"] > 0:
            L["recon_x"] = ops.masked_recon_loss(Xself_loss, Xgt, self.mask, B, M, 0, p)
        if cfg["lambda_recon_x_mix"] > 0:
            L["recon_x_mix"] = ops.masked_recon_loss(Xmix, Xgt, self.mask, B, M, 1, p)
        if cfg["lambda_kl"] > 0:
            L["kl"] = ops.kl_loss(mu, lv, self.mask, B, M, mu.shape[1])
        mu_new = None
        if cfg["lambda_latent_z"] > 0:
            # cycle: re-encode x_fake (src/main_missing.py:230-231).  With mod_enc_s False the second anatomy
            # encoding has no gradient path (Q7) but must still run: it updates the BatchNorm running statistics.
            S_new = None
            if use_s:
                S_new = model.anatomy_encoding_nhwc(Xself_ana, self.mask_img)
            elif training:
                # the code itself is unused: only the BatchNorm running-statistics updates of the reference's call are reproduced
                # (the blocks after the last BatchNorm are skipped; in eval mode the call has no effect at all)
                self._stats_pass(Xself)
            _, mu_new, _ = model.modality_encoding_nhwc(Xself_cyc, S_new if use_s else None, "test")
            L["latent_z"] = ops.latent_z_loss(mu, mu_new, self.mask, B, M, mu.shape[1])
        if cfg["lambda_sim_s"] > 0 and M > 1:
            L["sim_s"] = ops.sim_s_loss(model.compact_nhwc(S_sim), self.mask, self.pair, 0.1, B, M)
        if cfg["lambda_sim_z"] > 0 and M > 1:
            L["sim_z"] = ops.sim_z_loss(z_sim, self.mask, 0.1, B, M, z.shape[1])
        L["all"] = ops.weighted_sum(self.lambdas_eval if eval_total else self.lambdas, [L[k] for k in LOSS_KEYS[:-1]])
        if not getattr(self, "_defer_stats_join", False):
            self._join_stats_pass()                # called on its own (tests, evaluation): no work left on the side stream on return
        out = {"losses": L}
        if keep:
            out["tensors"] = {"S": S, "z": z, "z_mean": mu, "z_log_var": lv, "x_fake": Xself, "x_fake_mix": Xmix, "x_gt": Xgt,
                              "y_fake_list": y_list, "y_fake_fused": y_fused, "z_mean_new": mu_new}
        return out

    def _recon_y_fused_loss(self, y_fused, brats: bool, p: int):
        """src/main_missing.py:200-205: compute_segmentation_loss_y / compute_recon_loss_y of the K fused rows against the B
        targets.  The reference only works for compatible row counts and raises otherwise; so does this (same conditions):
        BraTS (cross_entropy): K == B; other datasets (broadcast of `gt - y`): K == B, or B == 1 (the single target broadcasts
        over the K rows), or K == 1."""
        B = self.B
        Kr = y_fused.shape[0]
        if brats:
            if Kr != B:
                raise ValueError("Expected input batch_size (%d) to match target batch_size (%d)." % (Kr, B))   # F.cross_entropy's message
            return ops.seg_loss(y_fused, self.targets.reshape(B, -1))
        tg = ops.to_nhwc(self.targets, torch.float32)
        if Kr == B:
            G = tg
        elif B == 1:
            G = ops.gather_blocks(tg, [0] * Kr, 1)
        else:
            raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (%d) at non-singleton dimension 0" % (B, Kr))
        ones = torch.ones(Kr, 1, device=self.dev)
        return ops.masked_recon_loss(y_fused, G, ones, Kr, 1, 0, p)

    # ------------------------------------------------------------------ one iteration
    def _decoder_grads_ready(self):
        ops.flush_mix_bwd()
        if self.ddp is not None and self.ddp.world > 1:
            self.ddp.stage_ready(0)

    def _stage_grads_ready(self, k: int):
        """Tape markers inside the encoders (data parallel only): stage k's gradients are final when the reducer's accumulation count
        says so — flush the stage's deferred expert-mixing backward and start its all-reduce under the rest of the backward."""
        if self.ddp is not None and self.ddp.stage_can_launch(k):
            ops.flush_mix_bwd()
            self.ddp.stage_ready(k)

    def _fwd_bwd(self, with_y: bool = False, keep: bool = False):
        # (callers other than _body — tools/ddp_parity.py, tests — join the statistics pass themselves or through _clip_step)
        if self.dev.type == "cuda" and self.use_mix_plan:
            if self.mix_plan is None:
                self.mix_plan = K.MixFwdPlan(self.dev)
            self.mix_plan.prepare()              # one launch: every CondConv layer's experts mixed for this iteration
            ops.MIX_FWD = self.mix_plan
        try:
            self._defer_stats_join = True          # the forked statistics pass is joined at the end of the iteration (_body)
            out = self.forward_losses(with_y=with_y, keep=keep)
        finally:
            self._defer_stats_join = False
            ops.MIX_FWD = None
        L = out["losses"]
        if self.dev.type == "cuda":
            if self.mix_batch is None:
                self.mix_batch = K.MixBwdBatch(self.dev)
            self.mix_batch.begin_iteration()
            ops.MIX_BATCH = self.mix_batch
        try:
            L["all"].backward()
            ops.flush_mix_bwd()
        finally:
            ops.MIX_BATCH = None
        K.cast(torch.stack([L[k].detach().reshape(()) for k in LOSS_KEYS]), self.loss_vec)
        return out

    def _clip_step(self, do_step: bool):
        self._join_stats_pass()
        fp = self.fp
        K.grad_norm(fp.grad, fp.segments, fp.nseg, fp.partial, fp.scalars, 1.0)
        if do_step:
            # clip scale, Adam and zero_grad in one pass; gradients outside the active segments are never written, so they stay zero
            # ... with torch.optim.Adam's "grad is None -> skip this parameter" rule (per-parameter step counters): modules the masked
            # loss terms did not reach in this accumulation window are left untouched like in the reference
            K.clip_adam_amsgrad_gated(fp.flat, fp.grad, fp.m, fp.v, fp.vmax, fp.segments, fp.seg_param, fp.nseg, fp.partial,
                                      fp.param_flags, fp.param_steps, self.hyper, fp.scalars, True)
        else:
            K.grad_scale(fp.grad, fp.segments, fp.nseg, fp.scalars)      # accumulation iteration: the clipped gradient stays (main_missing.py:272)

    def _stats_pass(self, Xself):
        """The cycle's second anatomy encoding (no gradient path, result unused: BatchNorm running statistics only).  Nothing on the main
        stream depends on it, so on a GPU it is forked onto its own stream — inside the captured iteration a parallel branch of the graph —
        and joined at the end of the iteration (`_join_stats_pass`); its small, latency-bound launches then fill the gaps of the loss
        kernels and the backward instead of standing in their way.  RD_B200_STATS_STREAM=0: inline on the main stream."""
        full = os.environ.get("RD_B200_FULL_CYCLE_ENC") is not None
        x = Xself.detach()
        if self.dev.type != "cuda" or os.environ.get("RD_B200_STATS_STREAM", "1") == "0":
            with torch.no_grad():
                self.model.anatomy_encoding_nhwc(x, self.mask_img, stats_only=not full)
            return
        if self._stats_stream is None:
            self._stats_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream()
        side = self._stats_stream
        side.wait_stream(main)
        self._stats_keep = x                        # the main stream must not recycle x_fake's memory before the join
        plan, batch_q = ops.MIX_FWD, ops.MIX_BATCH
        with torch.cuda.stream(side), torch.no_grad():
            self.model.anatomy_encoding_nhwc(x, self.mask_img, stats_only=not full)
        ops.MIX_FWD, ops.MIX_BATCH = plan, batch_q
        self._stats_pending = True

    def _join_stats_pass(self):
        if getattr(self, "_stats_pending", False):
            torch.cuda.current_stream().wait_stream(self._stats_stream)
            self._stats_pending = False
            self._stats_keep = None

    def _body(self, do_step: bool, with_y: bool = False, keep: bool = False):
        out = self._fwd_bwd(with_y, keep)
        if self.ddp is not None:
            self.ddp.finish(self.fp)
        self._clip_step(do_step)
        self._join_stats_pass()
        return out

    def train_iteration(self, batch: Optional[dict] = None, eps=None, pair=None, with_y: bool = False, keep: bool = False):
        """One loop body.  Returns the device loss vector (order LOSS_KEYS); nothing is synchronised.
        In CUDA-graph mode all work (eager warm-up iterations, capture, replays) runs on one side stream that is
        fenced against the caller's current stream on both sides, so autograd never ties the captured region to
        work on the legacy default stream."""
        if not self.use_graph:
            return self._iteration(batch, eps, pair, with_y, keep)
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            r = self._iteration(batch, eps, pair, with_y, keep)
        cur.wait_stream(self.side)
        return r

    def _narrow_rows(self, b: int):
        """Context: the static buffers narrowed to the first b rows (a final batch smaller than the trainer's batch size)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            names = ("B", "inputs", "targets", "mask", "mask_img", "eps")
            saved = {k: getattr(self, k) for k in names}
            try:
                if b != saved["B"]:
                    self.B = b
                    self.inputs, self.targets = saved["inputs"][:b], saved["targets"][:b]
                    self.mask, self.mask_img = saved["mask"][:b], saved["mask_img"][:b]
                    self.eps = torch.zeros(self.M * b, self.model.z_size, device=self.dev)
                yield
            finally:
                for k, v in saved.items():
                    setattr(self, k, v)
        return cm()

    def eval_iteration(self, batch: dict, with_y: bool, pair=None):
        """One loop body of the reference's evaluate() (src/main_missing.py:383-533) under no_grad with the model in eval mode and
        phase 'test' (z = z_mean): the loss terms, then the metrics of :503-517 computed ON THE DEVICE (rd_metrics_*): PSNR / SSIM / MSE
        of channel 0 of every cross reconstruction against the contrast it imitates when no y-loss is configured, else Dice / IoU
        (BraTS) or PSNR / SSIM / MSE of the fused output against the target.  Returns (loss vector clone (LOSS_KEYS order), metrics
        dict name -> device tensor with one value per image, kept tensors)."""
        cfg = self.cfg
        b = int(batch["inputs"].shape[0])
        if not 0 < b <= self.B:
            raise ValueError("rd_b200 Trainer.eval_iteration: batch of %d rows, trainer built for at most %d" % (b, self.B))
        with self._narrow_rows(b):
            self.load_batch(batch, None, pair, draw_eps=False)
            self.model.eval()
            with torch.no_grad():
                out = self.forward_losses(with_y=with_y, keep=True, eval_total=True)
            L, T = out["losses"], out["tensors"]
            K.cast(torch.stack([L[k].detach().reshape(()) for k in LOSS_KEYS]), self.loss_vec)
            M, B = self.M, b
            y_on = cfg["lambda_recon_y"] > 0 or cfg["lambda_recon_y_fused"] > 0
            metrics = {}
            if not (cfg["lambda_recon_y"] == 0 and cfg["lambda_recon_y_fused"] == 0):
                yf = T["y_fake_fused"]
                if yf is None:        # lambda_recon_y_fused == 0 after iteration 0: the reference's Python variable still holds iteration 0's
                    yf = getattr(self, "_stale_y_fused", None)       # fused output and the metrics are computed on it (src/main_missing.py:422-430, 512-515)
                    if yf is None:
                        raise RuntimeError("evaluate(): the fused output is first produced with with_y=True (iteration 0)")
                else:
                    self._stale_y_fused = yf
                if cfg["dataset_name"] == "BraTS":
                    n = min(yf.shape[0], B)       # zip-like: the reference iterates range(target.shape[0]) and needs K >= B rows
                    res = torch.empty(n, 2, dtype=torch.float32, device=self.dev)
                    K.metrics_seg(self.targets.reshape(B, -1)[:n].contiguous(), yf[:n].contiguous(), res)
                    metrics = {"dice": res[:, 0], "iou": res[:, 1]}
                else:
                    n = min(yf.shape[0], B)
                    res = torch.empty(n, 3, dtype=torch.float32, device=self.dev)
                    K.metrics_recon(ops.to_nhwc(self.targets, torch.float32), yf[:n].contiguous(), res)
                    metrics = {"ssim": res[:, 0], "psnr": res[:, 1], "rmse": res[:, 2]}
            else:
                Xmix, Xgt = T["x_fake_mix"], T["x_gt"]
                tidx = torch.tensor([j * B + r for i in range(M) for j in range(M) if i != j for r in range(B)], dtype=torch.int32).to(self.dev)
                res = torch.empty(Xmix.shape[0], 3, dtype=torch.float32, device=self.dev)
                if Xmix.shape[0]:
                    K.metrics_recon(Xgt, Xmix, res, t_index=tidx)
                metrics = {"ssim": res[:, 0], "psnr": res[:, 1], "rmse": res[:, 2]}
            return self.loss_vec.clone(), metrics, T

    def _ragged_iteration(self, batch, eps, pair, with_y, keep):
        """A final batch with fewer rows than the trainer was built for (the reference's loaders keep it, src/util.py:706): one eager
        iteration on buffers of its own size; the accumulation counter and the optimizer advance as for any other iteration."""
        b = int(batch["inputs"].shape[0])
        names = ("B", "inputs", "targets", "mask", "mask_img", "eps")
        saved = {k: getattr(self, k) for k in names}
        try:
            self.B = b
            self.inputs, self.targets = saved["inputs"][:b], saved["targets"][:b]
            self.mask, self.mask_img = saved["mask"][:b], saved["mask_img"][:b]
            self.eps = torch.zeros(self.M * b, self.model.z_size, device=self.dev)
            self.load_batch(batch, eps, pair)
            self.model.train()
            do_step = ((self.epoch_iter + 1) % self.accum_every) == 0
            self.iter += 1
            self.epoch_iter += 1
            out = self._body(do_step, with_y, keep)
            self.last = out if keep else None
        finally:
            for k, v in saved.items():
                setattr(self, k, v)
        return self.loss_vec

    def _iteration(self, batch, eps, pair, with_y, keep):
        if batch is not None and 0 < int(batch["inputs"].shape[0]) < self.B:
            return self._ragged_iteration(batch, eps, pair, with_y, keep)
        if batch is not None:
            self.load_batch(batch, eps, pair)
        elif self._staged:
            self._swap_in_staged()
        self.model.train()
        do_step = ((self.epoch_iter + 1) % self.accum_every) == 0       # `(iter+1) % (16 // batch_size)` with the PER-EPOCH index, main_missing.py:155, 282
        self.iter += 1
        self.epoch_iter += 1
        if self.use_graph and not with_y and not keep and self.iter > self.graph_warmup:
            multi = self.ddp is not None and self.ddp.world > 1
            if multi and not self.ddp_in_graph:
                return self._iteration_split(do_step)
            # one graph for the whole iteration; with DDP the NCCL all-reduces (stream-ordered, capturable) are part of it
            g = self.graphs.get(do_step)
            if g is None:
                try:
                    g = self._capture(("all", do_step), lambda: self._body(do_step))
                except RuntimeError as e:
                    if not multi:
                        raise
                    import warnings
                    warnings.warn("rd_b200: capturing NCCL in the CUDA graph failed (%s); using two graphs around eager "
                                  "collectives" % str(e)[:200])
                    self.ddp_in_graph = False
                    self.ddp.reset()
                    torch.cuda.synchronize()
                    return self._iteration_split(do_step)
                self.graphs[do_step] = g
            g.replay()
            return self.loss_vec
        out = self._body(do_step, with_y, keep)
        self.last = out if keep else None     # never keep an autograd graph alive across iterations
        return self.loss_vec

    def _iteration_split(self, do_step):
        """Fallback for DDP when NCCL cannot be captured: graph A (forward + backward), eager bucketed all-reduce,
        graph B (clip + Adam)."""
        ga = self.graphs.get("fwd_bwd")
        if ga is None:
            self.ddp.enabled_marker = False      # a captured marker callback must not launch eager collectives
            ga = self.graphs["fwd_bwd"] = self._capture("fwd_bwd", lambda: self._fwd_bwd())
        ga.replay()
        self.ddp.finish(self.fp)
        gb = self.graphs.get(("clip", do_step))
        if gb is None:
            gb = self.graphs[("clip", do_step)] = self._capture(("clip", do_step), lambda: self._clip_step(do_step))
        gb.replay()
        return self.loss_vec

    def _capture(self, key, fn):
        """Capture a region in a CUDA graph (stream capture records the launches; nothing runs until replay)."""
        from . import lib as _lib
        torch.cuda.synchronize()
        before = _lib.launch_count(self.dev.index or 0)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.side):
            fn()
        torch.cuda.synchronize()
        self.launches_per_graph[key] = _lib.launch_count(self.dev.index or 0) - before
        return g

    def make_reducer(self, world: int, bucket_mb: float = 25.0, group=None):
        """GradReducer whose early buckets are the input decoders' gradients (final when the decode backward is done)."""
        from .ddp import GradReducer

        def span(prefixes):
            lo = hi = None
            for n, p, o in zip(self.fp.names, self.fp.params, self.fp.offsets):
                if n.startswith(prefixes):
                    lo = o if lo is None else min(lo, o)
                    hi = max(hi or 0, o + (p.numel() + 3) // 4 * 4)
            return None if lo is None else (lo, hi)
        # readiness stages in the order the backward finishes them (created last = differentiated first): the decoders (+ the output
        # decoder in stage 2), the modality encoder, the anatomy decoder, the two deepest anatomy-encoder blocks (4/5 of its bytes).
        # What remains for `finish` is anatomy_encoder_enc_list down_1..down_3 (2 MB).
        stage_defs = [("decoders", ("input_decoder_list.", "output_decoder.")), ("modality_encoder", ("modality_encoder_list.",)),
                      ("anatomy_decoder", ("anatomy_encoder_dec.",))]
        if self.model.shared_ana_enc:
            stage_defs.append(("anatomy_encoder_deep", ("anatomy_encoder_enc_list.0.down_4.", "anatomy_encoder_enc_list.0.down_5.")))
        stages = []
        for name, prefixes in stage_defs:
            ranges = [r for r in (span((pf,)) for pf in prefixes) if r is not None]
            stages.append((name, ranges))
        self.ddp = GradReducer(self.fp, world, bucket_mb, group, stages=stages)
        if world > 1:
            ops.SINK_HOOK = self.ddp.note_sink
            self.model.bwd_markers = {"modality_encoder": lambda: self._stage_grads_ready(1),
                                      "anatomy_decoder": lambda: self._stage_grads_ready(2),
                                      "anatomy_encoder_deep": lambda: self._stage_grads_ready(3)}
        # rank 0's parameters, optimizer state and BatchNorm buffers everywhere (ranks may have been seeded or restored differently)
        self.ddp.broadcast_state(self.fp, extra=[self.hyper] + [b for b in self.model.buffers()])
        return self.ddp

    # ------------------------------------------------------------------ optimizer state (checkpoint contract, src/main_missing.py:126, 330-335)
    def set_lr(self, lr: float):
        """Drive the learning rate from a scheduler (ReduceLROnPlateau in the reference, src/main_missing.py:119, 320)."""
        self.hyper[0:1].copy_(torch.tensor([float(lr)], dtype=torch.float32), non_blocking=False)

    def get_lr(self) -> float:
        return float(self.hyper[0].item())

    def optimizer_state_dict(self) -> dict:
        """`torch.optim.Adam(amsgrad=True).state_dict()` format, see export_adam_state."""
        return export_adam_state(self.fp, self.hyper)

    def load_optimizer_state_dict(self, sd: dict):
        """Accepts the `optimizer` entry of a reference checkpoint (src/util.py:870-903), see import_adam_state."""
        import_adam_state(self.fp, self.hyper, sd)

    def losses_host(self) -> Dict[str, float]:
        v = self.loss_vec.tolist()
        return dict(zip(LOSS_KEYS, v))

    def grad_norm_host(self) -> float:
        return float(self.fp.scalars[0].item())
