"""TEST INFRASTRUCTURE — CPU oracle for the representation-disentanglement hot path.

A from-scratch *functional* restatement (plain PyTorch fp32 on CPU, autograd for the
gradients) of the reference path `src/model.py` (MultimodalModel and the blocks it
instantiates with `src/config.yaml`) and of the loop body of `src/main_missing.py`.
It is keyed by the reference `state_dict` names, so any checkpoint / state dict of the
reference (or of the product model, which keeps the same keys) can be evaluated.

Pinned (oracle/make_golden.py, run in the build container where /root/reference exists)
against the unmodified reference on identical weights / inputs / eps / (i,j):
losses, every parameter gradient, s_i, z, x_fake, x_fake_mix, y_fake, BN running stats.
The committed fixtures live in tests/golden/.  The reference has no tests or golden vectors
of its own, so "parity" = agreement with the reference *run here*.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package never does.

Bug-compatible behaviours reproduced on purpose: Q1 (activation strings
'lrelu'/'relu' resolve to identity), Q2 (per-sample conv loop), Q3 (boolean gather fusion),
Q4 (x_mix index lag), Q5 (decoder half indexed by the anatomy source), Q7 (mod_enc_s False).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

DEFAULT_CFG = {  # src/config.yaml:1-91 (+ derived keys of src/main_missing.py:26-28,75-86)
    "phase": "train", "dataset_name": "BraTS", "contrast_list": ["T1", "T1c", "T2", "T2_FLAIR"],
    "norm_type": "z-score", "block_size": 3, "batch_size": 8, "lr": 0.0002, "p": 1,
    "s_num_ch": 4, "z_size": 16, "lambda_recon_y": 0.0, "lambda_recon_y_fused": 0.0,
    "lambda_recon_x": 1.0, "lambda_recon_x_mix": 2.0, "lambda_sim_s": 10.0, "lambda_sim_z": 2.0,
    "s_compact_method": "max", "s_sim_method": "cosine", "z_sim_method": "cosine",
    "lambda_kl": 0.0, "lambda_latent_z": 0.1, "lambda_adv_s": 0.0, "is_cond": True,
    "is_distri_z": False, "shared_ana_enc": True, "shared_mod_enc": True, "shared_inp_dec": False,
    "others": {"mod_enc_s": False, "ana_dec_act": "softmax", "old": False, "softmax_remove_mask": True},
    "out_num_ch": 1, "input_height": 160, "input_width": 192, "fuse_method": "mean",
    "target_model_name": "U+SA", "input_output_act": "no", "target_output_act": "no",
}


class RDOracle:
    """Functional evaluator over a reference-keyed state dict.

    `state` maps reference state_dict keys to CPU fp32 tensors.  Float parameters that should
    receive gradients must already have requires_grad=True; BN buffers are updated in place in
    training mode (momentum 0.1, unbiased running variance — torch.nn.BatchNorm2d defaults,
    reference src/model.py:2132,2179).
    """

    def __init__(self, state: Dict[str, Tensor], cfg: dict, training: bool = True,
                 batched_condconv: bool = False):
        self.P = state
        self.cfg = cfg
        self.training = training
        self.M = len(cfg["contrast_list"])
        self.H, self.W = cfg["input_height"], cfg["input_width"]
        # exact algebraic identity: inputs_type is constant over the batch at every
        # call site, so one batched conv equals the per-sample loop.  Default False = faithful.
        self.batched_condconv = batched_condconv

    # ------------------------------------------------------------------ primitive blocks
    def cond_conv(self, x: Tensor, pre: str, t: float, stride: int, pad: int) -> Tensor:
        """CondConv2d.forward, src/model.py:2108-2117 (routing :2071-2073)."""
        P = self.P
        bs = x.shape[0]
        tt = torch.full((bs, 1), float(t), dtype=x.dtype)
        r = torch.sigmoid(F.linear(tt, P[pre + "._routing_fn.fc.weight"], P[pre + "._routing_fn.fc.bias"]))
        kern = torch.sum(r[:, :, None, None, None, None] * P[pre + ".weight"], 1)
        if self.batched_condconv:
            return F.conv2d(x, kern[0], P[pre + ".bias"], stride, pad)
        outs = [F.conv2d(x[b:b + 1], kern[b], P[pre + ".bias"], stride, pad) for b in range(bs)]
        return torch.cat(outs, 0)

    def conv(self, x: Tensor, pre: str, stride: int, pad: int, bias: bool = True) -> Tensor:
        return F.conv2d(x, self.P[pre + ".weight"], self.P[pre + ".bias"] if bias else None, stride, pad)

    def bn(self, x: Tensor, pre: str) -> Tensor:
        P = self.P
        if self.training:
            P[pre + ".num_batches_tracked"] += 1
        return F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"], P[pre + ".weight"],
                            P[pre + ".bias"], self.training, 0.1, 1e-5)

    @staticmethod
    def up2_ac(x: Tensor) -> Tensor:  # nn.Upsample(scale_factor=2, bilinear, align_corners=True) :2175
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)

    @staticmethod
    def up2(x: Tensor) -> Tensor:  # nn.Upsample(scale_factor=(2,2), bilinear) :2501
        return F.interpolate(x, scale_factor=(2, 2), mode="bilinear", align_corners=False)

    @staticmethod
    def resize(x: Tensor, size) -> Tensor:  # nn.Upsample(size=..., bilinear) :2432
        return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)

    # ------------------------------------------------------------------ anatomy encoder (a2, a3, a4)
    def anatomy_features(self, x: Tensor, t: float, enc: str) -> List[Tensor]:
        """AnatomyEncoderEncNew.forward src/model.py:2233-2245 (activations identity, Q1)."""
        d1 = F.leaky_relu(self.cond_conv(x, enc + ".down_1", t, 2, 1), 0.2)
        feats = [d1]
        h = d1
        for k in (2, 3, 4, 5):
            h = self.bn(self.cond_conv(h, enc + ".down_%d.conv" % k, t, 2, 1), enc + ".down_%d.bn" % k)
            feats.append(h)
        return feats

    def anatomy_logits(self, feats: List[Tensor], t: float) -> Tensor:
        """AnatomyEncoderDecNew.forward src/model.py:2285-2296 via Act_Deconv_BN_Concat_New :2182-2195."""
        dec = "anatomy_encoder_dec"
        h = feats[4]
        for k, skip in ((4, feats[3]), (3, feats[2]), (2, feats[1]), (1, feats[0])):
            u = self.cond_conv(self.up2_ac(h), dec + ".up_%d.conv" % k, t, 1, 1)
            u = self.bn(u, dec + ".up_%d.bn" % k)
            h = torch.cat([skip, u], 1)
        return self.cond_conv(self.up2_ac(h), dec + ".output.conv", t, 1, 1)

    def compute_anatomy_encoding(self, inputs_list: Sequence[Tensor], mask_img: Tensor) -> List[Tensor]:
        """src/model.py:3135-3157."""
        others = self.cfg["others"]
        out = []
        for i in range(self.M):
            enc = "anatomy_encoder_enc_list.%d" % (0 if self.cfg["shared_ana_enc"] else i)
            s = self.anatomy_logits(self.anatomy_features(inputs_list[i], 1 + i, enc), 1 + i)
            if others.get("ana_dec_act") == "softplus":
                s_act = F.softplus(s)
            elif others.get("softmax_remove_mask", False):
                s_act = F.softmax(torch.cat([100 * mask_img.unsqueeze(1), s], 1), dim=1)[:, 1:]
            else:
                s_act = F.softmax(s, dim=1)
            out.append(s_act)
        return out

    # ------------------------------------------------------------------ modality encoder (a5, a6)
    def modality_stats(self, x: Tensor, s: Tensor, t: float, pre: str) -> Tuple[Tensor, Tensor]:
        """ModalityEncoderNew.forward src/model.py:2366-2400."""
        use_s = self.cfg["others"].get("mod_enc_s", True)
        h = torch.cat([x, s], 1) if use_s else x
        for k in range(1, 6):
            h = F.leaky_relu(self.cond_conv(h, pre + ".conv%d" % k, t, 2, 1), 0.2)
        h = h.reshape(-1, 5 * 6 * 128)
        h = F.leaky_relu(F.linear(h, self.P[pre + ".fcs.0.weight"], self.P[pre + ".fcs.0.bias"]), 0.2)
        mu = F.linear(h, self.P[pre + ".mean.weight"], self.P[pre + ".mean.bias"])
        lv = F.linear(h, self.P[pre + ".log_var.weight"], self.P[pre + ".log_var.bias"])
        return mu, lv

    def compute_modality_encoding(self, inputs_list, si_list, phase="train", eps_list=None):
        """src/model.py:3164-3185; eps injected instead of the CPU torch.normal of :3159-3162 (Q8)."""
        z, mus, lvs = [], [], []
        for i in range(self.M):
            pre = "modality_encoder_list.%d" % (0 if self.cfg["shared_mod_enc"] else i)
            mu, lv = self.modality_stats(inputs_list[i], si_list[i], 1 + i, pre)
            if phase == "train":
                eps = eps_list[i] if eps_list is not None else torch.zeros_like(mu)
                zi = mu + eps * torch.exp(0.5 * lv)
            else:
                zi = mu
            z.append(zi), mus.append(mu), lvs.append(lv)
        return z, mus, lvs

    # ------------------------------------------------------------------ SPADE decoder (a7-a10)
    def spade_block(self, pre: str, size, s: Tensor, z: Tensor, t: float) -> Tensor:
        """SPADEBlockNew.forward src/model.py:2438-2454."""
        zn = F.instance_norm(z, eps=1e-5)
        a = self.cond_conv(self.resize(s, size), pre + ".si_layers", t, 1, 1)
        g = self.cond_conv(a, pre + ".gamma", t, 1, 1)
        b = self.cond_conv(a, pre + ".beta", t, 1, 1)
        return self.cond_conv(zn * (1 + g) + b, pre + ".out", t, 1, 1)

    def decode_shared(self, s: Tensor, z: Tensor, t: float) -> Tensor:
        """SPADENewShared.forward src/model.py:2564-2582 (module index -1 of input_decoder_list)."""
        pre = "input_decoder_list.%d" % self.M
        H, W = self.H, self.W
        h = F.linear(z, self.P[pre + ".zi_scaler.weight"], self.P[pre + ".zi_scaler.bias"])
        h = h.reshape(-1, 128, H // 32, W // 32)
        h = self.spade_block(pre + ".sp1", (H // 32, W // 32), s, h, t)
        h = self.spade_block(pre + ".sp2", (H // 16, W // 16), s, self.up2(h), t)
        h = self.spade_block(pre + ".sp3", (H // 8, W // 8), s, self.up2(h), t)
        return self.up2(h)

    def decode_private(self, k: int, s: Tensor, mid: Tensor, t: float) -> Tensor:
        """SPADENewNotShared.forward src/model.py:2615-2632."""
        pre = "input_decoder_list.%d" % k
        H, W = self.H, self.W
        h = self.spade_block(pre + ".sp4", (H // 4, W // 4), s, mid, t)
        h = self.spade_block(pre + ".sp5", (H // 2, W // 2), s, self.up2(h), t)
        h = self.spade_block(pre + ".sp6", (H, W), s, self.up2(h), t)
        h = self.cond_conv(h, pre + ".out", t, 1, 0)
        return F.softplus(h) if self.cfg["input_output_act"] == "softplus" else h

    def decode_full(self, s: Tensor, z: Tensor, t: float) -> Tensor:
        """SPADENew.forward src/model.py:2519-2538 (`shared_inp_dec: True`: module 0 of input_decoder_list)."""
        pre = "input_decoder_list.0"
        H, W = self.H, self.W
        h = F.linear(z, self.P[pre + ".zi_scaler.weight"], self.P[pre + ".zi_scaler.bias"])
        h = h.reshape(-1, 128, H // 32, W // 32)
        for k, d in enumerate((32, 16, 8, 4, 2, 1)):
            if k:
                h = self.up2(h)
            h = self.spade_block(pre + ".sp%d" % (k + 1), (H // d, W // d), s, h, t)
        h = self.cond_conv(h, pre + ".out", t, 1, 0)
        return F.softplus(h) if self.cfg["input_output_act"] == "softplus" else h

    def decode(self, i: int, j: int, si_list, zi_list) -> Tensor:
        """One (anatomy i, modality j) decode: type 1+j, private half i (Q5) src/model.py:3199-3200,3221-3222;
        `shared_inp_dec`: the single SPADENew decoder (:3195, 3216)."""
        if self.cfg["shared_inp_dec"]:
            return self.decode_full(si_list[i], zi_list[j], 1 + j)
        mid = self.decode_shared(si_list[i], zi_list[j], 1 + j)
        return self.decode_private(i, si_list[i], mid, 1 + j)

    def reconstruct_input_si_zi(self, si_list, zi_list):
        return [self.decode(i, i, si_list, zi_list) for i in range(self.M)]

    def reconstruct_input_si_zj(self, si_list, zi_list):
        return [self.decode(i, j, si_list, zi_list) for i in range(self.M) for j in range(self.M) if i != j]

    # ------------------------------------------------------------------ fusion + output decoder (a11, a12)
    def attention_gate(self, pre: str, x: Tensor, g: Tensor) -> Tuple[Tensor, Tensor]:
        """SpatialAttentionLayer.forward src/model.py:1316-1327."""
        xp = self.conv(x, pre + ".W_x", 2, 0, bias=False)
        gp = self.resize(self.conv(g, pre + ".W_g", 1, 0), xp.shape[2:])
        alpha = torch.sigmoid(self.conv(F.relu(xp + gp), pre + ".W_psi", 1, 0))
        alpha_up = self.resize(alpha, x.shape[2:])
        out = self.bn(self.conv(alpha_up * x, pre + ".W_out.0", 1, 0), pre + ".W_out.1")
        return out, alpha_up

    def channel_attention(self, pre: str, x: Tensor) -> Tuple[Tensor, Tensor]:
        """ChannelAttentionLayer.forward src/model.py:1425-1433 (squeeze and excitation, residual form)."""
        gp = x.mean((2, 3))
        down = F.relu(F.linear(gp, self.P[pre + ".W_down.weight"], self.P[pre + ".W_down.bias"]))
        alpha = torch.sigmoid(F.linear(down, self.P[pre + ".W_up.weight"], self.P[pre + ".W_up.bias"]))
        return (1 + alpha[:, :, None, None]) * x, alpha

    def symmetry_gate(self, pre: str, x: Tensor, g: Tensor) -> Tuple[Tensor, Tensor]:
        """SymmetryGateResidualSpatialAttentionLayer.forward src/model.py:1406-1415: the gate sees g and |g - flip_H(g)|."""
        g_diff = (g - torch.flip(g, dims=[2])).abs()
        g_post = F.relu(self.conv(g, pre + ".W_g", 1, 0) + self.conv(g_diff, pre + ".W_g_diff", 1, 0))
        alpha = torch.sigmoid(self.conv(g_post, pre + ".W_psi", 1, 0))
        alpha_up = self.resize(alpha, x.shape[2:])
        out = self.bn(self.conv((1 + alpha_up) * x, pre + ".W_out.0", 1, 0), pre + ".W_out.1")
        return out, alpha_up

    def output_decoder(self, x: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
        """GANShortGeneratorWithSpatialAttention.forward src/model.py:374-390 (U+SA); GANShortGenerator.forward
        src/model.py:287-299 (U: the same U-Net, the skip connections are concatenated without the attention gate);
        GANShortGeneratorWithChannelAttentionAllAndSpatialAttention.forward src/model.py:1111-1135 (U+SA+CA: the skip connection is
        channel attention + spatial attention) and ...AndSymmetrySpatialAttention.forward src/model.py:1041-1065 (U+SSA+CA)."""
        name = self.cfg["target_model_name"]
        assert name in ("U+SA", "U", "U+SA+CA", "U+SSA+CA")
        plain = name == "U"
        pre = "output_decoder"
        d = [F.leaky_relu(self.conv(x, pre + ".down_1.0", 2, 1), 0.2)]
        for k in (2, 3, 4, 5):
            d.append(self.bn(self.conv(d[-1], pre + ".down_%d.conv.0" % k, 2, 1), pre + ".down_%d.conv.1" % k))
        h = d[4]
        alphas = {}
        for k in (4, 3, 2, 1):
            if plain:
                gated = d[k - 1]
            elif name == "U+SA":
                gated, alphas["alpha_%d" % k] = self.attention_gate(pre + ".att_%d" % k, d[k - 1], h)
            else:
                cc, _ = self.channel_attention(pre + ".att_%d_c" % k, d[k - 1])
                gate = self.attention_gate if name == "U+SA+CA" else self.symmetry_gate
                cs, alphas["alpha_%d" % k] = gate(pre + ".att_%d_s" % k, d[k - 1], h)
                gated = cc + cs
            u = self.bn(self.conv(self.up2_ac(h), pre + ".up_%d.up.1" % k, 1, 1), pre + ".up_%d.bn" % k)
            h = torch.cat([gated, u], 1)
        y = self.conv(self.up2_ac(h), pre + ".output.up.1", 1, 1)
        act = self.cfg["target_output_act"]
        if act == "sigmoid":
            y = torch.sigmoid(y)
        elif act == "tanh":
            y = torch.tanh(y)
        elif act != "no":
            y = F.softplus(y)
        return y, alphas

    def reconstruct_output_si_fused(self, si_list, mask: Tensor) -> Tensor:
        """src/model.py:3239-3258 — boolean gather, row-major over (b, m) (Q3)."""
        cat = torch.stack(list(si_list), 1)
        sel = cat[mask == 1]
        if cat.dim() != sel.dim():
            sel = sel.unsqueeze(1)
        fm = self.cfg["fuse_method"]
        if fm == "mean":
            fused = sel.mean(1)
        elif fm == "max":
            fused = sel.max(1)[0]
        elif fm == "mean-max-min":
            fused = torch.cat([sel.mean(1), sel.max(1)[0], sel.min(1)[0]], 1)
        else:
            raise ValueError("No fused method")
        return self.output_decoder(fused)[0]

    def reconstruct_output_si(self, si_list):
        """src/model.py:3230-3237."""
        bs = si_list[0].shape[0]
        return [self.reconstruct_output_si_fused([s], torch.ones(bs, 1)) for s in si_list]

    # ------------------------------------------------------------------ losses (a13-a17)
    @staticmethod
    def recon(gt: Tensor, out: Tensor, p: int) -> Tensor:  # :3260-3266
        dims = list(range(1, gt.dim()))
        return (gt - out).abs().mean(dims) if p == 1 else (gt - out).pow(2).mean(dims)

    def recon_loss_x_list(self, gt_list, x_list, mask, p):  # :3315-3325
        loss, cnt = torch.zeros(()), 0
        for i in range(len(x_list)):
            if mask[:, i].sum() == 0:
                continue
            cnt += 1
            loss = loss + (mask[:, i] * self.recon(gt_list[i], x_list[i], p)).sum() / mask[:, i].sum()
        return loss if cnt == 0 else loss / cnt

    def recon_loss_x_mix_list(self, gt_list, x_list, mask, p):  # :3327-3341, index lag Q4
        loss, idx = torch.zeros(()), 0
        Mn = mask.shape[1]
        for i in range(Mn):
            for j in range(Mn):
                if i == j:
                    continue
                mm = mask[:, i] * mask[:, j]
                if mm.sum() == 0:
                    continue
                loss = loss + (mm * self.recon(gt_list[j], x_list[idx], p)).sum() / mm.sum()
                idx += 1
        return loss if idx == 0 else loss / idx

    def recon_loss_y_list(self, gt, y_list, mask, p):  # :3268-3278
        loss, cnt = torch.zeros(()), 0
        for i in range(len(y_list)):
            if mask[:, i].sum() == 0:
                continue
            cnt += 1
            loss = loss + (mask[:, i] * self.recon(gt, y_list[i], p)).sum() / mask[:, i].sum()
        return loss if cnt == 0 else loss / cnt

    def recon_loss_y(self, gt, y, p):  # :3280-3285
        return self.recon(gt, y, p).mean()

    @staticmethod
    def segmentation_loss_y(gt, y, weight=(1., 5., 5., 5.)):  # :3287-3297
        ce = F.cross_entropy(y, gt.squeeze(1).long(), weight=torch.tensor(weight))
        act = F.softmax(y, dim=1)
        dice = 0
        for c in range(1, 4):
            g = (gt[:, 0] == c).float()
            dice = dice + 1 - 2 * (act[:, c] * g).sum() / ((act[:, c] ** 2 + g ** 2).sum() + 1e-6)
        return ce + dice / 3

    def segmentation_loss_y_list(self, gt, y_list, mask):  # :3299-3313
        loss, cnt = torch.zeros(()), 0
        for i in range(len(y_list)):
            if mask[:, i].sum() == 0:
                continue
            cnt += 1
            loss = loss + self.segmentation_loss_y(gt, y_list[i])
        return loss if cnt == 0 else loss / cnt

    @staticmethod
    def kl_loss_list_standard(mu_list, lv_list, mask):  # :3343-3360
        mu, lv = torch.cat(list(mu_list), 0), torch.cat(list(lv_list), 0)
        m = torch.cat([mask[:, i] for i in range(mask.shape[1])], 0)
        kl = 0.5 * torch.sum(torch.exp(lv) + mu ** 2 - 1. - lv, 1)
        return (kl * m).sum() / m.sum() / len(mu_list)

    @staticmethod
    def latent_z_loss(mu_list, mu_new_list, mask):  # :3384-3394
        loss, cnt = torch.zeros(()), 0
        for i in range(len(mu_list)):
            if mask[:, i].sum() == 0:
                continue
            cnt += 1
            loss = loss + (mask[:, i].unsqueeze(1) * (mu_list[i] - mu_new_list[i]).abs()).sum() / mask[:, i].sum()
        return loss if cnt == 0 else loss / cnt

    @staticmethod
    def cosine(x, y):  # :3407-3415
        xn = torch.sqrt((x ** 2).sum(1) + 1e-8).clamp_min(1e-8)
        yn = torch.sqrt((y ** 2).sum(1) + 1e-8).clamp_min(1e-8)
        return (x * y).sum(1) / (xn * yn)

    def compact_s(self, x):  # :3448-3475
        m = self.cfg["s_compact_method"]
        if m == "max":
            return F.max_pool2d(x, (16, 16)).reshape(x.shape[0], -1)
        if m == "mean":
            return F.avg_pool2d(x, (16, 16)).reshape(x.shape[0], -1)
        raise ValueError("vgg compaction is out of scope (needs downloaded weights)")

    def similarity_s_loss(self, si_list, mask, pair: Optional[Tuple[int, int]], margin=0.1):
        """:3478-3513; the np.random.choice pair (Q9) is passed in as `pair`."""
        if len(si_list) == 1:
            return torch.zeros(())
        i, j = (0, 1) if len(si_list) == 2 else pair
        si, sj = si_list[i], si_list[j]
        si_perm = torch.cat([si[1:], si[0:1]], 0)
        mi_perm = torch.cat([mask[1:, i], mask[0:1, i]], 0)
        mm = mask[:, i] * mask[:, j] * mi_perm
        if mm.sum() > 0:
            a, b, c = self.compact_s(si), self.compact_s(sj), self.compact_s(si_perm)
            sim, sim_mix = self.cosine(a, b), self.cosine(c, a)
            return (mm * torch.clamp_min(margin - sim + sim_mix, 0)).sum() / mm.sum()
        return torch.zeros(())  # reference returns python int 0 here

    def similarity_z_loss(self, zi_list, mask, margin=0.1):  # :3537-3557
        loss, cnt = torch.zeros(()), 0
        if len(zi_list) == 1:
            return loss
        for i in range(len(zi_list) - 1):
            zi = zi_list[i]
            zp = torch.cat([zi[1:], zi[0:1]], 0)
            mp = torch.cat([mask[1:, i], mask[0:1, i]], 0)
            for j in range(i + 1, len(zi_list)):
                mm = mask[:, i] * mask[:, j] * mp
                if mm.sum() == 0:
                    continue
                cnt += 1
                c, cm = self.cosine(zi, zi_list[j]), self.cosine(zi, zp)
                loss = loss + (mm * torch.clamp_min(margin - cm + c, 0)).sum() / mm.sum()
        return loss if cnt == 0 else loss / cnt

    # ------------------------------------------------------------------ one loop body (a18)
    def forward_losses(self, inputs: Tensor, targets: Tensor, mask: Tensor, mask_img: Tensor,
                       eps_list: Sequence[Tensor], pair: Tuple[int, int], with_y: bool = False,
                       keep: bool = False) -> Dict[str, Tensor]:
        """src/main_missing.py:165-251 (forward + loss sum).  Returns a dict of loss tensors and,
        if keep=True, the intermediate tensors under 'tensors'."""
        cfg = self.cfg
        C = 2 * cfg["block_size"] + 1
        xs = [inputs[:, m * C:(m + 1) * C] for m in range(self.M)]
        si = self.compute_anatomy_encoding(xs, mask_img)
        zi, mu, lv = self.compute_modality_encoding(xs, si, "train" if self.training else "test", eps_list)
        x_self = self.reconstruct_input_si_zi(si, zi)
        x_mix = self.reconstruct_input_si_zj(si, zi)
        out: Dict[str, Tensor] = {}
        y_list = y_fused = None
        if with_y or cfg["lambda_recon_y"] > 0:
            y_list = self.reconstruct_output_si(si)
        if with_y or cfg["lambda_recon_y_fused"] > 0:
            y_fused = self.reconstruct_output_si_fused(si, mask)
        total = torch.zeros(())
        z0 = torch.zeros(())
        brats = cfg["dataset_name"] == "BraTS"
        if cfg["lambda_recon_y"] > 0:
            out["recon_y"] = (self.segmentation_loss_y_list(targets, y_list, mask) if brats
                              else self.recon_loss_y_list(targets, y_list, mask, cfg["p"]))
            total = total + cfg["lambda_recon_y"] * out["recon_y"]
        else:
            out["recon_y"] = z0
        if cfg["lambda_recon_y_fused"] > 0:
            out["recon_y_fused"] = (self.segmentation_loss_y(targets, y_fused) if brats
                                    else self.recon_loss_y(targets, y_fused, cfg["p"]))
            total = total + cfg["lambda_recon_y_fused"] * out["recon_y_fused"]
        else:
            out["recon_y_fused"] = z0
        if cfg["lambda_recon_x"] > 0:
            out["recon_x"] = self.recon_loss_x_list(xs, x_self, mask, cfg["p"])
            total = total + cfg["lambda_recon_x"] * out["recon_x"]
        else:
            out["recon_x"] = z0
        if cfg["lambda_recon_x_mix"] > 0:
            out["recon_x_mix"] = self.recon_loss_x_mix_list(xs, x_mix, mask, cfg["p"])
            total = total + cfg["lambda_recon_x_mix"] * out["recon_x_mix"]
        else:
            out["recon_x_mix"] = z0
        if cfg["lambda_kl"] > 0:
            out["kl"] = self.kl_loss_list_standard(mu, lv, mask)
            total = total + cfg["lambda_kl"] * out["kl"]
        else:
            out["kl"] = z0
        mu_new = None
        if cfg["lambda_latent_z"] > 0:
            si_new = self.compute_anatomy_encoding(x_self, mask_img)
            _, mu_new, _ = self.compute_modality_encoding(x_self, si_new, "train" if self.training else "test", None)
            out["latent_z"] = self.latent_z_loss(mu, mu_new, mask)
            total = total + cfg["lambda_latent_z"] * out["latent_z"]
        else:
            out["latent_z"] = z0
        if cfg["lambda_sim_s"] > 0:
            out["sim_s"] = self.similarity_s_loss(si, mask, pair)
            total = total + cfg["lambda_sim_s"] * out["sim_s"]
        else:
            out["sim_s"] = z0
        if cfg["lambda_sim_z"] > 0:
            out["sim_z"] = self.similarity_z_loss(zi, mask)
            total = total + cfg["lambda_sim_z"] * out["sim_z"]
        else:
            out["sim_z"] = z0
        out["all"] = total
        if keep:
            out["tensors"] = {"si": si, "zi": zi, "z_mean": mu, "z_log_var": lv, "x_fake": x_self,
                              "x_fake_mix": x_mix, "y_fake_list": y_list, "y_fake_fused": y_fused,
                              "z_mean_new": mu_new}
        return out


# ---------------------------------------------------------------------- helpers used by tests / bench
def param_keys(state: Dict[str, Tensor]) -> List[str]:
    """Keys that are nn.Parameters in the reference (everything except BN buffers)."""
    return [k for k in state if not (k.endswith("running_mean") or k.endswith("running_var")
                                     or k.endswith("num_batches_tracked"))]


def clone_state(state: Dict[str, Tensor], requires_grad: bool = True) -> Dict[str, Tensor]:
    out = {}
    pk = set(param_keys(state))
    for k, v in state.items():
        t = v.detach().clone().cpu()
        if t.is_floating_point():
            t = t.float()
        if k in pk and requires_grad:
            t.requires_grad_(True)
        out[k] = t
    return out


def clip_grad_norm(grads: Sequence[Optional[Tensor]], max_norm: float = 1.0) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (src/main_missing.py:272): in-place scale, returns total norm."""
    gs = [g for g in grads if g is not None]
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in gs]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in gs:
        g.mul_(coef)
    return total


def adam_amsgrad_step(params, grads, state, lr, step, betas=(0.9, 0.999), eps=1e-8, wd=1e-5):
    """torch.optim.Adam(amsgrad=True, weight_decay=1e-5) single step (src/main_missing.py:118).
    `state` is a dict name -> (m, v, vmax).  Params with grad None are skipped like torch does."""
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    with torch.no_grad():
        for name, p in params.items():
            g = grads.get(name)
            if g is None:
                continue
            g = g + wd * p
            m, v, vm = state.setdefault(name, (torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)))
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            torch.maximum(vm, v, out=vm)
            denom = (vm.sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / bc1)


def train_iteration(oracle: RDOracle, batch: dict, eps_list, pair, keep: bool = False):
    """Forward + backward + clip of one reference loop body (src/main_missing.py:165-272).
    Returns (losses dict of floats, grads dict name->tensor|None, grad_norm float, tensors|None)."""
    pk = param_keys(oracle.P)
    for k in pk:
        oracle.P[k].grad = None
    out = oracle.forward_losses(batch["inputs"], batch["targets"], batch["mask"], batch["mask_img"],
                                eps_list, pair, keep=keep)
    out["all"].backward()
    grads = {k: oracle.P[k].grad for k in pk}
    gn = clip_grad_norm(list(grads.values()), 1.0)
    losses = {k: float(v) for k, v in out.items() if k != "tensors"}
    return losses, grads, float(gn), out.get("tensors")
