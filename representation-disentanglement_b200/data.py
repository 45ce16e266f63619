"""Synthetic BraTS / NCANDA / ZeroDose shaped batches with the reference's batch-dict layout.

The dict keys and tensor shapes follow `ZeroDoseDataset.__getitem__` (reference src/util.py:471-566):
`inputs (B, M*(2*block+1), H, W)`, `targets (B, 1, H, W)`, `mask (B, M)`, `mask_img (B, H, W)`.
Value distribution: i.i.d. N(0,1) inside an elliptical "brain" support
(z-score data, src/data_preprocessing_BraTS.py:93), constant -10 outside (:95), all zeros for a
missing / dropped contrast (src/util.py:513,541), `mask_img = (inputs[:,0] == 0)` (src/util.py:564)
optionally with a zero border so the +100 softmax logit is exercised on part of the image,
BraTS targets = integer label blobs in {0,1,2,3}.  Random modality dropout follows
src/util.py:538-542 (p = 0.2, only if more than one contrast is present).
Everything is generated on the CPU with a seeded torch.Generator so that the oracle and the
device path can be fed identical tensors.
"""
import torch


def brain_support(H: int, W: int) -> torch.Tensor:
    yy = (torch.arange(H, dtype=torch.float32) - (H - 1) / 2) / (0.42 * H)
    xx = (torch.arange(W, dtype=torch.float32) - (W - 1) / 2) / (0.40 * W)
    return (yy[:, None] ** 2 + xx[None, :] ** 2) <= 1.0


def synthetic_batch(batch_size: int, modality_num: int = 4, block_size: int = 3, H: int = 160, W: int = 192,
                    seed: int = 10, dropoff: bool = False, missing=None, zero_border: int = 0,
                    labels: bool = True) -> dict:
    """Return a reference-layout batch dict of CPU fp32 tensors.

    missing: optional (B, M) 0/1 tensor / nested list; 0 = contrast absent (channels all zero).
    dropoff: apply the reference's random modality dropout on top of `missing`.
    zero_border: width of a border set to exactly 0 in every contrast (mask_img becomes 1 there).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    C = 2 * block_size + 1
    sup = brain_support(H, W)
    x = torch.randn(batch_size, modality_num * C, H, W, generator=g)
    x = torch.where(sup[None, None], x, torch.full_like(x, -10.0))
    if zero_border > 0:
        zb = zero_border
        x[:, :, :zb, :] = 0
        x[:, :, -zb:, :] = 0
        x[:, :, :, :zb] = 0
        x[:, :, :, -zb:] = 0
    mask = torch.ones(batch_size, modality_num)
    if missing is not None:
        mask = torch.as_tensor(missing, dtype=torch.float32).reshape(batch_size, modality_num).clone()
    if dropoff:
        for b in range(batch_size):
            if mask[b].sum() > 1 and torch.rand((), generator=g).item() > 0.8:
                present = torch.nonzero(mask[b] == 1).flatten()
                drop = present[torch.randint(len(present), (1,), generator=g).item()].item()
                mask[b, drop] = 0
    for b in range(batch_size):
        for m in range(modality_num):
            if mask[b, m] == 0:
                x[b, m * C:(m + 1) * C] = 0
    if labels:
        # label blobs: thresholded smooth noise inside the support, values in {0,1,2,3}
        n = torch.randn(batch_size, 1, H // 8, W // 8, generator=g)
        n = torch.nn.functional.interpolate(n, size=(H, W), mode="bilinear", align_corners=False)
        t = torch.zeros(batch_size, 1, H, W)
        t[n > 0.3] = 1
        t[n > 0.7] = 2
        t[n > 1.1] = 3
        t = t * sup[None, None].float()
    else:
        t = torch.zeros(batch_size, 1, H, W)
    mask_img = (x[:, 0] == 0).float()
    return {"inputs": x.contiguous(), "targets": t.contiguous(), "mask": mask, "mask_img": mask_img,
            "subj_id": ["synthetic_%04d" % b for b in range(batch_size)],
            "slice_idx": torch.full((batch_size,), 77, dtype=torch.long)}


def synthetic_eps(batch_size: int, modality_num: int, z_size: int, seed: int = 11):
    """The reparameterisation noise of `MultimodalModel.sample` (reference src/model.py:3159-3162),
    drawn once on the CPU so it can be injected into both the oracle and the device path."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return [torch.randn(batch_size, z_size, generator=g) for _ in range(modality_num)]
