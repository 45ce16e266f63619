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


# ======================================================================================================================
# Device-resident volume store + slab assembly: the data feed of the reference (ZeroDoseDataAll / ZeroDoseDataset,
# src/util.py:445-720) with the HDF5 volumes held in HBM (a BraTS training set — 369 subjects x (4 contrasts + labels) x
# 155 x 160 x 192 fp32 = 35 GB — fits in the 180 GB of one B200) and the per-sample NumPy work of `__getitem__`
# (7-slice window, missing / dropped contrasts, mask, mask_img, label remap) done by ONE kernel launch per batch.
# Only what needs the host stays there: the index lists, the shuffle (torch.randperm like DataLoader's RandomSampler) and the
# random dropoff decisions, drawn with the reference's own NumPy calls so that seeded runs drop the same contrasts.
import numpy as np

from . import kernels as K


class VolumeStore:
    """vols (S, M, D, H, W) fp32, present (S, M) uint8, tvols (S, D, H, W) fp32 | None, has_target (S) uint8 — all on `device`."""

    TARGET_KEY = {"ZeroDose": "PET", "BraTS": "seg", "Tau": "pet_nifti/fulldose"}

    def __init__(self, vols, present, tvols, has_target, subj_ids, contrast_list, dataset_name, brain_mask=None):
        self.vols, self.present, self.tvols, self.has_target = vols, present, tvols, has_target
        self.subj_ids, self.contrast_list, self.dataset_name = list(subj_ids), list(contrast_list), dataset_name
        self.brain_mask = brain_mask
        self.present_host = present.cpu().numpy().astype(np.int64)
        self.index = {s: k for k, s in enumerate(self.subj_ids)}

    @property
    def device(self):
        return self.vols.device

    @classmethod
    def from_dict(cls, data, subj_ids, contrast_list, dataset_name, device, brain_mask=None):
        """`data`: an h5py.File-like mapping 'subj/contrast' -> (H, W, D) array (the layout the reference's preprocessing writes,
        src/data_preprocessing_BraTS.py:85-97).  Volumes are transposed to (D, H, W) planes once, here."""
        S, M = len(subj_ids), len(contrast_list)
        first = None
        for s in subj_ids:
            for c in contrast_list:
                if s + "/" + c in data:
                    first = np.asarray(data[s + "/" + c])
                    break
            if first is not None:
                break
        if first is None:
            raise ValueError("VolumeStore.from_dict: no volume found for the given subjects / contrasts")
        H, W, D = first.shape
        vols = torch.zeros(S, M, D, H, W, dtype=torch.float32)
        present = torch.zeros(S, M, dtype=torch.uint8)
        tkey = cls.TARGET_KEY.get(dataset_name)
        tvols = torch.zeros(S, D, H, W, dtype=torch.float32) if tkey else None
        has_t = torch.zeros(S, dtype=torch.uint8)
        for k, s in enumerate(subj_ids):
            for m, c in enumerate(contrast_list):
                if s + "/" + c in data:
                    vols[k, m] = torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(data[s + "/" + c]), (2, 0, 1))).astype(np.float32))
                    present[k, m] = 1
            if tkey and s + "/" + tkey in data:
                tvols[k] = torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(data[s + "/" + tkey]), (2, 0, 1))).astype(np.float32))
                has_t[k] = 1
        bm = None
        if brain_mask is not None:
            bm = torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(brain_mask), (2, 0, 1))).astype(np.float32)).to(device)
        return cls(vols.to(device), present.to(device), tvols.to(device) if tvols is not None else None, has_t.to(device), subj_ids,
                   contrast_list, dataset_name, bm)

    @classmethod
    def synthetic(cls, subjects, contrast_list, dataset_name="BraTS", D=32, H=160, W=192, seed=10, device="cpu", missing_prob=0.0):
        """Random z-score-like volumes with the -10 background of src/data_preprocessing_BraTS.py:93-95 and label volumes."""
        g = torch.Generator().manual_seed(seed)
        M = len(contrast_list)
        sup = brain_support(H, W)
        vols = torch.randn(subjects, M, D, H, W, generator=g)
        vols = torch.where(sup[None, None, None], vols, torch.full_like(vols, -10.0))
        present = (torch.rand(subjects, M, generator=g) >= missing_prob).to(torch.uint8)
        present[:, 0] = 1
        tv = torch.randint(0, 5, (subjects, D, H // 8, W // 8), generator=g).float()
        tv = torch.nn.functional.interpolate(tv[:, None], size=(D, H, W), mode="nearest")[:, 0] * sup[None, None].float()
        has_t = torch.ones(subjects, dtype=torch.uint8)
        ids = ["synthetic_%04d" % k for k in range(subjects)]
        return cls(vols.to(device), present.to(device), tv.contiguous().to(device), has_t.to(device), ids, contrast_list, dataset_name)


class SlabLoader:
    """One of ZeroDoseDataAll's DataLoaders (src/util.py:699-706) over (subject, slice) pairs, batches assembled on the device.
    Yields the reference's batch dict (inputs, targets, mask, mask_img, subj_id, slice_idx) with DEVICE tensors."""

    def __init__(self, store: VolumeStore, subj_list, idx_list, batch_size, shuffle=False, dropoff=False, block_size=3, drop_last=False):
        self.store, self.B, self.shuffle, self.dropoff, self.block = store, int(batch_size), shuffle, dropoff, int(block_size)
        self.clamp_hi = 89 if store.dataset_name == "Tau" else 155                 # the constants of src/util.py:478-483
        D = store.vols.shape[2]
        keep = []
        for s, i in zip(subj_list, idx_list):
            sl = min(max(int(i), self.block), self.clamp_hi - self.block)
            # the reference's __getitem__ raises (-> sample skipped by nonechucks.SafeDataset) when the window leaves the volume or the
            # subject is unknown
            if str(s) in store.index and sl - self.block >= 0 and sl + self.block + 1 <= D:
                keep.append((store.index[str(s)], int(i)))
        self.items = keep
        self.drop_last = drop_last
        M = len(store.contrast_list)
        C = 2 * self.block + 1
        H, W = store.vols.shape[3], store.vols.shape[4]
        dev = store.device
        self._buf = {"inputs": torch.empty(self.B, M * C, H, W, device=dev), "targets": torch.empty(self.B, 1, H, W, device=dev),
                     "mask": torch.empty(self.B, M, device=dev), "mask_img": torch.empty(self.B, H, W, device=dev)}
        self._idx = torch.empty(3, self.B, dtype=torch.int32, device=dev)

    def __len__(self):
        n = len(self.items)
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def _drop_decisions(self, rows):
        """src/util.py:538-542 per sample, in sample order, with the reference's NumPy RNG calls."""
        drop = []
        for (s, _) in rows:
            d = -1
            mask = self.store.present_host[s]
            if self.dropoff and mask.sum() > 1:
                if np.random.rand() > 0.8:
                    d = int(np.random.choice(np.where(mask == 1)[0], 1)[0])
            drop.append(d)
        return drop

    def assemble(self, rows, drop=None):
        """rows: list of (store subject index, slice index); returns the batch dict (views of the loader's reusable buffers when the
        batch is full)."""
        b = len(rows)
        if drop is None:
            drop = self._drop_decisions(rows)
        host = torch.tensor([[s for s, _ in rows], [i for _, i in rows], drop], dtype=torch.int32)
        idx = self._idx[:, :b]
        idx.copy_(host, non_blocking=True)
        st = self.store
        out = {k: v[:b] for k, v in self._buf.items()}
        K.assemble_slabs(st.vols, st.present, st.tvols, st.has_target, st.brain_mask, idx[0].contiguous(), idx[1].contiguous(),
                         idx[2].contiguous(), out["inputs"], out["targets"], out["mask"], out["mask_img"], self.block,
                         st.dataset_name == "BraTS", self.clamp_hi)
        out["subj_id"] = [st.subj_ids[s] for s, _ in rows]
        out["slice_idx"] = torch.tensor([min(max(i, self.block), self.clamp_hi - self.block) for _, i in rows], dtype=torch.long)
        return out

    def __iter__(self):
        n = len(self.items)
        if self.shuffle:       # torch.utils.data.RandomSampler: a generator seeded from the default generator, then randperm
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        for k in range(0, n, self.B):
            rows = [self.items[j] for j in order[k:k + self.B]]
            if len(rows) < self.B and self.drop_last:
                return
            yield self.assemble(rows)
