"""TEST INFRASTRUCTURE — CPU restatement of the reference's evaluation metrics (src/util.py:935-992).

`compute_segmentation_metrics*` is plain NumPy in the reference and is restated line by line (and, in the build container,
checked against the reference function itself by oracle/make_golden.py).  `compute_reconstruction_metrics_single` calls
`skimage.metrics.{mean_squared_error, peak_signal_noise_ratio, structural_similarity}` — scikit-image is a third-party dependency
that is ABSENT from this image and from /root/reference (the reference pins no version: it has no requirements file).  Their
published algorithms (scikit-image >= 0.19, `skimage/metrics/simple_metrics.py`, `_structural_similarity.py`) are restated here
with scipy.ndimage.uniform_filter, which is what skimage itself calls.  PARITY UNPINNED against skimage: no skimage run is
possible here; the restatement is anchored on the reference's call site (float32 images shifted to min 0, data_range =
max(target - min), default win_size 7, uniform window, sample covariance, K1 0.01, K2 0.03)."""
import numpy as np
from scipy.ndimage import uniform_filter


def mean_squared_error(a, b):
    """skimage.metrics.mean_squared_error: np.mean((a - b) ** 2, dtype=np.float64) on the float32 images."""
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return float(np.mean((a - b) ** 2, dtype=np.float64))


def peak_signal_noise_ratio(a, b, data_range):
    err = mean_squared_error(a, b)
    with np.errstate(divide="ignore"):
        return float(10 * np.log10((float(data_range) ** 2) / err))


def structural_similarity(im1, im2, data_range, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity with its defaults (gaussian_weights False, use_sample_covariance True), float32 in."""
    im1, im2 = np.asarray(im1, dtype=np.float32), np.asarray(im2, dtype=np.float32)
    ndim = im1.ndim
    NP = win_size ** ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    R = data_range
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def compute_reconstruction_metrics_single(target, pred):
    """src/util.py:956-978."""
    target = target - target.min()
    pred = pred - pred.min()
    rng = target.max()
    return {"ssim": structural_similarity(target, pred, rng), "rmse": mean_squared_error(target, pred),
            "psnr": peak_signal_noise_ratio(target, pred, rng)}


def compute_reconstruction_metrics(target, pred):
    """src/util.py:935-944: channel 0 of every image."""
    out = {"ssim": [], "psnr": [], "rmse": []}
    for i in range(target.shape[0]):
        m = compute_reconstruction_metrics_single(target[i, 0], pred[i, 0])
        for k in out:
            out[k].append(m[k])
    return out


def compute_segmentation_metrics_single(target, pred):
    """src/util.py:980-992 (class i + 1 of the target against prediction channel i, +1 smoothing)."""
    if target.shape[0] == 1:
        target = target.squeeze(0)
    dice_list, iou_list = [], []
    for i in range(3):
        inter = np.logical_and(target == i + 1, pred[i] > 0.5)
        union = np.logical_or(target == i + 1, pred[i] > 0.5)
        dice_list.append((2.0 * inter.sum() + 1) / ((target == i + 1).sum() + (pred[i] > 0.5).sum() + 1))
        iou_list.append((np.sum(inter) + 1) / (np.sum(union) + 1))
    return {"dice": np.mean(dice_list), "iou": np.mean(iou_list)}


def compute_segmentation_metrics(target, pred):
    """src/util.py:946-954."""
    out = {"dice": [], "iou": []}
    for i in range(target.shape[0]):
        m = compute_segmentation_metrics_single(target[i], pred[i])
        out["dice"].append(m["dice"])
        out["iou"].append(m["iou"])
    return out


def assemble_sample(data, subj_id, slice_idx, contrast_list, block_size, dataset_name, image_size=(160, 192), drop_idx=None,
                    brain_mask=None):
    """ZeroDoseDataset.__getitem__ (src/util.py:471-566) for one sample from a dict `data[subj/contrast] -> (H, W, D) array`;
    `drop_idx` = the contrast the random dropoff removed (None: no drop), so that the RNG stays with the caller."""
    bs = block_size
    if slice_idx < bs:
        slice_idx = bs
    hi = 89 if dataset_name == "Tau" else 155
    if slice_idx > hi - bs:
        slice_idx = hi - bs
    imgs, mask = [], []
    for c in contrast_list:
        key = subj_id + "/" + c
        if key in data:
            imgs.append(np.asarray(data[key])[:, :, slice_idx - bs:slice_idx + bs + 1])
            mask.append(1)
        else:
            imgs.append(np.zeros((image_size[0], image_size[1], 2 * bs + 1)))
            mask.append(0)
    mask = np.array(mask)
    inputs = np.concatenate(imgs, 2).astype(np.float64).copy()
    tkey = {"ZeroDose": "/PET", "BraTS": "/seg", "Tau": "/pet_nifti/fulldose"}.get(dataset_name)
    if tkey is not None and subj_id + tkey in data:
        targets = np.asarray(data[subj_id + tkey])[:, :, slice_idx:slice_idx + 1].astype(np.float64).copy()
        if dataset_name == "BraTS":
            targets[targets == 4] = 3.0
    else:
        targets = np.zeros((image_size[0], image_size[1], 1))
    if drop_idx is not None:
        inputs[:, :, drop_idx * (2 * bs + 1):(drop_idx + 1) * (2 * bs + 1)] = 0
        mask[drop_idx] = 0
    if brain_mask is not None:
        bm = brain_mask[:, :, slice_idx - bs:slice_idx + bs + 1]
        inputs = inputs * np.tile(bm, (1, 1, len(contrast_list)))
        targets = targets * brain_mask[:, :, slice_idx:slice_idx + 1]
    inputs = np.transpose(inputs, (2, 0, 1))
    targets = np.transpose(targets, (2, 0, 1))
    mask_img = (inputs[0] == 0).astype(float)
    return {"inputs": inputs, "targets": targets, "subj_id": subj_id, "slice_idx": slice_idx, "mask": mask, "mask_img": mask_img}
