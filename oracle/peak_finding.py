"""Oracle restatement of sleap/nn/peak_finding.py (NumPy float32).  Test infrastructure only.

Function names, argument order and return tuples mirror the reference module so the
parity tests read like the reference's own tests (tests/nn/test_peak_finding.py).
"""
import numpy as np

from .tf_ops import (F32, crop_bboxes, dilation2d_nms_max, make_centered_bboxes)


def find_offsets_local_direction(centered_patches, delta=0.25):
    """sleap/nn/peak_finding.py:78-132.  patches (N,3,3,1) -> offsets (N,2) (dx, dy)."""
    p = np.asarray(centered_patches, dtype=F32)
    dx = p[:, 1, 2, :] - p[:, 1, 0, :]
    dy = p[:, 2, 1, :] - p[:, 0, 1, :]
    off = np.sign(np.stack([dx, dy], axis=1)[..., 0]).astype(F32) * F32(delta)
    return off.astype(F32)


def integral_regression(cms, xv, yv):
    """sleap/nn/peak_finding.py:311-334.  cms (N,h,w,C) -> x_hat, y_hat (N,C)."""
    cms = np.asarray(cms, dtype=F32)
    xv = np.asarray(xv, dtype=F32)
    yv = np.asarray(yv, dtype=F32)
    z = cms.sum(axis=(1, 2), dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        x_hat = (xv.reshape(1, 1, -1, 1) * cms).astype(F32).sum(axis=(1, 2), dtype=F32) / z
        y_hat = (yv.reshape(1, -1, 1, 1) * cms).astype(F32).sum(axis=(1, 2), dtype=F32) / z
    return x_hat.astype(F32), y_hat.astype(F32)


def find_global_peaks_rough(cms, threshold=0.1):
    """sleap/nn/peak_finding.py:193-246."""
    cms = np.asarray(cms, dtype=F32)
    B, H, W, C = cms.shape
    max_img_rows = cms.max(axis=2)                     # (B, H, C)
    argmax_rows = np.argmax(max_img_rows, axis=1).reshape(-1)   # first index on ties
    max_img_cols = cms.max(axis=1)                     # (B, W, C)
    argmax_cols = np.argmax(max_img_cols, axis=1).reshape(-1)
    total = argmax_cols.shape[0]
    sample_subs = np.arange(total) // C
    channel_subs = np.arange(total) % C
    peak_vals = cms[sample_subs, argmax_rows, argmax_cols, channel_subs].reshape(-1, C)
    peak_points = np.stack([argmax_cols, argmax_rows], axis=-1).astype(F32).reshape(-1, C, 2)
    peak_points = np.where((peak_vals < F32(threshold))[..., None], F32(np.nan), peak_points)
    return peak_points.astype(F32), peak_vals.astype(F32)


def find_local_peaks_rough(cms, threshold=0.2):
    """sleap/nn/peak_finding.py:249-308.  Order = tf.where row-major (sample, y, x, channel)."""
    cms = np.asarray(cms, dtype=F32)
    max_img = dilation2d_nms_max(cms)
    mask = (cms > max_img) & (cms > F32(threshold))
    subs = np.argwhere(mask)                            # row-major, like tf.where
    peak_vals = cms[mask].astype(F32)
    peak_points = subs[:, [2, 1]].astype(F32)
    return (peak_points, peak_vals, subs[:, 0].astype(np.int32), subs[:, 3].astype(np.int32))


def _refine(cms, rough_peaks, box_sample_inds, refinement, integral_patch_size):
    """Shared refinement (peak_finding.py:386-420 / 503-532): crops on (B*C,H,W,1) maps."""
    if refinement == "integral":
        crop_size = integral_patch_size
    elif refinement == "local":
        crop_size = 3
    else:
        return None
    B, H, W, C = cms.shape
    bboxes = make_centered_bboxes(rough_peaks, crop_size, crop_size)
    flat = np.transpose(cms, (0, 3, 1, 2)).reshape(B * C, H, W, 1)
    crops = crop_bboxes(flat, bboxes, box_sample_inds)
    if refinement == "integral":
        gv = np.arange(crop_size, dtype=F32) - F32((crop_size - 1) / 2)
        dx, dy = integral_regression(crops, gv, gv)
        return np.concatenate([dx, dy], axis=1).astype(F32)
    return find_offsets_local_direction(crops, 0.25)


def find_global_peaks(cms, threshold=0.2, refinement=None, integral_patch_size=5):
    """sleap/nn/peak_finding.py:337-420."""
    cms = np.asarray(cms, dtype=F32)
    rough, vals = find_global_peaks_rough(cms, threshold)
    if refinement is None or np.all(np.isnan(rough)):
        return rough, vals
    if refinement not in ("integral", "local"):
        return rough, vals
    B, H, W, C = cms.shape
    flat_peaks = rough.reshape(B * C, 2).copy()
    valid_idx = np.nonzero(~np.isnan(flat_peaks[:, 0]))[0]
    offsets = _refine(cms, flat_peaks[valid_idx], valid_idx, refinement, integral_patch_size)
    flat_peaks[valid_idx] = (flat_peaks[valid_idx] + offsets).astype(F32)
    return flat_peaks.reshape(B, C, 2).astype(F32), vals


def find_local_peaks(cms, threshold=0.2, refinement=None, integral_patch_size=5):
    """sleap/nn/peak_finding.py:451-532."""
    cms = np.asarray(cms, dtype=F32)
    rough, vals, sample_inds, channel_inds = find_local_peaks_rough(cms, threshold)
    if rough.shape[0] == 0 or refinement not in ("integral", "local"):
        return rough, vals, sample_inds, channel_inds
    C = cms.shape[3]
    box_inds = sample_inds.astype(np.int64) * C + channel_inds
    offsets = _refine(cms, rough, box_inds, refinement, integral_patch_size)
    return (rough + offsets).astype(F32), vals, sample_inds, channel_inds


def find_global_peaks_with_offsets(cms, offsets, threshold=0.2):
    """sleap/nn/peak_finding.py:566-643.  offsets (B,H,W,2C) -> reshaped (B,H,W,C,2)."""
    cms = np.asarray(cms, dtype=F32)
    rough, vals = find_global_peaks_rough(cms, threshold)
    if np.all(np.isnan(rough)):
        return rough, vals
    B, H, W, C = cms.shape
    off = np.asarray(offsets, dtype=F32).reshape(B, H, W, -1, 2)
    out = rough.copy()
    for s in range(B):
        for c in range(C):
            if not np.isnan(rough[s, c, 0]):
                x, y = int(rough[s, c, 0]), int(rough[s, c, 1])
                out[s, c] = rough[s, c] + off[s, y, x, c]
    return out.astype(F32), vals


def find_local_peaks_with_offsets(cms, offsets, threshold=0.2):
    """sleap/nn/peak_finding.py:646-707."""
    cms = np.asarray(cms, dtype=F32)
    rough, vals, sample_inds, channel_inds = find_local_peaks_rough(cms, threshold)
    if rough.shape[0] == 0:
        return rough, vals, sample_inds, channel_inds
    B, H, W, C = cms.shape
    off = np.asarray(offsets, dtype=F32).reshape(B, H, W, -1, 2)
    po = off[sample_inds, rough[:, 1].astype(np.int32), rough[:, 0].astype(np.int32), channel_inds]
    return (rough + po).astype(F32), vals, sample_inds, channel_inds
