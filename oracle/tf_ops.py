"""NumPy float32 restatements of the TensorFlow-2.7 kernels the hot path dispatches.

Oracle / test infrastructure only (see oracle/__init__.py).  Each function cites the
reference call site; the kernel semantics themselves live in TensorFlow (pinned
``tensorflow==2.7.0``, reference ``environment.yml:39``), which is not vendored under
/root/reference, so they are restated from the published TF 2.7 kernel definitions.

All arithmetic is performed in float32 without fused multiply-add, op for op.
"""
import numpy as np

F32 = np.float32


def dilation2d_nms_max(cms: np.ndarray) -> np.ndarray:
    """``tf.nn.dilation2d`` with the kernel [[0,0,0],[0,-1,0],[0,0,0]], SAME padding.

    Reference call site: sleap/nn/peak_finding.py:274-287.  TF semantics: output =
    max over the *in-bounds* taps of (input + filter); out-of-image taps are skipped.

    Args:
        cms: (B, H, W, C) float32.
    Returns:
        (B, H, W, C) float32 = max(8 in-bounds neighbours, centre - 1).
    """
    cms = np.asarray(cms, dtype=F32)
    B, H, W, C = cms.shape
    out = cms - F32(1.0)
    pad = np.full((B, H + 2, W + 2, C), -np.inf, dtype=F32)
    pad[:, 1:-1, 1:-1, :] = cms
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            if dy == 1 and dx == 1:
                continue
            out = np.maximum(out, pad[:, dy:dy + H, dx:dx + W, :])
    return out


def make_centered_bboxes(centroids: np.ndarray, box_height: int, box_width: int) -> np.ndarray:
    """sleap/nn/data/instance_cropping.py:124-166.  centroids (N,2) xy -> (N,4) y1x1y2x2."""
    centroids = np.asarray(centroids, dtype=F32).reshape(-1, 2)
    delta = np.array(
        [[-box_height + 1, -box_width + 1, box_height - 1, box_width - 1]], dtype=F32
    ) * F32(0.5)
    return (centroids[:, [1, 0, 1, 0]] + delta).astype(F32)


def normalize_bboxes(bboxes: np.ndarray, image_height: int, image_width: int) -> np.ndarray:
    """sleap/nn/data/instance_cropping.py:58-90: divide by (H-1, W-1, H-1, W-1) in f32."""
    factor = np.array([[image_height, image_width, image_height, image_width]], dtype=F32) - F32(1)
    return (np.asarray(bboxes, dtype=F32) / factor).astype(F32)


def crop_and_resize_bilinear(images, boxes, box_inds, crop_h, crop_w, extrapolation=0.0):
    """``tf.image.crop_and_resize(method="bilinear")`` (TF 2.7 CPU kernel), float32.

    Reference call site: sleap/nn/peak_finding.py:180-186.

    Args:
        images: (B, H, W, C) any real dtype (cast to f32 per pixel, as TF does).
        boxes: (N, 4) normalised y1, x1, y2, x2 float32.
        box_inds: (N,) int.
    Returns:
        (N, crop_h, crop_w, C) float32.
    """
    images = np.asarray(images)
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    box_inds = np.asarray(box_inds).reshape(-1)
    B, H, W, C = images.shape
    N = boxes.shape[0]
    out = np.full((N, crop_h, crop_w, C), F32(extrapolation), dtype=F32)
    Hm1 = F32(H - 1)
    Wm1 = F32(W - 1)
    for b in range(N):
        y1, x1, y2, x2 = [F32(v) for v in boxes[b]]
        bi = int(box_inds[b])
        img = images[bi]
        height_scale = (F32(F32(y2 - y1) * Hm1) / F32(crop_h - 1)) if crop_h > 1 else F32(0)
        width_scale = (F32(F32(x2 - x1) * Wm1) / F32(crop_w - 1)) if crop_w > 1 else F32(0)
        for y in range(crop_h):
            if crop_h > 1:
                in_y = F32(F32(y1 * Hm1) + F32(F32(y) * height_scale))
            else:
                in_y = F32(F32(F32(0.5) * F32(y1 + y2)) * Hm1)
            if in_y < 0 or in_y > Hm1:
                continue
            top = int(np.floor(in_y))
            bot = int(np.ceil(in_y))
            y_lerp = F32(in_y - F32(top))
            for x in range(crop_w):
                if crop_w > 1:
                    in_x = F32(F32(x1 * Wm1) + F32(F32(x) * width_scale))
                else:
                    in_x = F32(F32(F32(0.5) * F32(x1 + x2)) * Wm1)
                if in_x < 0 or in_x > Wm1:
                    continue
                left = int(np.floor(in_x))
                right = int(np.ceil(in_x))
                x_lerp = F32(in_x - F32(left))
                tl = img[top, left].astype(F32)
                tr = img[top, right].astype(F32)
                bl = img[bot, left].astype(F32)
                br = img[bot, right].astype(F32)
                t = (tl + (tr - tl) * x_lerp).astype(F32)
                bt = (bl + (br - bl) * x_lerp).astype(F32)
                out[b, y, x] = (t + (bt - t) * y_lerp).astype(F32)
    return out


def crop_bboxes(images, bboxes, sample_inds):
    """sleap/nn/peak_finding.py:135-190 (box size from the FIRST box; cast back to dtype)."""
    images = np.asarray(images)
    bboxes = np.asarray(bboxes, dtype=F32).reshape(-1, 4)
    y1x1 = bboxes[0, 0:2]
    y2x2 = bboxes[0, 2:4]
    box_size = np.round((y2x2 - y1x1) + F32(1)).astype(np.int32)  # np.round == round-half-even == tf.round
    H, W = images.shape[1], images.shape[2]
    nb = normalize_bboxes(bboxes, H, W)
    crops = crop_and_resize_bilinear(images, nb, sample_inds, int(box_size[0]), int(box_size[1]))
    if images.dtype == np.uint8:
        # tf.cast(float32 -> uint8) truncates toward zero (values are within [0,255]).
        return np.trunc(crops).astype(np.uint8)
    return crops.astype(images.dtype)


def resize_bilinear_half_pixel(images: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``tf.image.resize(method="bilinear", antialias=False)`` (TF2: half-pixel centres).

    Call sites: sleap/nn/data/resizing.py:96-106; Keras ``UpSampling2D(bilinear)`` at
    sleap/nn/architectures/encoder_decoder.py:335-339.
    """
    images = np.asarray(images, dtype=F32)
    B, H, W, C = images.shape

    def weights(out_size, in_size):
        scale = F32(in_size) / F32(out_size)
        i = np.arange(out_size, dtype=F32)
        src = ((i + F32(0.5)) * scale - F32(0.5)).astype(F32)
        src_f = np.floor(src)
        lower = np.maximum(src_f.astype(np.int64), 0)
        upper = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
        lerp = (src - src_f).astype(F32)
        return lower, upper, lerp

    ylo, yhi, yl = weights(out_h, H)
    xlo, xhi, xl = weights(out_w, W)
    top_rows = images[:, ylo]
    bot_rows = images[:, yhi]
    xl_ = xl.reshape(1, 1, -1, 1)
    yl_ = yl.reshape(1, -1, 1, 1)
    tl, tr = top_rows[:, :, xlo], top_rows[:, :, xhi]
    bl, br = bot_rows[:, :, xlo], bot_rows[:, :, xhi]
    top = (tl + (tr - tl) * xl_).astype(F32)
    bot = (bl + (br - bl) * xl_).astype(F32)
    return (top + (bot - top) * yl_).astype(F32)


def tf_linspace(start, stop, num: int) -> np.ndarray:
    """``tf.linspace(start, stop, num, axis=-1 appended)`` in f32.

    TF: delta = (stop - start) / (num - 1); values = start + delta * range(num - 1),
    with the last value set to exactly ``stop`` (math_ops.linspace_nd).
    Call site: sleap/nn/paf_grouping.py:192.
    Returns array of shape start.shape + (num,).
    """
    start = np.asarray(start, dtype=F32)
    stop = np.asarray(stop, dtype=F32)
    if num == 1:
        return start[..., None].astype(F32)
    delta = ((stop - start) / F32(num - 1)).astype(F32)
    idx = np.arange(num - 1, dtype=F32)
    body = (start[..., None] + (delta[..., None] * idx).astype(F32)).astype(F32)
    return np.concatenate([body, stop[..., None]], axis=-1).astype(F32)
