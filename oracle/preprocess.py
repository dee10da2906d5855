"""Oracle restatement of the preprocessing ops (NumPy float32).  Test infrastructure only.

Sources: sleap/nn/data/normalization.py:34-114, sleap/nn/data/resizing.py:10-106,
sleap/nn/inference.py:940-967 (InferenceLayer.preprocess).  TF kernel semantics restated
from TF 2.7 (convert_image_dtype, rgb_to_grayscale, resize, pad).
"""
import numpy as np

from .tf_ops import F32, resize_bilinear_half_pixel


def ensure_float(image):
    """normalization.py:34-49: convert_image_dtype(u8 -> f32) = cast * (1/255) in f32."""
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return (image.astype(F32) * F32(1.0 / 255)).astype(F32)   # scale = 1.0 / dtype.max
    return image.astype(F32)


def ensure_grayscale(image):
    """normalization.py:81-95.  tf.image.rgb_to_grayscale keeps the input dtype:
    u8 -> f32 (x/255), dot [0.2989, 0.5870, 0.1140], f32 -> u8 via trunc(x * 255.5)."""
    image = np.asarray(image)
    if image.shape[-1] != 3:
        return image
    w = np.array([0.2989, 0.5870, 0.1140], dtype=F32)
    flt = ensure_float(image)
    gray = (flt[..., 0] * w[0] + flt[..., 1] * w[1] + flt[..., 2] * w[2]).astype(F32)[..., None]
    if image.dtype == np.uint8:
        return np.clip(np.trunc(gray * F32(255.5)), 0, 255).astype(np.uint8)
    return gray


def ensure_rgb(image):
    """normalization.py:98-114."""
    image = np.asarray(image)
    if image.shape[-1] == 1:
        return np.repeat(image, 3, axis=-1)
    return image


def resize_image(image, scale):
    """resizing.py:71-106: new size = int(W*s), int(H*s) (truncate), bilinear, no antialias."""
    image = np.asarray(image)
    H, W = image.shape[-3], image.shape[-2]
    new_w = int(F32(W) * F32(scale))
    new_h = int(F32(H) * F32(scale))
    if image.ndim == 3:                      # the reference accepts a single (H, W, C) image as well
        return resize_image(image[None], scale)[0]
    out = resize_bilinear_half_pixel(image.astype(F32), new_h, new_w)
    if image.dtype == np.uint8:
        return np.trunc(out).astype(np.uint8)
    return out.astype(image.dtype)


def pad_to_stride(image, max_stride):
    """resizing.py:34-68: zero-pad bottom/right to a multiple of max_stride."""
    image = np.asarray(image)
    H, W = image.shape[-3], image.shape[-2]
    pb = (max_stride - H % max_stride) % max_stride
    pr = (max_stride - W % max_stride) % max_stride
    if pb == 0 and pr == 0:
        return image
    pads = [(0, 0)] * (image.ndim - 3) + [(0, pb), (0, pr), (0, 0)]
    return np.pad(image, pads, mode="constant", constant_values=0)


def preprocess(imgs, ensure_gray, input_scale=1.0, pad_stride=1, do_float=True, resize_img=True):
    """inference.py:940-967."""
    imgs = ensure_grayscale(imgs) if ensure_gray else ensure_rgb(imgs)
    if do_float:
        imgs = ensure_float(imgs)
    if resize_img and input_scale != 1.0:
        imgs = resize_image(imgs, input_scale)
    if pad_stride > 1:
        imgs = pad_to_stride(imgs, pad_stride)
    return imgs
