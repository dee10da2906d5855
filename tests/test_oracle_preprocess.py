"""Oracle preprocessing against the reference's known answers:
tests/nn/data/test_normalization.py:17-62 (ensure_float / ensure_grayscale / ensure_rgb),
tests/nn/data/test_resizing.py:13-66 (find_padding_for_stride, pad_to_stride, resize_image),
tests/nn/test_inference.py:399-497 (InferenceLayer.preprocess with an identity network)."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle import preprocess as opre


def test_ensure_float_gray_rgb():
    assert opre.ensure_float(np.zeros((2, 2), np.uint8)).dtype == np.float32
    assert opre.ensure_float(np.zeros((2, 2), np.float32)).dtype == np.float32
    assert_array_equal(opre.ensure_grayscale(np.full((2, 2, 3), 255, np.uint8)), np.full((2, 2, 1), 255, np.uint8))
    assert_array_equal(opre.ensure_grayscale(np.full((2, 2, 1), 255, np.uint8)), np.full((2, 2, 1), 255, np.uint8))
    assert_allclose(opre.ensure_grayscale(np.ones((2, 2, 3), np.float32)), np.ones((2, 2, 1), np.float32), atol=1e-4)
    assert_array_equal(opre.ensure_rgb(np.full((2, 2, 3), 255, np.uint8)), np.full((2, 2, 3), 255, np.uint8))
    assert_array_equal(opre.ensure_rgb(np.full((2, 2, 1), 255, np.uint8)), np.full((2, 2, 3), 255, np.uint8))


def test_pad_to_stride_and_resize():
    want = np.asarray([[1, 1, 1, 1, 1, 0]] * 3 + [[0] * 6], np.float32)[..., None]
    assert_array_equal(opre.pad_to_stride(np.ones((3, 5, 1), np.float32), 2), want)
    assert opre.pad_to_stride(np.ones((3, 5, 1), np.uint8), 2).dtype == np.uint8
    assert opre.pad_to_stride(np.ones((4, 4, 1), np.float32), 2).shape == (4, 4, 1)
    # find_padding_for_stride(127, 129, 32) == (1, 31); (128, 128, 32) == (0, 0)
    assert opre.pad_to_stride(np.ones((127, 129, 1), np.uint8), 32).shape == (128, 160, 1)
    assert opre.pad_to_stride(np.ones((128, 128, 1), np.uint8), 32).shape == (128, 128, 1)
    r = opre.resize_image(np.ones((4, 8, 1), np.uint8), 0.5)
    assert r.shape == (2, 4, 1) and r.dtype == np.uint8
    assert opre.resize_image(np.ones((4, 8, 1), np.float32), 0.5).dtype == np.float32


def test_inference_layer_preprocess():
    full = np.full((1, 4, 4, 1), 255, np.uint8)
    out = opre.preprocess(full, ensure_gray=True)
    assert out.dtype == np.float32 and out.shape == (1, 4, 4, 1) and np.all(out == 1.0)
    out = opre.preprocess(full, ensure_gray=True, do_float=False)                       # ensure_float=False keeps uint8
    assert out.dtype == np.uint8 and np.all(out == 255)
    out = opre.preprocess(np.full((1, 4, 4, 3), 255, np.uint8), ensure_gray=True)       # rgb -> grayscale
    assert out.shape == (1, 4, 4, 1) and np.all(out == 1.0)
    out = opre.preprocess(full, ensure_gray=False)                                      # grayscale -> rgb
    assert out.shape == (1, 4, 4, 3) and np.all(out == 1.0)
    out = opre.preprocess(np.full((1, 8, 8, 1), 255, np.uint8), ensure_gray=True, input_scale=0.5)
    assert out.dtype == np.float32 and out.shape == (1, 4, 4, 1) and np.all(out == 1.0)
    assert opre.preprocess(np.full((1, 3, 3, 1), 255, np.uint8), True, 1.0, 2).shape == (1, 4, 4, 1)
    assert opre.preprocess(np.full((1, 6, 6, 1), 255, np.uint8), True, 0.5, 2).shape == (1, 4, 4, 1)
