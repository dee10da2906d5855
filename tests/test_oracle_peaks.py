"""Pins the CPU oracle against the reference's own known-answer vectors (no GPU)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import cases_peaks
from oracle import peak_finding as opf
from oracle import tf_ops


@pytest.mark.parametrize("case", cases_peaks.ALL, ids=lambda f: f.__name__)
def test_oracle_peak_case(case):
    case(opf)


def test_bboxes_known_answers():
    # reference tests/nn/data/test_instance_cropping.py:14-60
    bb = tf_ops.normalize_bboxes(np.array([[0, 0, 3, 3]], np.float32), 9, 9)
    assert_allclose(bb, [[0, 0, 0.375, 0.375]])
    assert_array_equal(tf_ops.make_centered_bboxes(np.array([[1, 1]], np.float32), 3, 3), [[0, 0, 2, 2]])
    assert_array_equal(tf_ops.make_centered_bboxes(np.array([[2, 2]], np.float32), 4, 4), [[0.5, 0.5, 3.5, 3.5]])
    img = np.arange(81, dtype=np.float32).reshape(1, 9, 9, 1)
    crop = tf_ops.crop_bboxes(img, np.array([[0, 0, 2, 2]], np.float32), [0])
    assert_array_equal(crop[0, :, :, 0], img[0, :3, :3, 0])
    bb = tf_ops.make_centered_bboxes(np.array([[464.42838, 550.14276]], np.float32), 100, 100)
    crop = tf_ops.crop_bboxes(np.zeros((1, 1024, 1024, 1), np.float32), bb, [0])
    assert crop.shape == (1, 100, 100, 1)


def test_nms_border_and_plateau():
    cms = np.zeros((1, 5, 5, 1), np.float32)
    cms[0, 0, 0, 0] = 0.9       # corner: out-of-image taps skipped -> still a peak
    cms[0, 2, 2, 0] = 0.5
    cms[0, 2, 3, 0] = 0.5       # plateau: strict '>' -> neither is a peak
    pts, vals, si, ci = opf.find_local_peaks_rough(cms, threshold=0.2)
    assert_array_equal(pts, [[0, 0]])
    assert_array_equal(vals, np.array([0.9], np.float32))


def test_global_tie_and_threshold_asymmetry():
    cms = np.zeros((1, 4, 4, 2), np.float32)
    cms[0, 1, 2, 0] = 0.2      # == threshold: global keeps it (strict '<'), local drops it (strict '>')
    cms[0, 1, 2, 1] = 0.7
    cms[0, 3, 0, 1] = 0.7      # tie -> first index on each axis: row 1, col 0 -> value there is 0
    gp, gv = opf.find_global_peaks_rough(cms, threshold=0.2)
    assert_array_equal(gp[0, 0], [2, 1])
    assert np.all(np.isnan(gp[0, 1])) and gv[0, 1] == 0
    lp = opf.find_local_peaks_rough(cms, threshold=0.2)
    assert_array_equal(lp[3], [1, 1])
