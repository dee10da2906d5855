"""sleap_b200.nn.identity (host logic) against the reference's known-answer vectors, tests/nn/test_inference_identity.py."""
import numpy as np
from numpy.testing import assert_array_equal

from sleap_b200.nn.identity import classify_peaks_from_maps, classify_peaks_from_vectors, group_class_peaks


def test_group_class_peaks():
    """tests/nn/test_inference_identity.py:15-38."""
    peak_class_probs = np.array([[0.1, 0.9], [0.9, 0.1], [0.95, 0.05], [0.8, 0.2], [0.9, 0.1], [0.85, 0.15], [0.1, 0.9]])
    peak_sample_inds = np.array([0, 0, 0, 0, 1, 1, 1])
    peak_channel_inds = np.array([0, 0, 1, 1, 0, 0, 0])
    peak_inds, class_inds = group_class_peaks(peak_class_probs, peak_sample_inds, peak_channel_inds, n_samples=2, n_channels=2)
    assert_array_equal(peak_inds, [0, 1, 2, 4, 6])
    assert_array_equal(class_inds, [1, 0, 0, 0, 1])


def test_classify_peaks_from_maps():
    """tests/nn/test_inference_identity.py:41-79."""
    peak_class_probs = np.array([[0.1, 0.9], [0.91, 0.09], [0.95, 0.05], [0.8, 0.2], [0.92, 0.08], [0.85, 0.15], [0.07, 0.93]])
    peak_sample_inds = np.array([0, 0, 0, 0, 1, 1, 1])
    peak_channel_inds = np.array([0, 0, 1, 1, 0, 0, 0])
    peak_points = np.arange(7 * 2, dtype=np.float32).reshape(7, 2)
    peak_vals = np.ones([7], np.float32)
    class_maps = np.zeros([2, 14, 14, 2], dtype="float32")
    for s, (x, y), pr in zip(peak_sample_inds, peak_points, peak_class_probs):
        class_maps[s, int(y), int(x), :] = pr
    points, point_vals, class_probs = classify_peaks_from_maps(class_maps, peak_points, peak_vals, peak_sample_inds,
                                                               peak_channel_inds, n_channels=2)
    assert points.shape == (2, 2, 2, 2) and point_vals.shape == (2, 2, 2) and class_probs.shape == (2, 2, 2)
    assert_array_equal(points[0][0], peak_points[[1, 2]])
    assert_array_equal(points[0][1], [peak_points[0], [np.nan, np.nan]])
    assert_array_equal(points[1][0], [peak_points[4], [np.nan, np.nan]])
    assert_array_equal(points[1][1], [peak_points[6], [np.nan, np.nan]])
    assert np.isclose(class_probs[0, 0, 0], 0.91) and np.isclose(class_probs[1, 1, 0], 0.93) and np.isnan(class_probs[0, 1, 1])
    assert point_vals[0, 0, 0] == 1.0 and np.isnan(point_vals[1, 0, 1])


def test_classify_peaks_from_vectors_and_empty():
    """Top-down grouping (sleap/nn/identity.py:182-254): crops of one sample compete for the classes; the crop that is not
    the best match of its class is dropped; no peaks -> all NaN."""
    pts = np.arange(3 * 2 * 2, dtype=np.float32).reshape(3, 2, 2)
    vals = np.ones((3, 2), np.float32)
    probs = np.array([[0.8, 0.2], [0.3, 0.7], [0.9, 0.1]], np.float32)
    points, point_vals, class_probs = classify_peaks_from_vectors(pts, vals, probs, np.array([0, 0, 1]), n_samples=2)
    assert points.shape == (2, 2, 2, 2) and class_probs.shape == (2, 2)
    assert_array_equal(points[0, 0], pts[0])
    assert_array_equal(points[0, 1], pts[1])
    assert_array_equal(points[1, 0], pts[2])
    assert np.isnan(points[1, 1]).all() and np.isnan(class_probs[1, 1]) and np.isclose(class_probs[0, 1], 0.7)
    e_pts, e_vals, e_probs = classify_peaks_from_maps(np.zeros((1, 4, 4, 2), np.float32), np.zeros((0, 2)), np.zeros((0,)),
                                                      np.zeros((0,), np.int32), np.zeros((0,), np.int32), n_channels=3)
    assert e_pts.shape == (1, 2, 3, 2) and np.isnan(e_pts).all() and np.isnan(e_probs).all()
