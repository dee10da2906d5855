"""Known-answer cases ported from the reference's tests/nn/test_peak_finding.py and
tests/nn/data/test_instance_cropping.py.  Each ``check_*`` takes a
module-like object ``pf`` exposing the reference's function-level API, so the very same
assertions run against the CPU oracle (``-m "not gpu"``) and the CUDA path (``-m gpu``)."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle.synth import make_confmaps, make_grid_vectors, make_multi_confmaps

F32 = np.float32


def check_find_local_offsets(pf):
    # reference tests/nn/test_peak_finding.py:28-46
    off = pf.find_offsets_local_direction(
        np.array([[0.0, 1.0, 0.0], [1.0, 3.0, 2.0], [0.0, 1.0, 0.0]], F32).reshape(1, 3, 3, 1), 0.25)
    assert tuple(off.shape) == (1, 2)
    assert off[0][0] == 0.25 and off[0][1] == 0.0
    off = pf.find_offsets_local_direction(
        np.array([[0.0, 1.0, 0.0], [1.0, 3.0, 1.0], [0.0, 1.0, 0.0]], F32).reshape(1, 3, 3, 1), 0.25)
    assert off[0][0] == 0.0 and off[0][1] == 0.0


def check_global_peaks_rough(pf):
    # :49-73
    xv, yv = make_grid_vectors(8, 8, 1)
    points = np.array([[1, 2], [3, 4], [5, 6]], F32)
    cm = make_confmaps(points, xv, yv, sigma=1.0)
    points2 = points + 1
    cms = np.stack([cm, make_confmaps(points2, xv, yv, sigma=1.0)])
    peaks, vals = pf.find_global_peaks(cms, threshold=0.1, refinement=None)
    assert peaks.shape == (2, 3, 2) and vals.shape == (2, 3)
    assert_array_equal(peaks[0], points)
    assert_array_equal(vals[0], [1, 1, 1])
    assert_array_equal(peaks[1], points2)
    assert_array_equal(vals[1], [1, 1, 1])
    peaks, vals = pf.find_global_peaks_rough(np.zeros((1, 8, 8, 3), F32), threshold=0.1)
    assert peaks.shape == (1, 3, 2) and vals.shape == (1, 3)
    assert np.all(np.isnan(peaks))
    assert_array_equal(vals, [[0, 0, 0]])


def check_global_peaks_integral(pf):
    # :76-121
    xv, yv = make_grid_vectors(12, 12, 1)
    points = np.array([[1.5, 2.5], [3.5, 4.5], [5.5, 6.5]], F32)
    cm = make_confmaps(points, xv, yv, sigma=1.0)
    peaks, vals = pf.find_global_peaks(cm[None], threshold=0.1, refinement="integral", integral_patch_size=5)
    assert peaks.shape == (1, 3, 2) and vals.shape == (1, 3)
    assert_allclose(peaks[0], points, atol=0.1)
    assert_allclose(vals[0], [1, 1, 1], atol=0.3)
    peaks, vals = pf.find_global_peaks(np.zeros((1, 8, 8, 3), F32), threshold=0.1,
                                       refinement="integral", integral_patch_size=5)
    assert np.all(np.isnan(peaks))
    assert_array_equal(vals, [[0, 0, 0]])
    peaks, vals = pf.find_global_peaks(np.stack([np.zeros((12, 12, 3), F32), cm]), threshold=0.1,
                                       refinement="integral", integral_patch_size=5)
    assert peaks.shape == (2, 3, 2)
    assert np.all(np.isnan(peaks[0]))
    assert_allclose(peaks[1], points, atol=0.1)


def check_global_peaks_local(pf):
    # :124-138
    xv, yv = make_grid_vectors(12, 12, 1)
    points = np.array([[1.6, 2.6], [3.6, 4.6], [5.6, 6.6]], F32)
    cm = make_confmaps(points, xv, yv, sigma=1.0)
    peaks, vals = pf.find_global_peaks(cm[None], threshold=0.1, refinement="local")
    assert_allclose(peaks[0], np.array([[1.75, 2.75], [3.75, 4.75], [5.75, 6.75]]))
    assert_allclose(vals[0], [1, 1, 1], atol=0.3)


def _local_case(scale=1.0, shift=0.0, size=16):
    xv, yv = make_grid_vectors(size, size, 1)
    inst = np.array([[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[np.nan, np.nan], [11, 12]]], F32) * scale + shift
    cms = make_multi_confmaps(inst, xv, yv, sigma=1.0)
    inst2 = np.array([[[2, 3], [4, 5]], [[6, 7], [8, 9]]], F32) * scale + shift
    cms = np.stack([cms, make_multi_confmaps(inst2, xv, yv, sigma=1.0)], axis=0)
    expected = np.array([[1, 2], [3, 4], [5, 6], [7, 8], [11, 12], [2, 3], [4, 5], [6, 7], [8, 9]], F32) * scale + shift
    return cms, expected


def check_local_peaks_rough(pf):
    # :141-198
    cms, expected = _local_case()
    pts, vals, si, ci = pf.find_local_peaks(cms, threshold=0.1, refinement=None)
    assert pts.shape == (9, 2) and vals.shape == (9,) and si.shape == (9,) and ci.shape == (9,)
    assert_array_equal(pts, expected)
    assert_array_equal(vals, np.ones(9))
    assert_array_equal(si, [0, 0, 0, 0, 0, 1, 1, 1, 1])
    assert_array_equal(ci, [0, 1, 0, 1, 1, 0, 1, 0, 1])
    pts, vals, si, ci = pf.find_local_peaks(np.zeros((1, 4, 4, 3), F32), threshold=0.1, refinement=None)
    assert pts.shape == (0, 2) and vals.shape == (0,) and si.shape == (0,) and ci.shape == (0,)


def check_local_peaks_integral(pf):
    # :201-256
    cms, expected = _local_case(2.0, 0.3, 32)
    pts, vals, si, ci = pf.find_local_peaks(cms, threshold=0.1, refinement="integral", integral_patch_size=5)
    assert pts.shape == (9, 2)
    assert_allclose(pts, expected, atol=0.2)
    assert_allclose(vals, np.ones(9), atol=0.1)
    assert_array_equal(si, [0, 0, 0, 0, 0, 1, 1, 1, 1])
    assert_array_equal(ci, [0, 1, 0, 1, 1, 0, 1, 0, 1])
    pts, vals, si, ci = pf.find_local_peaks(np.zeros((1, 4, 4, 3), F32), refinement="integral", integral_patch_size=5)
    assert pts.shape == (0, 2) and vals.shape == (0,)


def check_local_peaks_local(pf):
    # :283-337
    cms, expected = _local_case(2.0, 0.25, 32)
    pts, vals, si, ci = pf.find_local_peaks(cms, threshold=0.1, refinement="local")
    assert_allclose(pts, expected)
    assert_allclose(vals, np.ones(9), atol=0.1)
    assert_array_equal(si, [0, 0, 0, 0, 0, 1, 1, 1, 1])
    assert_array_equal(ci, [0, 1, 0, 1, 1, 0, 1, 0, 1])


def check_peaks_with_offsets(pf):
    # semantics of :340-391 without the labels fixture: offsets are gathered at rough peaks.
    cms, expected = _local_case()
    rng = np.random.default_rng(0)
    B, H, W, C = cms.shape
    offs = rng.uniform(-0.5, 0.5, size=(B, H, W, 2 * C)).astype(F32)
    pts, vals, si, ci = pf.find_local_peaks_with_offsets(cms, offs, threshold=0.1)
    o5 = offs.reshape(B, H, W, C, 2)
    want = expected + o5[si, expected[:, 1].astype(int), expected[:, 0].astype(int), ci]
    assert_allclose(pts, want, atol=1e-6)
    gp, gv = pf.find_global_peaks_with_offsets(cms, offs, threshold=0.1)
    rp, _ = pf.find_global_peaks_rough(cms, threshold=0.1)
    for s in range(B):
        for c in range(C):
            x, y = int(rp[s, c, 0]), int(rp[s, c, 1])
            assert_allclose(gp[s, c], rp[s, c] + o5[s, y, x, c], atol=1e-6)


ALL = [check_find_local_offsets, check_global_peaks_rough, check_global_peaks_integral,
       check_global_peaks_local, check_local_peaks_rough, check_local_peaks_integral,
       check_local_peaks_local, check_peaks_with_offsets]
