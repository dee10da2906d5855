"""GPU parity of the peak-finding kernels: reference known-answer vectors + seeded comparison with
the CPU oracle (indices bit-exact, sub-pixel offsets / values within 1e-4)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import cases_peaks
from oracle import peak_finding as opf
from oracle import synth
from oracle import tf_ops

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def pf():
    from sleap_b200.nn import peak_finding
    return peak_finding


@pytest.mark.parametrize("case", cases_peaks.ALL, ids=lambda f: f.__name__)
def test_reference_known_answers(pf, case):
    case(pf)


def _random_cms(seed, B, H, W, C, n_inst=4, sigma=2.0, noise=0.02):
    rng = np.random.default_rng(seed)
    xv, yv = synth.make_grid_vectors(H, W, 1)
    cms = []
    for b in range(B):
        pts = rng.uniform(-1, [W + 1, H + 1], size=(n_inst, C, 2)).astype(np.float32)   # some near / past the border
        pts[0, 0] = [0.3, 0.2]                 # border peaks (crop_and_resize extrapolation cases)
        pts[n_inst - 1, min(1, C - 1)] = [W - 1.2, H - 1.1]
        cm = np.zeros((H, W, C), np.float32)
        for p in pts:
            cm = np.maximum(cm, synth.make_confmaps(p, xv, yv, sigma))
        cms.append(cm + rng.normal(0, noise, cm.shape).astype(np.float32))
    return np.stack(cms).astype(np.float32)


@pytest.mark.parametrize("shape", [(2, 64, 64, 5), (3, 40, 56, 13), (1, 256, 256, 13), (2, 33, 47, 3)])
@pytest.mark.parametrize("refinement", [None, "integral", "local"])
def test_local_peaks_match_oracle(pf, shape, refinement):
    cms = _random_cms(11 + shape[1], *shape)
    want = opf.find_local_peaks(cms, threshold=0.2, refinement=refinement, integral_patch_size=5)
    got = pf.find_local_peaks(cms, threshold=0.2, refinement=refinement, integral_patch_size=5)
    assert len(got[0]) == len(want[0]) and len(want[0]) > 0
    assert_array_equal(got[2], want[2])          # sample inds
    assert_array_equal(got[3], want[3])          # channel inds
    assert_array_equal(got[1], want[1])          # values are copies of the map
    assert_array_equal(np.floor(got[0] + 0.5), np.floor(want[0] + 0.5)) if refinement is None else None
    assert_allclose(got[0], want[0], atol=TOL, rtol=0, equal_nan=True)


@pytest.mark.parametrize("patch", [3, 5, 7])
def test_integral_patch_sizes(pf, patch):
    cms = _random_cms(5, 2, 48, 48, 4)
    want = opf.find_local_peaks(cms, 0.2, "integral", patch)
    got = pf.find_local_peaks(cms, 0.2, "integral", patch)
    assert_array_equal(got[3], want[3])
    assert_allclose(got[0], want[0], atol=TOL, rtol=0, equal_nan=True)


@pytest.mark.parametrize("shape", [(2, 64, 64, 5), (4, 40, 40, 13), (1, 256, 256, 13), (3, 31, 45, 2)])
@pytest.mark.parametrize("refinement", [None, "integral", "local"])
def test_global_peaks_match_oracle(pf, shape, refinement):
    cms = _random_cms(7 + shape[2], *shape, n_inst=1)
    cms[0, :, :, 0] = 0.0                                    # below threshold -> NaN point
    want = opf.find_global_peaks(cms, threshold=0.2, refinement=refinement, integral_patch_size=5)
    got = pf.find_global_peaks(cms, threshold=0.2, refinement=refinement, integral_patch_size=5)
    assert_array_equal(np.isnan(got[0]), np.isnan(want[0]))
    assert_array_equal(got[1], want[1])
    assert_allclose(got[0], want[0], atol=TOL, rtol=0, equal_nan=True)


def test_global_ties_first_index(pf):
    cms = np.zeros((1, 6, 7, 2), np.float32)
    cms[0, 4, 1, 0] = 0.9
    cms[0, 2, 5, 0] = 0.9       # tie: row = first row with the max (2), col = first col with the max (1)
    cms[0, 3, 3, 1] = 0.5
    want = opf.find_global_peaks_rough(cms, 0.1)
    got = pf.find_global_peaks_rough(cms, 0.1)
    assert_array_equal(np.isnan(got[0]), np.isnan(want[0]))
    assert_array_equal(np.nan_to_num(got[0]), np.nan_to_num(want[0]))
    assert_array_equal(got[1], want[1])


def test_empty_and_plateau_and_border(pf):
    cms = np.zeros((2, 9, 9, 3), np.float32)
    cms[1, 0, 0, 0] = 0.9
    cms[1, 8, 8, 2] = 0.8
    cms[1, 4, 4, 1] = 0.5
    cms[1, 4, 5, 1] = 0.5       # plateau -> no peak
    for ref in (None, "integral", "local"):
        want = opf.find_local_peaks(cms, 0.2, ref, 5)
        got = pf.find_local_peaks(cms, 0.2, ref, 5)
        assert len(want[0]) == 2
        assert_array_equal(got[2], want[2]); assert_array_equal(got[3], want[3])
        assert_allclose(got[0], want[0], atol=TOL, rtol=0, equal_nan=True)


def test_capacity_truncation_is_ordered(pf):
    cms = _random_cms(3, 1, 64, 64, 5, noise=0.0)
    full = pf._local(cms, 0.2, None, 5, None, None)
    cut = pf._local(cms, 0.2, None, 5, None, None, max_peaks_per_sample=3)
    assert len(cut[0]) == 3
    assert_array_equal(cut[0], full[0][:3])


def test_crops_match_oracle(pf):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(2, 70, 90, 1), dtype=np.uint8)
    cent = np.array([[30.3, 20.7], [2.1, 3.4], [88.2, 68.9], [45.0, 35.0]], np.float32)
    sinds = np.array([0, 1, 1, 0], np.int32)
    for size in (8, 9):
        bb = tf_ops.make_centered_bboxes(cent, size, size)
        want = tf_ops.crop_bboxes(img, bb, sinds)
        got = pf.crop_bboxes(img, bb, sinds)
        assert_array_equal(got, want)
        wantf = tf_ops.crop_bboxes(img.astype(np.float32), bb, sinds)
        gotf = pf.crop_bboxes(img.astype(np.float32), bb, sinds)
        assert_allclose(gotf, wantf, atol=1e-3)


def test_integral_regression_and_local_dir(pf):
    rng = np.random.default_rng(1)
    cms = rng.uniform(0, 1, size=(6, 5, 5, 3)).astype(np.float32)
    gv = np.arange(5, dtype=np.float32) - 2
    wx, wy = opf.integral_regression(cms, gv, gv)
    gx, gy = pf.integral_regression(cms, gv, gv)
    assert_allclose(gx, wx, atol=1e-5); assert_allclose(gy, wy, atol=1e-5)
    patches = rng.normal(size=(10, 3, 3, 1)).astype(np.float32)
    assert_array_equal(pf.find_offsets_local_direction(patches, 0.25), opf.find_offsets_local_direction(patches, 0.25))
