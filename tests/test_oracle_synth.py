"""The synthetic-input generators the parity tests feed on (oracle/synth.py) against the reference's known answers:
tests/nn/data/test_confidence_maps.py:21-108 (make_confmaps, make_multi_confmaps)."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle import synth


def test_make_confmaps():
    xv, yv = synth.make_grid_vectors(4, 5, 1)
    points = np.asarray([[0.5, 1.0], [3, 3.5], [2.0, 2.0]], np.float32)
    cm = synth.make_confmaps(points, xv, yv, 1.0)
    assert cm.dtype == np.float32 and cm.shape == (4, 5, 3)
    want = [[[0.535, 0.0, 0.018], [0.535, 0.0, 0.082], [0.197, 0.001, 0.135], [0.027, 0.002, 0.082], [0.001, 0.001, 0.018]],
            [[0.882, 0.0, 0.082], [0.882, 0.006, 0.368], [0.325, 0.027, 0.607], [0.044, 0.044, 0.368], [0.002, 0.027, 0.082]],
            [[0.535, 0.004, 0.135], [0.535, 0.044, 0.607], [0.197, 0.197, 1.0], [0.027, 0.325, 0.607], [0.001, 0.197, 0.135]],
            [[0.119, 0.01, 0.082], [0.119, 0.119, 0.368], [0.044, 0.535, 0.607], [0.006, 0.882, 0.368], [0.0, 0.535, 0.082]]]
    assert_allclose(cm, want, atol=1e-3)
    assert synth.make_confmaps(np.asarray([[2, 3]], np.float32), xv, yv, 1.0)[3, 2] == 1.0        # grid-aligned peak
    xv, yv = synth.make_grid_vectors(8, 8, 2)                                                       # output stride
    cm = synth.make_confmaps(np.asarray([[2, 4]], np.float32), xv, yv, 1.0)
    assert cm.shape == (4, 4, 1) and cm[2, 1] == 1.0
    cmn = synth.make_confmaps(np.asarray([[2, 4], [np.nan, np.nan]], np.float32), xv, yv, 1.0)    # missing points -> zeros
    assert cmn.shape == (4, 4, 2) and cmn.dtype == np.float32
    assert_array_equal(cmn[:, :, 0], cm[:, :, 0])
    assert (cmn[:, :, 1] == 0).all()


def test_make_multi_confmaps():
    xv, yv = synth.make_grid_vectors(4, 5, 1)
    inst = np.asarray([[[0.5, 1.0], [2.0, 2.0]], [[1.5, 1.0], [2.0, 3.0]], [[np.nan, np.nan], [-1.0, 5.0]]], np.float32)
    cms = synth.make_multi_confmaps(inst, xv, yv, 1.0)
    assert cms.shape == (4, 5, 2) and cms.dtype == np.float32
    each = np.stack([synth.make_confmaps(i, xv, yv, 1.0) for i in inst], axis=-1)
    assert_array_equal(cms, each.max(axis=-1))
