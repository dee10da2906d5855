"""Pins the CPU oracle's PAF grouping against the reference's known-answer vectors (no GPU)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import cases_paf
from oracle import paf_grouping as opg
from oracle import synth


@pytest.mark.parametrize("case", cases_paf.FUNCTION_LEVEL + cases_paf.HOST_ONLY, ids=lambda f: f.__name__)
def test_oracle_paf_case(case):
    case(opg)


def test_edge_maps_known_answers():
    # reference tests/nn/data/test_edge_maps.py:12-58 (confirms the d^4 quirk)
    xv = np.array([0, 1, 2], np.float32)
    yv = np.array([0, 1, 2], np.float32)
    gx, gy = np.meshgrid(xv, yv)
    grid = np.stack([gx, gy], -1)
    es = np.array([[1, 0.5], [0, 0]], np.float32)
    ed = np.array([[1, 1.5], [2, 2]], np.float32)
    d = synth.distance_to_edge(grid, es, ed)
    assert_allclose(d[0], [[1.25, 0.0], [0.25, 0.5], [1.25, 2.0]], atol=1e-3)
    em = synth.make_edge_maps(xv, yv, es, ed, sigma=1.0)
    assert_allclose(em[0], [[0.458, 1.0], [0.969, 0.882], [0.458, 0.135]], atol=1e-3)


def test_flies13_toposort_is_bfs_from_thorax():
    order = opg.toposort_edges(synth.flies13_edge_inds())
    assert order == tuple(range(12))


def test_bottomup_synthetic_recovers_instances():
    pts, cms, pafs = synth.make_bottomup_frame(seed=3)
    from oracle import peak_finding as opf
    p, v, si, ci = opf.find_local_peaks(cms[None], 0.2, "integral", 5)
    p = p * np.float32(4)
    scorer = opg.PAFScorer(synth.FLIES13_NODES, synth.FLIES13_EDGES, pafs_stride=8)
    inst, ps, isc, *_ = scorer.predict(pafs[None], [p], [v], [ci])
    assert inst[0].shape[0] >= 4
    # every predicted node within 2 px of some true node of the same type
    for row in inst[0]:
        for n in range(13):
            if not np.isnan(row[n, 0]):
                assert np.min(np.linalg.norm(pts[:, n] - row[n], axis=1)) < 2.0
