"""GPU parity of PAF scoring / matching / grouping: reference known-answer vectors + seeded
comparison with the CPU oracle (candidate and assignment indices bit-exact, scores within 1e-4)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import cases_paf
from oracle import paf_grouping as opg
from oracle import peak_finding as opf
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def pg():
    from sleap_b200.nn import paf_grouping
    return paf_grouping


@pytest.mark.parametrize("case", cases_paf.FUNCTION_LEVEL, ids=lambda f: f.__name__)
def test_reference_known_answers(pg, case):
    case(pg)


def test_lsap_matches_scipy(pg):
    rng = np.random.default_rng(0)
    mats = []
    for n in range(1, 9):
        for m in range(1, 9):
            mats.append(rng.normal(size=(n, m)).astype(np.float32))
            mats.append(rng.integers(-2, 3, size=(n, m)).astype(np.float32))      # ties
    a = rng.normal(size=(5, 5)).astype(np.float32); a[1, 2] = np.nan; a[3, :2] = np.nan
    mats.append(a)
    b = rng.normal(size=(3, 4)).astype(np.float32); b[0, :] = np.nan               # infeasible -> no matches
    mats.append(b)
    mats.append(np.zeros((6, 6), np.float32))                                       # constant -> identity
    sols = pg._lsap_scores(mats)
    from scipy.optimize import linear_sum_assignment
    for mat, (r, c, s) in zip(mats, sols):
        cost = np.where(np.isnan(mat), np.inf, -mat)
        try:
            wr, wc = linear_sum_assignment(cost)
        except ValueError:
            wr = wc = np.zeros((0,), np.int64)
        assert_array_equal(r, wr, err_msg=str(mat))
        assert_array_equal(c, wc, err_msg=str(mat))
        assert_array_equal(s, mat[wr, wc])


def _frame(seed, **kw):
    return synth.make_bottomup_frame(seed, **kw)


@pytest.mark.parametrize("cfg", [dict(height=256, width=256, n_instances=3, centroid_margin=40.0, spread=25.0),
                                 dict(height=1024, width=1024, n_instances=5)])
def test_score_match_group_stagewise(pg, cfg):
    B = 2
    frames = [_frame(100 + i, noise=0.01, **cfg) for i in range(B)]
    cms = np.stack([f[1] for f in frames]); pafs = np.stack([f[2] for f in frames])
    p, v, si, ci = opf.find_local_peaks(cms, 0.2, "integral", 5)
    p = (p * np.float32(4)).astype(np.float32)
    peaks = [p[si == b] for b in range(B)]; vals = [v[si == b] for b in range(B)]; chans = [ci[si == b] for b in range(B)]
    edges = synth.flies13_edge_inds()
    want = opg.score_paf_lines_batch(pafs, peaks, chans, edges, 10, 8, 0.25, 1.0, 13)
    got = pg.score_paf_lines_batch(pafs, peaks, chans, edges, 10, 8, 0.25, 1.0, 13)
    for b in range(B):
        assert_array_equal(got[0][b], want[0][b]); assert_array_equal(got[1][b], want[1][b])
        assert_allclose(got[2][b], want[2][b], atol=TOL, rtol=0, equal_nan=True)
    wm = opg.match_candidates_batch(*want, 12)
    gm = pg.match_candidates_batch(*want, 12)          # same scores in -> identical assignment
    for k in range(4):
        for b in range(B):
            assert_array_equal(gm[k][b], wm[k][b])
    order = opg.toposort_edges(edges)
    wi = opg.group_instances_batch(peaks, vals, chans, *wm, 13, order, edges, 0, 0.25)
    gi = pg.group_instances_batch(peaks, vals, chans, *wm, 13, order, edges, 0, 0.25)
    for b in range(B):
        assert_array_equal(gi[0][b], wi[0][b]); assert_array_equal(gi[1][b], wi[1][b])
        assert_allclose(gi[2][b], wi[2][b], atol=1e-6)


@pytest.mark.parametrize("refinement", ["integral", "local", None])
def test_bottomup_from_maps_matches_oracle(pg, refinement):
    """The whole post-processing chain on identical maps (C4 shapes): bit-exact indices / assignments."""
    from sleap_b200.nn.inference import bottomup_from_maps
    B = 3
    frames = [_frame(200 + i, noise=0.01) for i in range(B)]
    cms = np.stack([f[1] for f in frames]); pafs = np.stack([f[2] for f in frames])
    oscorer = opg.PAFScorer(synth.FLIES13_NODES, synth.FLIES13_EDGES, 8)
    scorer = pg.PAFScorer(synth.FLIES13_NODES, synth.FLIES13_EDGES, 8)
    p, v, si, ci = opf.find_local_peaks(cms, 0.2, refinement, 5)
    p = (p * np.float32(4)).astype(np.float32)
    peaks = [p[si == b] for b in range(B)]; vals = [v[si == b] for b in range(B)]; chans = [ci[si == b] for b in range(B)]
    winst, wps, wisc, wei, wepi, wls = oscorer.predict(pafs, peaks, vals, chans)
    got = bottomup_from_maps(cms, pafs, scorer, 4, 0.2, refinement, 5)
    for b in range(B):
        assert_array_equal(got["peak_channel_inds"][b], chans[b])
        assert_allclose(got["peaks"][b], peaks[b], atol=TOL * 4, rtol=0)
        assert_array_equal(got["peak_vals"][b], vals[b])
        assert_array_equal(got["edge_inds"][b], wei[b]); assert_array_equal(got["edge_peak_inds"][b], wepi[b])
        assert_allclose(got["line_scores"][b], wls[b], atol=TOL, rtol=0, equal_nan=True)
        assert got["instance_peaks"][b].shape == winst[b].shape
        assert_array_equal(np.isnan(got["instance_peaks"][b]), np.isnan(winst[b]))      # assignment pattern
        assert_allclose(got["instance_peaks"][b], winst[b], atol=TOL * 4, rtol=0, equal_nan=True)
        assert_array_equal(got["instance_peak_vals"][b], wps[b])
        assert_allclose(got["instance_scores"][b], wisc[b], atol=TOL, rtol=0)
        assert got["flags"][b] == 0


def test_bottomup_from_maps_empty_and_input_scale(pg):
    from sleap_b200.nn.inference import bottomup_from_maps
    scorer = pg.PAFScorer(synth.FLIES13_NODES, synth.FLIES13_EDGES, 8)
    cms = np.zeros((2, 64, 64, 13), np.float32); pafs = np.zeros((2, 32, 32, 24), np.float32)
    got = bottomup_from_maps(cms, pafs, scorer, 4)
    assert got["instance_peaks"][0].shape == (0, 13, 2) and list(got["n_valid"]) == [0, 0]
    pts, cm, pf_ = _frame(9, height=256, width=256, n_instances=2, centroid_margin=50.0, spread=20.0)
    a = bottomup_from_maps(cm[None], pf_[None], scorer, 4, input_scale=1.0)
    b = bottomup_from_maps(cm[None], pf_[None], scorer, 4, input_scale=0.5)
    assert_allclose(b["instance_peaks"][0], a["instance_peaks"][0] / np.float32(0.5) + np.float32(0.5), atol=1e-4, equal_nan=True)


def test_greedy_merge_semantics(pg):
    """Unsorted edge order exercises the 'both assigned' merge/steal branches (paf_grouping.py:848-883)."""
    rng = np.random.default_rng(4)
    n_nodes, edges = 5, [(0, 1), (2, 3), (1, 2), (3, 4), (0, 4)]
    for trial in range(20):
        counts = rng.integers(1, 4, size=n_nodes)
        ch = np.concatenate([np.full(c, i) for i, c in enumerate(counts)]).astype(np.int32)
        perm = rng.permutation(len(ch)); ch = ch[perm]
        peaks = rng.uniform(0, 100, size=(len(ch), 2)).astype(np.float32)
        vals = rng.uniform(0.3, 1, size=len(ch)).astype(np.float32)
        me, ms, md, msc = [], [], [], []
        for k, (a, b) in enumerate(edges):
            n = min(counts[a], counts[b])
            s = rng.permutation(counts[a])[:n]; d = rng.permutation(counts[b])[:n]
            o = np.argsort(s)
            me += [k] * n; ms += list(s[o]); md += list(d[o]); msc += list(rng.uniform(0.1, 1.0, size=n))
        args = (peaks, vals, ch, np.array(me, np.int32), np.array(ms, np.int32), np.array(md, np.int32),
                np.array(msc, np.float32), n_nodes, tuple(range(len(edges))), edges)
        for mip in (0, 2):
            w = opg.group_instances_sample(*args, mip, 0.25)
            g = pg.group_instances_sample(*args, mip, 0.25)
            ctx = f"trial={trial} mip={mip} counts={counts} ch={ch} me={me} ms={ms} md={md} msc={np.round(msc, 3)}\nGOT={g[0]}\nWANT={w[0]}"
            assert g[0].shape == w[0].shape, ctx
            assert np.array_equal(np.nan_to_num(g[0], nan=-1), np.nan_to_num(w[0], nan=-1)), ctx
            assert_array_equal(g[1], w[1]); assert_allclose(g[2], w[2], atol=1e-6)
