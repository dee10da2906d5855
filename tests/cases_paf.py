"""Known-answer cases ported from the reference's tests/nn/test_paf_grouping.py and
tests/nn/test_nn_utils.py.  ``pg`` is a module-like object exposing the
reference's function-level API; ragged values are lists of per-sample arrays."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

F32 = np.float32


def check_connection_candidates(pg):
    # reference tests/nn/test_paf_grouping.py:28-41
    ei, epi = pg.get_connection_candidates([0, 0, 0, 1, 1, 2], [[0, 1], [1, 2], [2, 3]], 4)
    assert_array_equal(ei, [0, 0, 0, 0, 0, 0, 1, 1])
    assert_array_equal(epi, [[0, 3], [0, 4], [1, 3], [1, 4], [2, 3], [2, 4], [3, 5], [4, 5]])


def check_make_line_subs(pg):
    # :44-55
    subs = pg.make_line_subs(np.array([[0, 0], [4, 8]], F32), np.array([[0, 1]], np.int32),
                             np.array([0], np.int32), n_line_points=3, pafs_stride=2)
    assert_array_equal(subs, [[[[0, 0, 0], [0, 0, 1]], [[2, 1, 0], [2, 1, 1]], [[4, 2, 0], [4, 2, 1]]]])


def _paf_case():
    pafs = np.arange(6 * 4 * 2, dtype=F32).reshape(6, 4, 2)
    peaks = np.array([[0, 0], [4, 8]], F32)
    return pafs, peaks, np.array([[0, 1]], np.int32), np.array([0], np.int32)


def check_paf_lines(pg):
    # :58-72
    pafs, peaks, epi, ei = _paf_case()
    lines = pg.get_paf_lines(pafs, peaks, epi, ei, n_line_points=3, pafs_stride=2)
    assert_array_equal(lines, [[[0, 1], [18, 19], [36, 37]]])


def check_score_paf_lines(pg):
    # :75-90
    pafs, peaks, epi, ei = _paf_case()
    lines = pg.get_paf_lines(pafs, peaks, epi, ei, n_line_points=3, pafs_stride=2)
    scores = pg.score_paf_lines(lines, peaks, epi, max_edge_length=2)
    assert_allclose(scores, [24.27], atol=1e-2)


def check_distance_penalty(pg):
    # :93-102
    pen = pg.compute_distance_penalty(np.array([1, 2, 3, 4], F32), max_edge_length=2)
    assert_allclose(pen, [0, 0, 2 / 3 - 1, 2 / 4 - 1], atol=1e-6)
    pen = pg.compute_distance_penalty(np.array([1, 2, 3, 4], F32), max_edge_length=2, dist_penalty_weight=2)
    assert_allclose(pen, [0, 0, -0.6666666, -1], atol=1e-6)


def check_score_paf_lines_batch(pg):
    # :105-129
    pafs = np.arange(6 * 4 * 2, dtype=F32).reshape(1, 6, 4, 2)
    ei, epi, ls = pg.score_paf_lines_batch(pafs, [np.array([[0, 0], [4, 8]], F32)], [np.array([0, 1], np.int32)],
                                           np.array([[0, 1], [1, 2], [2, 3]], np.int32), 3, 2, 2 / 12, 1.0, 4)
    assert_array_equal(ei[0], [0])
    assert_array_equal(epi[0], [[0, 1]])
    assert_allclose(ls[0], [24.27], atol=1e-2)


def check_match_candidates_sample(pg):
    # :132-158
    me, ms, md, msc = pg.match_candidates_sample(np.array([0, 0], np.int32), np.array([[0, 1], [2, 1]], np.int32),
                                                 np.array([-0.5, 1.0], F32), 1)
    assert_array_equal(me, [0])
    assert_array_equal(ms, [1])
    assert_array_equal(md, [0])
    assert_array_equal(msc, [1.0])


def check_match_candidates_batch(pg):
    # :161-185
    me, ms, md, msc = pg.match_candidates_batch([np.array([0, 0], np.int32)], [np.array([[0, 1], [2, 1]], np.int32)],
                                                [np.array([-0.5, 1.0], F32)], 1)
    assert_array_equal(np.concatenate(me), [0])
    assert_array_equal(np.concatenate(ms), [1])
    assert_array_equal(np.concatenate(md), [0])
    assert_array_equal(np.concatenate(msc), [1.0])


def _group_case():
    return dict(
        peaks=np.arange(10, dtype=F32).reshape(5, 2), scores=np.arange(5, dtype=F32),
        ch=np.array([0, 1, 2, 0, 1], np.int32), me=np.array([0, 1, 0], np.int32),
        ms=np.array([0, 0, 1], np.int32), md=np.array([0, 0, 1], np.int32), msc=np.ones(3, F32))


def check_group_instances_sample(pg):
    # :188-230
    c = _group_case()
    inst, ps, isc = pg.group_instances_sample(c["peaks"], c["scores"], c["ch"], c["me"], c["ms"], c["md"], c["msc"],
                                              3, (0, 1), [(0, 1), (1, 2)], 0)
    assert_array_equal(inst, [[[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]], [[6.0, 7.0], [8.0, 9.0], [np.nan, np.nan]]])
    assert_array_equal(ps, [[0.0, 1.0, 2.0], [3.0, 4.0, np.nan]])
    assert_array_equal(isc, [2.0, 1.0])


def check_group_instances_batch(pg):
    # :233-299
    c = _group_case()
    inst, ps, isc = pg.group_instances_batch([c["peaks"]], [c["scores"]], [c["ch"]], [c["me"]], [c["ms"]], [c["md"]],
                                             [c["msc"]], 3, (0, 1), [(0, 1), (1, 2)], 0)
    assert_array_equal(inst[0], [[[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]], [[6.0, 7.0], [8.0, 9.0], [np.nan, np.nan]]])
    assert_array_equal(ps[0], [[0.0, 1.0, 2.0], [3.0, 4.0, np.nan]])
    assert_array_equal(isc[0], [2.0, 1.0])


TOPO_A = [(5, 7), (5, 8), (5, 9), (5, 6), (5, 11), (5, 12), (1, 0), (1, 3), (1, 2), (1, 10), (1, 13), (1, 14),
          (4, 5), (4, 1)]
TOPO_B = [(1, 4), (1, 5), (6, 8), (6, 7), (6, 9), (9, 10), (1, 0), (1, 3), (1, 2), (6, 1)]


def check_toposort(pg):
    # :302-339
    assert tuple(pg.toposort_edges(TOPO_A)) == (12, 13, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11)
    assert tuple(pg.toposort_edges(TOPO_B)) == (2, 3, 4, 9, 5, 0, 1, 6, 7, 8)


def greedy_connections():
    # :342-375
    conns = {
        (5, 7): [(0, 0, 1.0465653)], (5, 8): [(0, 0, 1.0607507)], (5, 9): [(0, 0, 0.9563284)],
        (5, 6): [(0, 1, 0.5797864)], (5, 11): [(0, 0, 0.9892818)], (5, 12): [(0, 0, 0.7557168)],
        (1, 0): [], (1, 3): [], (1, 2): [], (1, 10): [], (1, 13): [], (1, 14): [],
        (4, 5): [(0, 0, 0.9735552)], (4, 1): [(0, 0, 0.31536198)],
    }
    return conns


def check_assign_connections(pg):
    # :342-403
    conns = greedy_connections()
    a = pg.assign_connections_to_instances(conns, min_instance_peaks=0, n_nodes=15)
    assert a == {(5, 0): 0, (7, 0): 0, (8, 0): 0, (9, 0): 0, (6, 1): 0, (11, 0): 0, (12, 0): 0, (4, 0): 1, (1, 0): 1}
    types = list(conns.keys())
    order = pg.toposort_edges(types)
    a = pg.assign_connections_to_instances({types[i]: conns[types[i]] for i in order}, min_instance_peaks=0, n_nodes=15)
    assert all(v == 0 for v in a.values())


def check_lsap(pg):
    # reference tests/nn/test_nn_utils.py:12-24
    r, c = pg.linear_sum_assignment(np.array([[-1, 0], [0, -1]], F32))
    assert_array_equal(r, [0, 1])
    assert_array_equal(c, [0, 1])


FUNCTION_LEVEL = [check_connection_candidates, check_make_line_subs, check_paf_lines, check_score_paf_lines,
                  check_distance_penalty, check_score_paf_lines_batch, check_match_candidates_sample,
                  check_match_candidates_batch, check_group_instances_sample, check_group_instances_batch,
                  check_toposort, check_lsap]
HOST_ONLY = [check_assign_connections]
