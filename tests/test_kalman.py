"""sleap_b200.nn.kalman: the matching logic against the reference's known-answer tests (tests/nn/test_kalman.py), the
argument checks of ``Tracker.make_tracker_by_name`` (tests/nn/test_tracking_integration.py:27-95), the in-tree Kalman
filter / EM on analytic cases, and the whole tracker on two synthetic animals whose paths cross."""
import numpy as np
import pytest

from sleap_b200.nn import kalman as k
from sleap_b200.nn.tracking import Tracker, first_choice_matching, greedy_matching


def test_first_choice_matching():
    """tests/nn/test_kalman.py:8-66."""
    instances, tracks = ["instance a", "instance b"], ["track a", "track b"]
    cost = np.array([[10, 150], [50, 100]])
    mt = k.match_tuples_from_match_function(cost, instances, tracks, first_choice_matching)
    assert len(mt) == 2 and ("instance a", "track a", 10) in mt and ("instance b", "track a", 50) in mt
    by_track = k.match_dict_from_match_function(cost, instances, tracks, first_choice_matching)
    assert by_track == {"track a": "instance a"}
    by_inst = k.match_dict_from_match_function(cost, instances, tracks, first_choice_matching, key_by_column=False)
    assert by_inst == {"instance a": "track a", "instance b": "track a"}
    cost = np.array([[50, 100], [10, 150]])        # the best match per track regardless of the row order
    assert k.match_dict_from_match_function(cost, instances, tracks, first_choice_matching) == {"track a": "instance b"}


def test_greedy_matching():
    """tests/nn/test_kalman.py:69-94."""
    instances, tracks = ["instance a", "instance b"], ["track a", "track b"]
    cost = np.array([[10, 200], [75, 150]])
    m = k.matches_from_match_tuples(k.match_tuples_from_match_function(cost, instances, tracks, greedy_matching))
    assert [(x.track, x.instance, x.score) for x in m] == [("track a", "instance a", 10), ("track b", "instance b", 150)]


def test_track_instance_matches():
    """tests/nn/test_kalman.py:97-172."""
    instances, tracks = ["instance a", "instance b"], ["track a", "track b"]
    for cost, want in ((np.array([[10, 200], [75, 150]]), [("track a", "instance a", 10), ("track b", "instance b", 150)]),
                       (np.array([[10, 100], [50, 150]]), [("track a", "instance a", 10), ("track b", "instance b", 150)]),
                       (np.array([[50, 100], [10, 150]]), [("track a", "instance b", 10), ("track b", "instance a", 100)])):
        m = k.get_track_instance_matches(cost, instances=instances, tracks=tracks, are_too_close_function=lambda x, y: True)
        assert [(x.track, x.instance, x.score) for x in m] == want


def test_remove_second_bests():
    """kalman.py:578-667: a column whose two best entries are within the threshold is cleared, and so is the row whose best
    entry was in it."""
    cost = np.array([[1.0, 50.0], [2.0, 60.0], [40.0, 5.0]])
    out = k.remove_second_bests_from_cost_matrix(cost, thresh=3.0)
    assert np.isnan(out[:, 0]).all()                   # 1.0 vs 2.0: ambiguous track
    assert np.isnan(out[0]).all() and np.isnan(out[1]).all()      # their best (cleared) choice cannot be replaced
    assert out[2, 1] == 5.0
    assert np.array_equal(k.remove_second_bests_from_cost_matrix(np.array([[1.0, 50.0], [60.0, 2.0]]), 3.0), [[1.0, 50.0], [60.0, 2.0]])


def test_make_tracker_by_name_kalman_argument_checks():
    """tests/nn/test_tracking_integration.py:27-95 (same messages)."""
    kw = dict(match="greedy", track_window=5, kf_init_frame_count=10)
    with pytest.raises(ValueError, match="Kalman filter requires simple tracker for initial tracking."):
        Tracker.make_tracker_by_name(tracker="flow", max_tracking=True, max_tracks=2, similarity="instance", kf_node_indices=[0, 1], **kw)
    with pytest.raises(ValueError, match="Kalman filter does not support normalized_instance_similarity."):
        Tracker.make_tracker_by_name(tracker="simple", max_tracking=True, max_tracks=2, similarity="normalized_instance", kf_node_indices=[0, 1], **kw)
    with pytest.raises(ValueError, match="Kalman filter requires node indices for instance tracking."):
        Tracker.make_tracker_by_name(tracker="simple", max_tracking=True, max_tracks=2, similarity="instance", **kw)
    with pytest.raises(ValueError, match="Kalman filter requires max tracks or target instance count."):
        Tracker.make_tracker_by_name(tracker="simple", similarity="instance", kf_node_indices=[0, 1], **kw)
    t = Tracker.make_tracker_by_name(tracker="simple", similarity="instance", target_instance_count=2, kf_node_indices=[0, 1], **kw)
    assert isinstance(t, k.KalmanTracker) and t.get_name().startswith("kalman.") and t.kalman_tracker.instance_count == 2
    t = Tracker.make_tracker_by_name(tracker="simple", similarity="iou", max_tracking=True, max_tracks=2, kf_node_indices=[0, 1], **kw)
    assert isinstance(t, k.KalmanTracker) and t.kalman_tracker.instance_count == 2


def test_kalman_filter_constant_velocity_and_em():
    """Noise-free constant-velocity track: the filter's one-step prediction lands on the next point and skips a masked
    observation; EM on noisy data shrinks the observation covariance towards the true noise level and keeps Q, R symmetric
    positive semi-definite."""
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    C = np.array([[1.0, 0.0]])
    x = 3.0 + 2.0 * np.arange(30)
    kf = k.KalmanFilter(A, C, initial_state_mean=[x[0], 0.0])
    m, c = kf.filter(x[:, None])
    nm, _ = kf.filter_update(m[-1], c[-1], np.ma.masked)
    assert abs(nm[0] - (x[-1] + 2.0)) < 0.5 and abs(nm[1] - 2.0) < 0.2
    nm2, _ = kf.filter_update(m[-1], c[-1], np.array([x[-1] + 2.0]))
    assert abs(nm2[0] - (x[-1] + 2.0)) < 0.2
    rng = np.random.default_rng(0)
    z = np.ma.masked_invalid((x + rng.normal(0, 0.5, x.shape))[:, None])
    z[7] = np.ma.masked
    kf2 = k.KalmanFilter(A, C, initial_state_mean=[float(z[0, 0]), 0.0]).em(z, n_iter=20)
    assert 0.02 < kf2.R[0, 0] < 1.0                      # true variance 0.25; the default start is 1.0
    for M in (kf2.Q, kf2.R, kf2.P0):
        assert np.allclose(M, M.T, atol=1e-9) and np.all(np.linalg.eigvalsh((M + M.T) / 2) > -1e-9)
    m2, _ = kf2.filter(z)
    assert np.abs(m2[10:, 0] - x[10:]).mean() < 0.5


class _Inst:
    def __init__(self, pts, score=1.0):
        self.points = np.asarray(pts, np.float32)
        self.point_confidences = np.ones(len(self.points), np.float32)
        self.score, self.track, self.tracking_score = score, None, 0.0

    def numpy(self):
        return self.points


def test_kalman_tracker_keeps_identities_through_a_crossing():
    """Two animals (3 nodes) passing each other in opposite directions, 80 frames of jittered detections in random order: the
    regular tracker supplies the first 10 frames, the filters take over (``init_done``) and every frame keeps the two
    identities apart."""
    rng = np.random.default_rng(1)
    shape = np.array([[0.0, 0.0], [6.0, 0.0], [12.0, 0.0]])
    tr = Tracker.make_tracker_by_name(tracker="simple", similarity="centroid", match="greedy", track_window=5, target_instance_count=2,
                                      kf_init_frame_count=10, kf_node_indices=[0, 1])
    ids = {}
    for t in range(80):
        a = shape + np.array([20.0 + 3.0 * t, 40.0 + 0.5 * t])          # passes b around t = 40, 30 px apart in y: never ambiguous
        b = shape + np.array([260.0 - 3.0 * t, 90.0 + 0.5 * t])          # (coincident detections are left untracked by design)
        insts = [_Inst(a + rng.normal(0, 0.3, a.shape)), _Inst(b + rng.normal(0, 0.3, b.shape))]
        order = rng.permutation(2)
        out = tr.track([insts[i] for i in order], t=t)           # the regular tracker returns tracked COPIES, the filters tag in place
        assert len(out) == 2
        for inst in out:
            who = "a" if np.abs(inst.numpy() - a).mean() < np.abs(inst.numpy() - b).mean() else "b"
            assert inst.track is not None, (t, who)
            ids.setdefault(who, inst.track)
            assert inst.track is ids[who], (t, who)
    assert tr.init_done and ids["a"] is not ids["b"]
    assert set(tr.kalman_tracker.tracks) == {ids["a"], ids["b"]}


def test_kalman_tracker_culls_and_resets_after_a_gap():
    """The pre-cull keeps the ``instance_count`` best instances before matching (a low-score spurious detection never gets a
    track); when BOTH filters have gone unmatched for more than ``reset_gap_size`` frames their identities are replaced by
    fresh tracks whose spawn frame is set at their first match (kalman.py:150-160, 236-242)."""
    rng = np.random.default_rng(3)
    shape = np.array([[0.0, 0.0], [6.0, 0.0]])
    tr = Tracker.make_tracker_by_name(tracker="simple", similarity="centroid", match="greedy", track_window=5, target_instance_count=2,
                                      pre_cull_to_target=True, kf_init_frame_count=6, kf_node_indices=[0, 1])

    def frame(t, jump=0.0):
        a = shape + np.array([20.0 + 2.0 * t + jump, 30.0])
        b = shape + np.array([20.0 + 2.0 * t + jump, 120.0])
        return a, b

    first = {}
    for t in range(20):
        a, b = frame(t)
        out = tr.track([_Inst(a + rng.normal(0, 0.2, a.shape)), _Inst(b + rng.normal(0, 0.2, b.shape)),
                        _Inst(shape + np.array([300.0, 300.0]), score=0.05)], t=t)
        assert len(out) == 2 and all(i.score > 0.5 for i in out)          # the spurious detection was culled
        for i in out:
            first.setdefault("a" if abs(i.numpy()[0, 1] - 30.0) < 10 else "b", i.track)
    assert tr.init_done
    assert set(tr.kalman_tracker.tracks) == {first["a"], first["b"]}


def test_kalman_tracker_replaces_identities_after_a_gap():
    """kalman.py:236-242: when more than one filter has gone unmatched for more than ``reset_gap_size`` frames (here two of three
    animals leave the field of view while the third keeps being tracked), those filters continue under fresh tracks whose
    spawn frame is set at their first match (:150-160, :228-232)."""
    shape = np.array([[0.0, 0.0], [6.0, 0.0]])
    tr = Tracker.make_tracker_by_name(tracker="simple", similarity="centroid", match="greedy", track_window=5, target_instance_count=3,
                                      kf_init_frame_count=6, kf_node_indices=[0, 1])
    pos = lambda t: [shape + np.array([20.0 + 2.0 * t, y]) for y in (30.0, 120.0, 210.0)]
    for t in range(15):
        out = tr.track([_Inst(p) for p in pos(t)], t=t)
    assert tr.init_done and len(tr.kalman_tracker.tracks) == 3
    by_y = {round(float(i.numpy()[0, 1])): i.track for i in out}
    stay, gone = by_y[30], {by_y[120], by_y[210]}
    for t in range(15, 23):                                        # only the first animal is visible
        out = tr.track([_Inst(pos(t)[0])], t=t)
        assert out[0].track is stay
    now = set(tr.kalman_tracker.tracks)
    assert stay in now and not (now & gone) and len(now) == 3      # the two lost identities were replaced ...
    assert all(t_.spawned_on == -1 for t_ in now if t_ is not stay)
    out = tr.track([_Inst(p) for p in pos(23)], t=23)              # ... and are picked up again by the same (coasting) filters
    tracks = {round(float(i.numpy()[0, 1])): i.track for i in out}
    assert tracks[30] is stay and tracks[120] in now and tracks[210] in now and tracks[120] is not tracks[210]
    assert tracks[120].spawned_on == 23 and tracks[210].spawned_on == 23
