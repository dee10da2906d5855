"""Tracker against the reference's component / max-tracking known answers
(tests/nn/test_tracker_components.py:104-170 nms, :173-232 FrameMatches + greedy matching, :235-430 max tracking)."""
import numpy as np
import pytest

from sleap_b200.nn import tracking as T
from sleap_b200.nn.inference import LabeledFrame, PredictedInstance


def test_nms():
    boxes = np.array([[10, 10, 20, 20], [10, 10, 15, 15], [30, 30, 40, 40], [32, 32, 42, 42]])
    assert sorted(T.nms_fast(boxes, np.array([1, 0.3, 1, 0.5]), iou_threshold=0.5)) == [0, 2]
    assert sorted(T.nms_fast(boxes, np.array([1, 0.3, 1, 0.5]), iou_threshold=0.5, target_count=3)) == [0, 2, 3]
    assert sorted(T.nms_fast(boxes, np.array([1, 0.5, 1, 0.3]), iou_threshold=0.5, target_count=3)) == [0, 1, 2]
    assert T.nms_fast(np.zeros((0, 4)), np.zeros(0), 0.5) == []


def _inst(a, b, score):
    return PredictedInstance.from_numpy(np.asarray([a, b], np.float32), np.ones(2, np.float32), score)


def test_nms_instances_to_remove():
    insts = [_inst((10, 10), (20, 20), 1), _inst((10, 10), (15, 15), 0.3), _inst((30, 30), (40, 40), 1), _inst((32, 32), (42, 42), 0.5)]
    keep, remove = T.nms_instances(insts, iou_threshold=0.5, target_count=3)
    assert len(remove) == 1 and remove[0] is insts[1] and len(keep) == 3
    culled = T.cull_frame_instances(list(insts), 2, iou_threshold=0.5)
    assert len(culled) == 2 and {id(x) for x in culled} == {id(insts[0]), id(insts[2])}


def test_frame_match_object():
    instances, tracks = ["instance a", "instance b"], ["track a", "track b"]
    fm = T.FrameMatches.from_cost_matrix(np.array([[10, 200], [75, 150]]), instances, tracks, T.greedy_matching)
    assert not fm.has_only_first_choice_matches and len(fm.matches) == 2
    assert (fm.matches[0].track, fm.matches[0].instance, fm.matches[0].score) == ("track a", "instance a", -10)
    assert (fm.matches[1].track, fm.matches[1].instance, fm.matches[1].score) == ("track b", "instance b", -150)
    fm = T.FrameMatches.from_cost_matrix(np.array([[10, 200], [150, 75]]), instances, tracks, T.greedy_matching)
    assert fm.has_only_first_choice_matches
    assert T.hungarian_matching(np.array([[10, 200], [75, 150]])) == [(0, 0), (1, 1)]
    assert T.first_choice_matching(np.array([[10, 200], [75, 150]])) == [(0, 0), (1, 0)]


def _make_insts(trx):
    out = []
    for frame in trx:
        out.append([PredictedInstance.from_numpy(np.array([[-0.1, -0.1], [0.0, 0.0], [0.1, 0.1]]) + np.array([[x, y]]), [1, 1, 1], 1)
                    for x, y in frame])
    return out


def _n_tracks(preds, **kw):
    tracker = T.Tracker.make_tracker_by_name(match="hungarian", track_window=2, **kw)
    tracked = [tracker.track(insts, img_hw=(1, 1)) for insts in preds]
    return len({id(inst.track) for frame in tracked for inst in frame}), tracked


CASES = {
    "large_gap_single_track": ([[(0, 0), (0, 1)], [(0.1, 0), (0.1, 1)], [(0.2, 0), (0.2, 1)], [(0.3, 0)], [(0.4, 0)], [(0.5, 0), (0.5, 1)],
                                [(0.6, 0), (0.6, 1)]], 3),
    "small_gap_on_both_tracks": ([[(0, 0), (0, 1)], [(0.1, 0), (0.1, 1)], [(0.2, 0), (0.2, 1)], [], [], [(0.5, 0), (0.5, 1)],
                                  [(0.6, 0), (0.6, 1)]], 4),
    "extra_detections": ([[(0, 0), (0, 1)], [(0.1, 0), (0.1, 1)], [(0.2, 0), (0.2, 1)], [(0.3, 0)], [(0.4, 0)], [(0.5, 0), (0.5, 1)],
                          [(0.6, 0), (0.6, 1), (0.6, 0.5)]], 4),
}


@pytest.mark.parametrize("name", list(CASES))
def test_max_tracking(name):
    """A gap longer than the window loses a track with the simple tracker; the max-tracks tracker keeps exactly 2."""
    trx, n_simple = CASES[name]
    assert _n_tracks(_make_insts(trx), tracker="simple")[0] == n_simple
    n, tracked = _n_tracks(_make_insts(trx), tracker="simplemaxtracks", max_tracks=2, max_tracking=True)
    assert n == 2
    # identities follow the animals: y = 0 and y = 1 never share a track
    by_track = {}
    for frame in tracked:
        for inst in frame:
            by_track.setdefault(id(inst.track), set()).add(round(float(inst.numpy()[1, 1])))
    assert all(len(v) == 1 or name == "extra_detections" for v in by_track.values())


@pytest.mark.parametrize("similarity", ["instance", "normalized_instance", "iou", "centroid", "object_keypoint"])
@pytest.mark.parametrize("match", ["greedy", "hungarian"])
@pytest.mark.parametrize("tracker", ["simple", "simplemaxtracks"])
def test_tracker_by_name(tracker, similarity, match):
    """Every (tracker, similarity, match) combination runs and keeps two well-separated animals apart
    (test_tracker_by_name, :46-68, on synthetic instead of recorded predictions)."""
    shape = np.array([[-5.0, -5.0], [0.0, 0.0], [5.0, 5.0]])          # 10 px animals moving 1 px per frame (boxes overlap in time)
    frames = [LabeledFrame(0, t, [PredictedInstance.from_numpy(shape + np.array([[10.0 + t, 10.0]]), [1, 1, 1], 1),
                                  PredictedInstance.from_numpy(shape + np.array([[60.0 - t, 40.0]]), [1, 1, 1], 1)]) for t in range(6)]
    tr = T.Tracker.make_tracker_by_name(tracker=tracker, similarity=similarity, match=match, max_tracks=2,
                                        max_tracking=tracker == "simplemaxtracks")
    out = T.run_tracker(frames, tr)
    ids = [[id(i.track) for i in lf.instances] for lf in out]
    assert all(len(set(x)) == 2 for x in ids) and all(x == ids[0] for x in ids)
    assert "." in tr.get_name()
    T.Tracker.make_tracker_by_name(tracker="none").track([])
    with pytest.raises(ValueError):
        T.Tracker.make_tracker_by_name(tracker="simple", max_tracks=2, kf_init_frame_count=10)     # Kalman variant: not built
    with pytest.raises(ValueError):
        T.Tracker.make_tracker_by_name(similarity="bogus")


def _blob_frame(centers, hw=(96, 128), sigma=4.0):
    """Textured uint8 frame: one Gaussian blob per animal node (what Lucas-Kanade can lock on to)."""
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]].astype(np.float64)
    img = np.zeros(hw)
    for cx, cy in centers:
        img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sigma ** 2))
    return np.clip(img * 255, 0, 255).astype(np.uint8)[..., None]


def test_flow_shift_instances_follows_motion():
    """FlowCandidateMaker.flow_shift_instances (tracking.py:262-360): the reference points move with the image content;
    a point the flow loses becomes NaN and does not count towards min_shifted_points."""
    shape = np.array([[-8.0, 0.0], [0.0, 0.0], [8.0, 0.0]])
    a, b = shape + [40.0, 40.0], shape + [43.0, 42.0]                      # whole animal moves by (+3, +2)
    ref = PredictedInstance.from_numpy(a, [1, 1, 1], 1.0)
    ref.track = T.Track(0, "t0")
    out = T.FlowCandidateMaker.flow_shift_instances([ref], _blob_frame(a), _blob_frame(b), window_size=21, max_levels=3)
    assert len(out) == 1 and out[0].track is ref.track
    assert np.allclose(out[0].numpy(), b, atol=0.35)
    assert out[0].shift_score <= 0
    half = T.FlowCandidateMaker.flow_shift_instances([ref], _blob_frame(a), _blob_frame(b), scale=0.5)
    assert np.allclose(half[0].numpy(), b, atol=1.0)
    assert T.FlowCandidateMaker.flow_shift_instances([ref], _blob_frame(a), _blob_frame(b), min_shifted_points=3) == []


@pytest.mark.parametrize("tracker,save", [("flow", False), ("flow", True), ("flowmaxtracks", False)])
def test_flow_tracker_keeps_identities_across_a_jump(tracker, save):
    """Two animals move 6 px per frame towards each other's positions: with the instance similarity (exp(-d^2)) the simple
    tracker cannot bridge a 6 px step, the flow tracker can because candidates are first shifted into the current frame."""
    shape = np.array([[-8.0, 0.0], [0.0, 0.0], [8.0, 0.0]])
    pos = lambda t: (shape + [30.0 + 6 * t, 30.0], shape + [100.0 - 6 * t, 70.0])
    frames = [LabeledFrame(0, t, [PredictedInstance.from_numpy(pos(t)[0], [1, 1, 1], 1.0), PredictedInstance.from_numpy(pos(t)[1], [1, 1, 1], 1.0)])
              for t in range(6)]
    imgs = {t: _blob_frame(np.concatenate(pos(t))) for t in range(6)}
    tr = T.Tracker.make_tracker_by_name(tracker=tracker, similarity="instance", match="hungarian", track_window=3, max_tracks=2,
                                        max_tracking=tracker == "flowmaxtracks", save_shifted_instances=save)
    assert tr.uses_image
    out = T.run_tracker(frames, tr, images=imgs)
    names = [[i.track.name for i in lf.instances] for lf in out]
    assert all(n == names[0] for n in names) and len(set(names[0])) == 2, names
    with pytest.raises(ValueError):
        tr.track(list(frames[0].instances), t=99)                          # flow needs the image
    simple = T.run_tracker([LabeledFrame(0, f.frame_idx, [PredictedInstance.from_numpy(i.numpy(), [1, 1, 1], 1.0) for i in f.instances])
                            for f in frames], T.Tracker.make_tracker_by_name(tracker="simple", similarity="instance", match="hungarian"))
    assert len({i.track.name for lf in simple for i in lf.instances}) >= 2


def test_connect_single_track_breaks_and_final_pass():
    """sleap/nn/tracker/components.py:417-466 through ``Tracker.final_pass`` (tracking.py:816-835): a track that is lost while
    exactly one new track appears is continued under the old identity for the rest of the video; two simultaneous
    changes are left alone; the option is off by default."""
    from sleap_b200.nn.tracking import Track, Tracker, connect_single_track_breaks

    class I:
        def __init__(self, track):
            self.track = track

    class F:
        def __init__(self, idx, tracks):
            self.frame_idx, self.instances = idx, [I(t) for t in tracks]

    a, b, c, d, e = (Track(0, n) for n in "abcde")
    frames = [F(0, [a, b]), F(1, [a, b]), F(2, [a, c]), F(3, [a, c]), F(4, [c, a]), F(5, [d, e])]
    connect_single_track_breaks(frames, 2)
    assert [[i.track.name for i in f.instances] for f in frames[:5]] == [["a", "b"], ["a", "b"], ["a", "b"], ["a", "b"], ["b", "a"]]
    assert [i.track.name for i in frames[5].instances] == ["d", "e"]          # two lost, two new: ambiguous, untouched
    frames = [F(0, [a, b]), F(1, [a, c])]
    t = Tracker.make_tracker_by_name(tracker="simple", target_instance_count=2)
    t.final_pass(frames)
    assert [i.track.name for i in frames[1].instances] == ["a", "c"]          # post_connect_single_breaks defaults to False
    t = Tracker.make_tracker_by_name(tracker="simple", max_tracking=True, max_tracks=2, post_connect_single_breaks=True)
    t.final_pass(frames)
    assert [i.track.name for i in frames[1].instances] == ["a", "b"] and t.target_instance_count == 2
