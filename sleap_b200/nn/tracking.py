"""Identity tracking across frames: the step that follows the inference path.

Restates the default tracker of the reference -- candidates from the last ``track_window`` frames, one similarity
value per (instance, track), greedy or Hungarian assignment, new tracks for the unmatched, optional cap on the
number of tracks -- and the optical-flow variant that first shifts the candidates of earlier frames into the current frame
(Lucas-Kanade through OpenCV, exactly the library call the reference makes); the Kalman variant is not built:
  sleap/nn/tracker/components.py:33-196   similarity functions (instance, normalized, object keypoint, centroid, IoU)
  sleap/nn/tracker/components.py:198-226  hungarian_matching / greedy_matching, :637-647 first_choice_matching
  sleap/nn/tracker/components.py:229-313  nms_instances / nms_fast, :316-422 cull_instances / cull_frame_instances
  sleap/nn/tracker/components.py:457-634  Match, FrameMatches
  sleap/nn/tracking.py:442-492            SimpleCandidateMaker, SimpleMaxTracksCandidateMaker
  sleap/nn/tracking.py:33-86, 108-360     ShiftedInstance, FlowCandidateMaker (flow_shift_instances, saved shifts, pruning)
  sleap/nn/tracking.py:363-440            FlowMaxTracksCandidateMaker
  sleap/nn/tracking.py:542-844            Tracker.track / spawn / queues, :844-995 make_tracker_by_name
Sequential host code by nature (frame t depends on t-1); instances are anything with ``numpy()`` (n_nodes, 2),
``score`` and a settable ``track`` (``sleap_b200.nn.inference.PredictedInstance``).
"""
import copy
from collections import deque
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.optimize import linear_sum_assignment


class Track:
    """sleap/instance.py Track: identity token; compared by identity."""

    def __init__(self, spawned_on: int = 0, name: str = ""):
        self.spawned_on, self.name = spawned_on, name

    def __repr__(self):
        return f"Track(spawned_on={self.spawned_on}, name={self.name!r})"


def _pts(inst) -> np.ndarray:
    return np.asarray(inst.numpy(), np.float64)


def n_visible_points(inst) -> int:
    return int(np.sum(~np.isnan(_pts(inst)).any(axis=1)))


def centroid(inst) -> np.ndarray:
    """Median of the visible points (instance.py:867-875)."""
    return np.nanmedian(_pts(inst), axis=0)


def bounding_box(inst) -> np.ndarray:
    """[y1, x1, y2, x2] over the visible points (instance.py:878-886)."""
    p = _pts(inst)
    if np.isnan(p).all():
        return np.full(4, np.nan)
    return np.concatenate([np.nanmin(p, axis=0)[::-1], np.nanmax(p, axis=0)[::-1]])


# ---- similarities (larger = more alike) -----------------------------------------------------------
def instance_similarity(ref, query) -> float:
    r, q = _pts(ref), _pts(query)
    n_ref = np.sum(~np.isnan(r).any(axis=1))
    d2 = np.sum((q - r) ** 2, axis=1)
    return float(np.nansum(np.exp(-d2)) / n_ref)


def normalized_instance_similarity(ref, query, img_hw: Tuple[int, int]) -> float:
    scale = np.asarray((img_hw[1], img_hw[0]), np.float64)
    r, q = _pts(ref) / scale, _pts(query) / scale
    n_ref = np.sum(~np.isnan(r).any(axis=1))
    return float(np.nansum(np.exp(-np.sum((q - r) ** 2, axis=1))) / n_ref)


def factory_object_keypoint_similarity(keypoint_errors=None, score_weighting: bool = False, normalization_keypoints: str = "all") -> Callable:
    errors = np.asarray(1 if keypoint_errors is None or (hasattr(keypoint_errors, "__len__") and len(keypoint_errors) == 0) else keypoint_errors,
                        np.float64)
    with np.errstate(divide="ignore"):
        precision = 1.0 / (2.0 * errors ** 2)

    def object_keypoint_similarity(ref, query) -> float:
        r, q = _pts(ref), _pts(query)
        ws = 1.0
        if score_weighting:
            ws = np.asarray(getattr(ref, "point_confidences", np.ones(len(r))), np.float64) * \
                np.asarray(getattr(query, "point_confidences", np.ones(len(q))), np.float64)
        vis_r = ~np.isnan(r).any(axis=1)
        if normalization_keypoints == "ref":
            denom = int(vis_r.sum())
        elif normalization_keypoints == "union":
            denom = int(np.logical_and(vis_r, ~np.isnan(q).any(axis=1)).sum())
        else:
            denom = len(r)
        if denom == 0:
            return 0.0
        prec = precision
        if prec.size > 1 and prec.size != len(r):               # fit the per-keypoint errors to the skeleton
            prec = prec[:len(r)] if prec.size > len(r) else np.pad(prec, (0, len(r) - prec.size), "edge")
        d = np.sum((q - r) ** 2, axis=1) * prec
        return float(np.nansum(ws * np.exp(-d)) / denom)

    return object_keypoint_similarity


def centroid_distance(ref, query) -> float:
    return float(-np.linalg.norm(centroid(ref) - centroid(query)))


def compute_iou(a: Sequence[float], b: Sequence[float]) -> float:
    """sleap/nn/utils.py:45-76: inclusive-pixel IoU of [y1, x1, y2, x2] boxes."""
    iy1, ix1, iy2, ix2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    inter = max(ix2 - ix1 + 1, 0) * max(iy2 - iy1 + 1, 0)
    area = lambda t: (t[3] - t[1] + 1) * (t[2] - t[0] + 1)
    return float(inter / (area(a) + area(b) - inter))


def instance_iou(ref, query) -> float:
    return compute_iou(bounding_box(ref), bounding_box(query))


# ---- assignment -----------------------------------------------------------------------------------
def hungarian_matching(cost: np.ndarray) -> List[Tuple[int, int]]:
    rows, cols = linear_sum_assignment(cost)
    return list(zip(rows.tolist(), cols.tolist()))


def greedy_matching(cost: np.ndarray) -> List[Tuple[int, int]]:
    """Cheapest remaining (row, column) pair first; ties in ``argsort`` order of the flattened matrix."""
    order = np.argsort(cost, axis=None)
    used_r, used_c, out = set(), set(), []
    for flat in order.tolist():
        r, c = divmod(flat, cost.shape[1])
        if r in used_r or c in used_c:
            continue
        used_r.add(r); used_c.add(c)
        out.append((r, c))
    return out


def first_choice_matching(cost: np.ndarray) -> List[Tuple[int, int]]:
    return list(zip(range(len(cost)), cost.argmin(axis=1).tolist()))


# ---- suppression of duplicate detections ----------------------------------------------------------
def nms_fast(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float, target_count: Optional[int] = None) -> List[int]:
    """Score-ordered box suppression (overlap measured against the *other* box's area, +1 pixel convention), then, when
    fewer than ``target_count`` survive, suppressed boxes are handed back best score first.  The hand-back count is
    ``min(n_suppressed, n_kept - target_count)`` used as a slice end, exactly as the reference computes it
    (components.py:306-309; negative ends drop from the tail, which its tests rely on)."""
    boxes = np.asarray(boxes)
    scores = np.asarray(scores, np.float64)
    if len(boxes) == 0:
        return []
    if target_count and len(boxes) < target_count:
        return list(range(len(boxes)))
    boxes = boxes.astype(np.float64)
    x1, y1, x2, y2 = boxes.T
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    alive = list(np.argsort(scores))
    kept, dropped = [], []
    while alive:
        top = alive.pop()                                   # best remaining score
        kept.append(int(top))
        rest = np.asarray(alive, np.int64)
        if len(rest) == 0:
            break
        w = np.maximum(0, np.minimum(x2[top], x2[rest]) - np.maximum(x1[top], x1[rest]) + 1)
        h = np.maximum(0, np.minimum(y2[top], y2[rest]) - np.maximum(y1[top], y1[rest]) + 1)
        over = (w * h) / area[rest] > iou_threshold
        dropped.extend(int(i) for i in rest[over])
        alive = [int(i) for i in rest[~over]]
    if target_count and dropped and len(kept) < target_count:
        dropped.sort(key=lambda i: -scores[i])
        kept.extend(dropped[:min(len(dropped), len(kept) - target_count)])
    return kept


def nms_instances(instances: list, iou_threshold: float, target_count: Optional[int] = None):
    boxes = np.asarray([bounding_box(i) for i in instances])
    scores = np.asarray([i.score for i in instances])
    picks = set(nms_fast(boxes, scores, iou_threshold, target_count))
    return [x for i, x in enumerate(instances) if i in picks], [x for i, x in enumerate(instances) if i not in picks]


def cull_frame_instances(instances: list, instance_count: int, iou_threshold: Optional[float] = None) -> list:
    """At most ``instance_count`` instances in this frame: overlapping duplicates go first (when a threshold is given),
    then the lowest scores.  Modifies and returns the list (components.py:365-422)."""
    if not instances or len(instances) <= instance_count:
        return instances
    keep = list(instances)
    if iou_threshold:
        keep, extra = nms_instances(keep, iou_threshold, target_count=instance_count)
        for x in extra:
            instances.remove(x)
    if len(keep) > instance_count:
        for x in sorted(keep, key=lambda i: i.score)[:-instance_count]:
            instances.remove(x)
    return instances


def cull_instances(frames: list, instance_count: int, iou_threshold: Optional[float] = None):
    for lf in sorted(frames, key=lambda f: f.frame_idx):
        cull_frame_instances(lf.instances, instance_count, iou_threshold)


def connect_single_track_breaks(frames: list, instance_count: int) -> list:
    """sleap/nn/tracker/components.py:417-466: whenever exactly one track disappears and exactly one new track appears
    between a frame and the last frame that had ``instance_count`` tracks, the new track is merged into the lost one
    (for the rest of the video).  Modifies the frames in place."""
    if not frames:
        return frames
    fix_track_map = {}
    last_good = {inst.track for inst in frames[0].instances}
    for lf in frames:
        frame_tracks = {inst.track for inst in lf.instances}
        if frame_tracks & set(fix_track_map):
            for inst in lf.instances:
                if inst.track in fix_track_map and fix_track_map[inst.track] not in frame_tracks:
                    inst.track = fix_track_map[inst.track]
                    frame_tracks = {i.track for i in lf.instances}
        extra, missing = frame_tracks - last_good, last_good - frame_tracks
        if len(extra) == 1 and len(missing) == 1:
            for inst in lf.instances:
                if inst.track in extra:
                    old, new = inst.track, missing.pop()
                    fix_track_map[old] = new
                    inst.track = new
                    break
        elif len(frame_tracks) == instance_count:
            last_good = frame_tracks
    return frames


class TrackCleaner:
    """tracking.py:1513-1539: cull each frame to ``instance_count`` instances, then join single track breaks."""

    def __init__(self, instance_count: int, iou_threshold: Optional[float] = None):
        self.instance_count, self.iou_threshold = instance_count, iou_threshold

    def run(self, frames: list):
        cull_instances(frames, self.instance_count, self.iou_threshold)
        connect_single_track_breaks(frames, self.instance_count)


# ---- matches of one frame ---------------------------------------------------------------------------
class Match:
    def __init__(self, track, instance, score=None, is_first_choice=False):
        self.track, self.instance, self.score, self.is_first_choice = track, instance, score, is_first_choice


class FrameMatches:
    """Matches of one frame + whether each instance got the track it would have picked alone (components.py:470-634)."""

    def __init__(self, matches: List[Match], cost_matrix: np.ndarray, unmatched_instances: list):
        self.matches, self.cost_matrix, self.unmatched_instances = matches, cost_matrix, unmatched_instances

    @property
    def has_only_first_choice_matches(self) -> bool:
        return all(m.is_first_choice for m in self.matches)

    @classmethod
    def from_cost_matrix(cls, cost_matrix: np.ndarray, instances: list, tracks: list, matching_function: Callable):
        matches, taken = [], set()
        if instances and tracks:
            first = cost_matrix.argmin(axis=1)
            for i, j in matching_function(cost_matrix):
                taken.add(i)
                matches.append(Match(tracks[j], instances[i], -cost_matrix[i, j], bool(first[i] == j)))
        return cls(matches, cost_matrix, [x for i, x in enumerate(instances) if i not in taken])

    @classmethod
    def from_candidate_instances(cls, untracked_instances: list, candidate_instances: list, similarity_function: Callable,
                                 matching_function: Callable, robust_best_instance: float = 1.0):
        cost, tracks = np.ndarray((0,)), []
        if candidate_instances:
            by_track: Dict[object, list] = {}
            for c in candidate_instances:                       # insertion order = order of first appearance
                by_track.setdefault(c.track, []).append(c)
            tracks = list(by_track)
            sim = np.full((len(untracked_instances), len(tracks)), np.nan)
            for i, u in enumerate(untracked_instances):
                for j, tr in enumerate(tracks):
                    s = [similarity_function(u, c) for c in by_track[tr]]
                    sim[i, j] = np.quantile(s, robust_best_instance) if 0 < robust_best_instance < 1 else np.max(s)
            cost = -sim
            cost[np.isnan(cost)] = np.inf
        return cls.from_cost_matrix(cost, untracked_instances, tracks, matching_function)


# ---- candidate pools --------------------------------------------------------------------------------
class SimpleCandidateMaker:
    """Every instance of the last ``track_window`` frames with enough visible points."""

    uses_image = False

    def __init__(self, min_points: int = 0):
        self.min_points = min_points

    def get_candidates(self, track_matching_queue, **kw) -> list:
        return [x for item in track_matching_queue for x in item[1] if n_visible_points(x) >= self.min_points]


class SimpleMaxTracksCandidateMaker(SimpleCandidateMaker):
    """Per-track history; with ``max_tracking`` only the first ``max_tracks`` tracks are matchable."""

    def __init__(self, min_points: int = 0, max_tracks: Optional[int] = None):
        super().__init__(min_points)
        self.max_tracks = max_tracks

    def get_candidates(self, track_matching_queue_dict, max_tracking: bool, **kw) -> list:
        out, n = [], 0
        for _, hist in track_matching_queue_dict.items():
            if not max_tracking or n < self.max_tracks:
                n += 1
                out.extend(item[1] for item in hist if n_visible_points(item[1]) >= self.min_points)
        return out


class ShiftedInstance:
    """A reference instance moved into the current frame by optical flow (tracking.py:33-86): same track, new points
    (NaN where the flow lost the point), ``shift_score`` = -mean tracking error."""

    def __init__(self, points_array: np.ndarray, track, shift_score: float = 0.0, source=None):
        self.points_array = np.asarray(points_array, np.float64)
        self.track, self.shift_score, self.source = track, shift_score, source
        self.score = getattr(source, "score", float("nan"))

    def numpy(self):
        return self.points_array

    @classmethod
    def from_instance(cls, ref_instance, new_points_array=None, shift_score: float = 0.0):
        pts = _pts(ref_instance) if new_points_array is None else new_points_array
        return cls(pts, getattr(ref_instance, "track", None), shift_score, ref_instance)


def _ensure_u8(img: np.ndarray) -> np.ndarray:
    """normalization.ensure_int (sleap/nn/data/normalization.py:52-66): float images in [0, 1] -> uint8."""
    img = np.asarray(img)
    if img.dtype == np.uint8:
        return img
    if np.issubdtype(img.dtype, np.floating) and img.size and float(np.nanmax(img)) <= 1.0:
        img = img * 255.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


class FlowCandidateMaker:
    """Candidates = the instances of the last ``track_window`` frames, each flow-shifted into the current frame
    (tracking.py:108-360).  ``save_shifted_instances`` chains the shifts frame to frame instead of always starting from
    the original frame (:139-160)."""

    uses_image = True

    def __init__(self, min_points: int = 0, img_scale: float = 1.0, of_window_size: int = 21, of_max_levels: int = 3,
                 save_shifted_instances: bool = False, track_window: int = 5):
        self.min_points, self.img_scale = min_points, img_scale
        self.of_window_size, self.of_max_levels = of_window_size, of_max_levels
        self.save_shifted_instances, self.track_window = save_shifted_instances, track_window
        self.shifted_instances: Dict[Tuple[int, int], tuple] = {}        # (ref_t, t) -> (instances, img)

    def get_shifted_instances_from_earlier_time(self, ref_t: int, ref_img, ref_instances: list, t: int):
        for ti in reversed(range(ref_t, t)):
            if (ref_t, ti) in self.shifted_instances:
                insts, img = self.shifted_instances[(ref_t, ti)]
                if len(insts) > 0:
                    return img, insts
        return ref_img, ref_instances

    def get_shifted_instances(self, ref_instances: list, ref_img, ref_t: int, img, t: int) -> list:
        shifted = self.flow_shift_instances(ref_instances, ref_img, img, min_shifted_points=self.min_points, scale=self.img_scale,
                                            window_size=self.of_window_size, max_levels=self.of_max_levels)
        if self.save_shifted_instances:
            self.shifted_instances[(ref_t, t)] = (shifted, img)
        return shifted

    def prune_shifted_instances(self, t: int):
        if not self.save_shifted_instances:
            return
        for k in list(self.shifted_instances):
            if t - k[0] > self.track_window:
                del self.shifted_instances[k]

    def get_candidates(self, track_matching_queue, t: int, img, **kw) -> list:
        if img is None:
            raise ValueError("the flow tracker needs the frame image: Tracker.track(instances, img=frame, ...)")
        out = []
        self.prune_shifted_instances(t)
        for item in track_matching_queue:
            ref_t, ref_instances, ref_img = item[0], item[1], item[2]
            if self.save_shifted_instances:
                ref_img, ref_instances = self.get_shifted_instances_from_earlier_time(ref_t, ref_img, ref_instances, t)
            if len(ref_instances) > 0:
                out.extend(self.get_shifted_instances(ref_instances, ref_img, ref_t, img, t))
        return out

    @staticmethod
    def flow_shift_instances(ref_instances: list, ref_img, new_img, min_shifted_points: int = 0, scale: float = 1.0,
                             window_size: int = 21, max_levels: int = 3) -> list:
        """tracking.py:262-360: pyramidal Lucas-Kanade (cv2.calcOpticalFlowPyrLK, 30 iterations / eps 0.01) of every
        reference point; instances keep the points the flow found (> min_shifted_points of them)."""
        import cv2
        ref_img, new_img = _ensure_u8(ref_img), _ensure_u8(new_img)
        if ref_img.ndim > 2 and ref_img.shape[-1] == 1:
            ref_img, new_img = ref_img[..., 0], new_img[..., 0]
        if ref_img.ndim > 2 and ref_img.shape[-1] == 3:
            ref_img, new_img = cv2.cvtColor(ref_img, cv2.COLOR_BGR2GRAY), cv2.cvtColor(new_img, cv2.COLOR_BGR2GRAY)
        if scale != 1:
            ref_img = cv2.resize(ref_img, None, None, scale, scale)
            new_img = cv2.resize(new_img, None, None, scale, scale)
        ref_pts = [_pts(x) for x in ref_instances]
        shifted, status, errs = cv2.calcOpticalFlowPyrLK(
            np.ascontiguousarray(ref_img), np.ascontiguousarray(new_img), (np.concatenate(ref_pts, axis=0)).astype("float32") * scale, None,
            winSize=(window_size, window_size), maxLevel=max_levels,
            criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
        shifted = shifted / scale
        sections = np.cumsum([len(x) for x in ref_pts])[:-1]
        out = []
        for ref, pts, found, err in zip(ref_instances, np.split(shifted, sections, axis=0), np.split(status, sections, axis=0),
                                        np.split(errs, sections, axis=0)):
            if found.sum() > min_shifted_points:
                found = found.reshape(-1).astype(bool)
                pts = pts.astype(np.float64)
                pts[~found] = np.nan
                out.append(ShiftedInstance.from_instance(ref, new_points_array=pts, shift_score=-float(np.mean(err.reshape(-1)[found]))))
        return out


class FlowMaxTracksCandidateMaker(FlowCandidateMaker):
    """Flow candidates from the per-track history, at most ``max_tracks`` tracks (tracking.py:363-440)."""

    def __init__(self, max_tracks: Optional[int] = None, **kw):
        super().__init__(**kw)
        self.max_tracks = max_tracks

    @staticmethod
    def get_ref_instances(ref_t: int, ref_img, track_matching_queue_dict) -> list:
        out = []
        for _, hist in track_matching_queue_dict.items():
            out += [item[1] for item in hist if item[0] == ref_t and np.all(item[2] == ref_img)]
        return out

    def get_candidates(self, track_matching_queue_dict, max_tracking: bool, t: int, img, **kw) -> list:
        if img is None:
            raise ValueError("the flow tracker needs the frame image: Tracker.track(instances, img=frame, ...)")
        out, tracks = [], []
        self.prune_shifted_instances(t)
        for track, hist in track_matching_queue_dict.items():
            if not max_tracking or len(tracks) < self.max_tracks:
                tracks.append(track)
                for item in hist:
                    ref_t, ref_img = item[0], item[2]
                    ref_instances = self.get_ref_instances(ref_t, ref_img, track_matching_queue_dict)
                    if self.save_shifted_instances:
                        ref_img, ref_instances = self.get_shifted_instances_from_earlier_time(ref_t, ref_img, ref_instances, t)
                    if len(ref_instances) > 0:
                        out.extend(self.get_shifted_instances(ref_instances, ref_img, ref_t, img, t))
        return out


SIMILARITIES = dict(instance=instance_similarity, centroid=centroid_distance, iou=instance_iou,
                    normalized_instance=normalized_instance_similarity, object_keypoint=factory_object_keypoint_similarity)
MATCHERS = dict(hungarian=hungarian_matching, greedy=greedy_matching)
CANDIDATE_MAKERS = dict(simple=SimpleCandidateMaker, simplemaxtracks=SimpleMaxTracksCandidateMaker, flow=FlowCandidateMaker,
                        flowmaxtracks=FlowMaxTracksCandidateMaker)


class Tracker:
    """One ``track(instances)`` call per frame, in order (tracking.py:542-844)."""

    def __init__(self, track_window: int = 5, similarity_function: Optional[Callable] = instance_similarity,
                 matching_function: Callable = greedy_matching, candidate_maker=None, max_tracks: Optional[int] = None,
                 max_tracking: bool = False, min_new_track_points: int = 0, robust_best_instance: float = 1.0,
                 pre_cull_function: Optional[Callable] = None, target_instance_count: int = 0):
        self.track_window = track_window
        self.similarity_function, self.matching_function = similarity_function, matching_function
        self.candidate_maker = candidate_maker
        self.max_tracks, self.max_tracking = max_tracks, max_tracking
        self.min_new_track_points, self.robust_best_instance = min_new_track_points, robust_best_instance
        self.pre_cull_function, self.target_instance_count = pre_cull_function, target_instance_count
        self.track_matching_queue: deque = deque(maxlen=track_window)        # (t, [instances])
        self.track_matching_queue_dict: Dict[Track, deque] = {}               # track -> deque of (t, instance)
        self.spawned_tracks: List[Track] = []
        self.last_matches: Optional[FrameMatches] = None

    @property
    def is_valid(self):
        return self.similarity_function is not None

    @property
    def has_max_tracking(self) -> bool:
        return isinstance(self.candidate_maker, (SimpleMaxTracksCandidateMaker, FlowMaxTracksCandidateMaker))

    @property
    def uses_image(self) -> bool:
        return bool(getattr(self.candidate_maker, "uses_image", False))

    @property
    def unique_tracks_in_queue(self) -> List[Track]:
        if self.has_max_tracking:
            return list(self.track_matching_queue_dict)
        return list({x.track for item in self.track_matching_queue for x in item[1]})

    def reset_candidates(self):
        self.track_matching_queue = deque(maxlen=self.track_window)
        for tr in self.track_matching_queue_dict:
            self.track_matching_queue_dict[tr] = deque(maxlen=self.track_window)

    def _next_t(self) -> int:
        if self.has_max_tracking:
            if not self.track_matching_queue_dict:
                return 0
            longest = max(self.track_matching_queue_dict, key=lambda tr: len(self.track_matching_queue_dict[tr]))
            return self.track_matching_queue_dict[longest][-1][0] + 1
        return self.track_matching_queue[-1][0] + 1 if self.track_matching_queue else 0

    def track(self, untracked_instances: list, img_hw: Tuple[int, int] = (1, 1), img=None, t: Optional[int] = None) -> list:
        if self.candidate_maker is None:
            return untracked_instances
        sim = self.similarity_function
        if sim is normalized_instance_similarity:
            sim = lambda a, b: normalized_instance_similarity(a, b, img_hw=img_hw)
        if t is None:
            t = self._next_t()
        tracked: list = []
        if untracked_instances:
            if self.pre_cull_function:
                self.pre_cull_function(untracked_instances)
            if self.has_max_tracking:
                cands = self.candidate_maker.get_candidates(track_matching_queue_dict=self.track_matching_queue_dict,
                                                            max_tracking=self.max_tracking, t=t, img=img)
            else:
                cands = self.candidate_maker.get_candidates(track_matching_queue=self.track_matching_queue, t=t, img=img)
            fm = FrameMatches.from_candidate_instances(untracked_instances, cands, sim, self.matching_function, self.robust_best_instance)
            self.last_matches = fm
            for m in fm.matches:                                # matched: inherit the track
                x = copy.copy(m.instance)
                x.track, x.tracking_score = m.track, m.score
                tracked.append(x)
            for inst in fm.unmatched_instances:                 # unmatched: new tracks, unless the cap is reached
                if n_visible_points(inst) < self.min_new_track_points:
                    continue
                if self.has_max_tracking and self.max_tracking and len(self.track_matching_queue_dict) >= self.max_tracks:
                    break
                tr = Track(spawned_on=t, name=f"track_{len(self.spawned_tracks)}")
                self.spawned_tracks.append(tr)
                x = copy.copy(inst)
                x.track = tr
                tracked.append(x)
        keep_img = img if self.uses_image else None              # only the flow makers look at earlier frames (:780-800)
        if self.has_max_tracking:
            for x in tracked:
                if x.track in self.track_matching_queue_dict:
                    self.track_matching_queue_dict[x.track].append((t, x, keep_img))
                elif not self.max_tracking or len(self.track_matching_queue_dict) < self.max_tracks:
                    self.track_matching_queue_dict[x.track] = deque([(t, x, keep_img)], maxlen=self.track_window)
        else:
            self.track_matching_queue.append((t, tracked, keep_img))
        return tracked

    cleaner: Optional["TrackCleaner"] = None          # deprecated --clean_instance_count path (:924-927)
    post_connect_single_breaks: bool = False

    def final_pass(self, frames: list):
        """:816-835: post-processing after the last frame -- the (deprecated) cleaner, or the single-break joining."""
        if self.cleaner is not None:
            self.cleaner.run(frames)
        elif (self.target_instance_count or self.max_tracks) and self.post_connect_single_breaks:
            if not self.target_instance_count:
                self.target_instance_count = self.max_tracks
            connect_single_track_breaks(frames, self.target_instance_count)

    def get_name(self) -> str:
        return f"{type(self.candidate_maker).__name__}.{getattr(self.similarity_function, '__name__', 'none')}." \
               f"{getattr(self.matching_function, '__name__', 'none')}"

    @classmethod
    def make_tracker_by_name(cls, tracker: str = "simple", similarity: str = "instance", match: str = "greedy", track_window: int = 5,
                             robust: float = 1.0, min_new_track_points: int = 0, min_match_points: int = 0,
                             target_instance_count: int = 0, pre_cull_to_target: bool = False,
                             pre_cull_iou_threshold: Optional[float] = None, max_tracks: Optional[int] = None,
                             max_tracking: bool = False, oks_errors=None, oks_score_weighting: bool = False,
                             oks_normalization: str = "all", img_scale: float = 1.0, of_window_size: int = 21,
                             of_max_levels: int = 3, save_shifted_instances: bool = False, kf_init_frame_count: int = 0,
                             kf_node_indices: Optional[list] = None, post_connect_single_breaks: bool = False,
                             clean_instance_count: int = 0, clean_iou_threshold: Optional[float] = None, **kwargs) -> "Tracker":
        max_tracking = max_tracking if max_tracks else False
        if max_tracking and tracker in ("simple", "flow"):          # :882-884
            tracker += "maxtracks"
        if tracker.lower() == "none":
            return cls(track_window=track_window, similarity_function=None, matching_function=None, candidate_maker=None)
        if tracker not in CANDIDATE_MAKERS:
            raise ValueError(f"{tracker} is not a valid tracker.")
        if similarity not in SIMILARITIES:
            raise ValueError(f"{similarity} is not a valid tracker similarity function.")
        if match not in MATCHERS:
            raise ValueError(f"{match} is not a valid tracker matching function.")
        maker = CANDIDATE_MAKERS[tracker](min_points=min_match_points)
        if tracker in ("flow", "flowmaxtracks"):                   # :913-918
            maker.img_scale, maker.of_window_size, maker.of_max_levels = img_scale, of_window_size, of_max_levels
            maker.save_shifted_instances, maker.track_window = save_shifted_instances, track_window
        if tracker in ("simplemaxtracks", "flowmaxtracks"):
            maker.max_tracks = max_tracks
        sim = SIMILARITIES[similarity]
        if similarity == "object_keypoint":
            sim = factory_object_keypoint_similarity(oks_errors, oks_score_weighting, oks_normalization)
        pre_cull = None
        if target_instance_count and pre_cull_to_target:
            pre_cull = lambda insts: cull_frame_instances(insts, target_instance_count, pre_cull_iou_threshold)
        tracker_obj = cls(track_window=track_window, similarity_function=sim, matching_function=MATCHERS[match], candidate_maker=maker,
                          max_tracks=max_tracks, max_tracking=max_tracking, min_new_track_points=min_new_track_points,
                          robust_best_instance=robust, pre_cull_function=pre_cull, target_instance_count=target_instance_count)
        tracker_obj.post_connect_single_breaks = bool(post_connect_single_breaks)
        if clean_instance_count:
            tracker_obj.cleaner = TrackCleaner(instance_count=int(clean_instance_count), iou_threshold=clean_iou_threshold)
        # Kalman filters on top of the regular tracker (:955-991; sleap_b200/nn/kalman.py)
        if (max_tracks or target_instance_count) and kf_init_frame_count:
            if not kf_node_indices:
                raise ValueError("Kalman filter requires node indices for instance tracking.")
            if tracker in ("flow", "flowmaxtracks"):
                raise ValueError("Kalman filter requires simple tracker for initial tracking.")
            if similarity == "normalized_instance":
                raise ValueError("Kalman filter does not support normalized_instance_similarity.")
            from sleap_b200.nn.kalman import KalmanTracker
            return KalmanTracker.make_tracker(init_tracker=tracker_obj, init_frame_count=int(kf_init_frame_count),
                                              node_indices=[int(i) for i in kf_node_indices],
                                              instance_count=int(target_instance_count or max_tracks),
                                              instance_iou_threshold=pre_cull_iou_threshold)
        if kf_init_frame_count and not (max_tracks or target_instance_count):
            raise ValueError("Kalman filter requires max tracks or target instance count.")
        return tracker_obj


def run_tracker(frames: list, tracker: Tracker, images=None) -> list:
    """Track the predicted instances of ``frames`` (sorted by frame index) in place (tracking.py:1542-1580).
    ``images``: ``frame_idx -> image`` (mapping or callable), needed by the flow trackers."""
    frames = sorted(frames, key=lambda lf: lf.frame_idx)
    for lf in frames:
        img = None
        if images is not None:
            img = images(lf.frame_idx) if callable(images) else images[lf.frame_idx]
        hw = tuple(np.asarray(img).shape[:2]) if img is not None else (1, 1)
        lf.instances = tracker.track(list(lf.instances), img_hw=hw, img=img, t=lf.frame_idx)
    tracker.final_pass(frames)
    return frames
