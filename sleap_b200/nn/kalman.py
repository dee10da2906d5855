"""Kalman-filter identity tracking (sleap/nn/tracker/kalman.py, sleap/nn/tracking.py:1235-1510).

The reference builds one ``pykalman.KalmanFilter`` per tracked identity (constant-velocity model over the coordinates of
a few reliable nodes), fits its noise covariances with 20 EM iterations on frames that the regular tracker has already
tracked, and from then on matches every frame's instances to the filters' predicted positions.

pykalman is a third-party dependency that is not part of this image, so ``KalmanFilter`` below restates the textbook
algorithms pykalman 0.9.5 implements for this use (Kalman filter, Rauch-Tung-Striebel smoother, Shumway-Stoffer EM for
``transition_covariance``, ``observation_covariance``, ``initial_state_mean``, ``initial_state_covariance`` -- pykalman's
default ``em_vars``; an observation with ANY masked component counts as missing, as in pykalman's ``_filter_correct``).
Numerical parity with pykalman itself is UNPINNED (nothing to run it against here); the matching logic around the
filters is pinned to the reference's known-answer tests (tests/nn/test_kalman.py) in tests/test_kalman.py.

Host-side, sequential, per-frame work on a handful of instances -- like the rest of the tracking step it is not a
device path.
"""
import itertools
from collections import defaultdict
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np

from sleap_b200.nn.tracking import (Match, Track, cull_frame_instances, first_choice_matching, greedy_matching)


class KalmanFilter:
    """Linear-Gaussian state-space model  x_{t+1} = A x_t + w (cov Q),  z_t = C x_t + v (cov R)."""

    def __init__(self, transition_matrices, observation_matrices, initial_state_mean, transition_covariance=None,
                 observation_covariance=None, initial_state_covariance=None):
        self.A = np.asarray(transition_matrices, np.float64)
        self.C = np.asarray(observation_matrices, np.float64)
        n, m = self.A.shape[0], self.C.shape[0]
        self.Q = np.eye(n) if transition_covariance is None else np.asarray(transition_covariance, np.float64)
        self.R = np.eye(m) if observation_covariance is None else np.asarray(observation_covariance, np.float64)
        self.mu0 = np.asarray(np.ma.filled(np.ma.asarray(initial_state_mean, np.float64), 0.0), np.float64)
        self.P0 = np.eye(n) if initial_state_covariance is None else np.asarray(initial_state_covariance, np.float64)

    # -- one step ---------------------------------------------------------------------------------------
    @staticmethod
    def _missing(obs) -> bool:
        if obs is np.ma.masked or obs is None:
            return True
        return bool(np.any(np.ma.getmaskarray(np.ma.asarray(obs)))) or bool(np.any(np.isnan(np.ma.filled(np.ma.asarray(obs, np.float64), np.nan))))

    def _correct(self, pm, pc, obs):
        if self._missing(obs):
            return pm, pc, np.zeros((len(pm), self.C.shape[0]))
        z = np.asarray(np.ma.filled(np.ma.asarray(obs, np.float64), 0.0), np.float64)
        S = self.C @ pc @ self.C.T + self.R
        K = pc @ self.C.T @ np.linalg.pinv(S)
        return pm + K @ (z - self.C @ pm), pc - K @ self.C @ pc, K

    def filter_update(self, filtered_state_mean, filtered_state_covariance, observation=None):
        """pykalman ``filter_update``: predict from the state at t, correct with the observation at t + 1."""
        pm = self.A @ np.asarray(filtered_state_mean, np.float64)
        pc = self.A @ np.asarray(filtered_state_covariance, np.float64) @ self.A.T + self.Q
        m, c, _ = self._correct(pm, pc, observation)
        return m, c

    # -- whole sequences --------------------------------------------------------------------------------
    def _filter(self, X):
        T, n = len(X), self.A.shape[0]
        pm, pc = np.zeros((T, n)), np.zeros((T, n, n))
        fm, fc = np.zeros((T, n)), np.zeros((T, n, n))
        for t in range(T):
            if t == 0:
                pm[t], pc[t] = self.mu0, self.P0
            else:
                pm[t] = self.A @ fm[t - 1]
                pc[t] = self.A @ fc[t - 1] @ self.A.T + self.Q
            fm[t], fc[t], _ = self._correct(pm[t], pc[t], X[t])
        return pm, pc, fm, fc

    def filter(self, X):
        X = np.ma.masked_invalid(np.ma.asarray(X, np.float64))
        _, _, fm, fc = self._filter(X)
        return fm, fc

    def _smooth(self, pm, pc, fm, fc):
        T = len(fm)
        sm, sc = fm.copy(), fc.copy()
        J = np.zeros((max(T - 1, 0),) + fc.shape[1:])
        for t in range(T - 2, -1, -1):
            J[t] = fc[t] @ self.A.T @ np.linalg.pinv(pc[t + 1])
            sm[t] = fm[t] + J[t] @ (sm[t + 1] - pm[t + 1])
            sc[t] = fc[t] + J[t] @ (sc[t + 1] - pc[t + 1]) @ J[t].T
        pair = np.zeros_like(sc)                     # pair[t] = Cov(x_t, x_{t-1} | all observations), t >= 1
        for t in range(1, T):
            pair[t] = sc[t] @ J[t - 1].T
        return sm, sc, pair

    def em(self, X, n_iter: int = 10):
        """Shumway-Stoffer EM over Q, R, the initial mean and the initial covariance (pykalman's default ``em_vars``)."""
        X = np.ma.masked_invalid(np.ma.asarray(X, np.float64))
        T = len(X)
        for _ in range(int(n_iter)):
            pm, pc, fm, fc = self._filter(X)
            sm, sc, pair = self._smooth(pm, pc, fm, fc)
            if T > 1:
                Q = np.zeros_like(self.Q)
                for t in range(T - 1):
                    err = sm[t + 1] - self.A @ sm[t]
                    VA = pair[t + 1] @ self.A.T
                    Q += np.outer(err, err) + self.A @ sc[t] @ self.A.T + sc[t + 1] - VA - VA.T
                self.Q = Q / (T - 1)
            R, n_obs = np.zeros_like(self.R), 0
            for t in range(T):
                if self._missing(X[t]):
                    continue
                z = np.asarray(np.ma.filled(X[t], 0.0), np.float64)
                err = z - self.C @ sm[t]
                R += np.outer(err, err) + self.C @ sc[t] @ self.C.T
                n_obs += 1
            if n_obs:
                self.R = R / n_obs
            self.mu0 = sm[0].copy()
            self.P0 = sc[0].copy()
        return self


# ---- matching helpers (kalman.py:436-667) ---------------------------------------------------------------
def match_dict_from_match_function(cost_matrix, row_items: list, column_items: list, match_function: Callable,
                                   key_by_column: bool = True) -> Dict[Any, Any]:
    """Keys from the columns (tracks), values from the rows (instances); the cheapest row wins a contested column."""
    match_dict, match_cost = {}, {}
    for i, j in match_function(cost_matrix):
        c = cost_matrix[i, j]
        if np.isfinite(c):
            key, val = (column_items[j], row_items[i]) if key_by_column else (row_items[i], column_items[j])
            if key not in match_dict or c < match_cost[key]:
                match_dict[key], match_cost[key] = val, c
    return match_dict


def match_tuples_from_match_function(cost_matrix, row_items: list, column_items: list, match_function) -> List[Tuple[Any, Any, float]]:
    return [(row_items[i], column_items[j], cost_matrix[i, j]) for (i, j) in match_function(cost_matrix) if np.isfinite(cost_matrix[i, j])]


def matches_from_match_tuples(match_tuples) -> List[Match]:
    return [Match(instance=inst, track=track, score=score) for (inst, track, score) in match_tuples]


def get_track_instance_matches(cost_matrix, instances: list, tracks: list, are_too_close_function: Callable) -> List[Match]:
    """Greedy matching in which an instance that lost its first-choice track to another instance only gets its second
    choice if the two instances are not "too close" (a redundant detection next to the winner)."""
    first = match_dict_from_match_function(cost_matrix, instances, tracks, first_choice_matching)
    greedy = matches_from_match_tuples(match_tuples_from_match_function(cost_matrix, instances, tracks, greedy_matching))
    good = []
    for m in greedy:
        if m.track in first:
            rival = first[m.track]
            if m.instance is not rival and m.instance != rival and are_too_close_function(m.instance, rival):
                continue
        good.append(m)
    return good


def remove_second_bests_from_cost_matrix(cost_matrix, thresh: float, invalid_val: float = np.nan) -> np.ndarray:
    """Clear a track's column when its best match is not better than the second best by ``thresh``; clear an instance's
    row when the same holds along the row or when its best entry has just been cleared."""
    cost_matrix = np.asarray(cost_matrix, np.float64)
    valid = np.full(cost_matrix.shape, True, dtype=bool)
    rows, cols = cost_matrix.shape
    for c in range(cols):
        col = cost_matrix[:, c]
        if np.all(np.isnan(col)):
            continue
        with np.errstate(invalid="ignore"):
            if (col < (col.min() + thresh)).sum() > 1:      # ndarray.min() propagates NaN like the reference's ``column.min()``
                valid[:, c] = False
    for r in range(rows):
        row = cost_matrix[r]
        if np.all(np.isnan(row)):
            continue
        k = int(row.argmin())
        with np.errstate(invalid="ignore"):
            if (row < (row[k] + thresh)).sum() > 1 or not valid[r, k]:
                valid[r] = False
    out = cost_matrix.copy()
    out[~valid] = invalid_val
    return out


def _track_points(inst, node_indices) -> np.ndarray:
    return np.asarray(inst.numpy(), np.float64)[node_indices, 0:2].flatten()


class BareKalmanTracker:
    """kalman.py:34-434: one filter per identity over ``node_indices``; ``init_filters`` fits them on tracked instances,
    ``track_frame`` assigns the tracks of one frame."""

    def __init__(self, node_indices: List[int], instance_count: int, instance_score_thresh: float = 0.3, reset_gap_size: int = 5):
        self.node_indices = list(node_indices)
        self.instance_count = instance_count
        self.instance_score_thresh = instance_score_thresh
        self.reset_gap_size = reset_gap_size
        self.kalman_filters: Dict[Track, KalmanFilter] = {}
        self.last_results: Dict[Track, Dict[str, Any]] = {}
        self.tracks: List[Track] = []
        self.last_frame_for_track: Dict[Track, int] = {}

    def init_filters(self, instances: Iterable):
        instances = list(instances)
        if not instances:
            raise ValueError("Kalman filter must be initialized with instances.")
        per_track = defaultdict(list)
        for inst in instances:
            per_track[inst.track].append(_track_points(inst, self.node_indices))
        filters, last, tracks = {}, {}, []
        for track, rows in per_track.items():
            frame_array = np.ma.masked_invalid(np.ma.asarray(rows, np.float64))
            n = frame_array[0].size                              # coordinates per frame: x0, y0, x1, y1, ...
            mu0 = np.zeros(2 * n)
            mu0[0::2] = np.ma.filled(frame_array[0], 0.0)        # state = (coord, velocity) pairs
            A = np.zeros((2 * n, 2 * n))
            C = np.zeros((n, 2 * n))
            for k in range(n):
                A[2 * k, 2 * k] = A[2 * k, 2 * k + 1] = A[2 * k + 1, 2 * k + 1] = 1.0
                C[k, 2 * k] = 1.0
            kf = KalmanFilter(transition_matrices=A, observation_matrices=C, initial_state_mean=mu0).em(frame_array, n_iter=20)
            means, covs = kf.filter(frame_array)
            tracks.append(track)
            filters[track] = kf
            last[track] = {"means": means[-1], "covariances": covs[-1]}
        self.kalman_filters, self.tracks, self.last_results, self.last_frame_for_track = filters, tracks, last, {}

    def replace_track(self, old_track: Track):
        """A long gap: the filter keeps running under a fresh identity (spawn frame set at its first match)."""
        new_track = Track(spawned_on=-1, name=old_track.name)
        self.kalman_filters[new_track] = self.kalman_filters.pop(old_track)
        self.tracks[self.tracks.index(old_track)] = new_track
        if old_track in self.last_results:
            self.last_results[new_track] = self.last_results.pop(old_track)

    def update_filters(self, track_instance_matches: Optional[dict] = None, only_update_matches: bool = False) -> dict:
        results = {}
        for track, kf in self.kalman_filters.items():
            if track_instance_matches and track in track_instance_matches:
                obs = np.ma.masked_invalid(np.ma.asarray(_track_points(track_instance_matches[track], self.node_indices)))
            elif only_update_matches:
                continue
            else:
                obs = np.ma.masked
            mean, cov = kf.filter_update(self.last_results[track]["means"], self.last_results[track]["covariances"], obs)
            results[track] = {"means": mean, "covariances": cov, "coordinate_means": np.asarray(mean[::2])}
        return results

    def get_instance_points_weight(self, instance) -> Tuple[np.ndarray, np.ndarray]:
        if not self.node_indices:
            raise ValueError("Kalman tracker must have node_indices set.")
        pts = _track_points(instance, self.node_indices)
        conf = getattr(instance, "point_confidences", None)
        w = np.ones(len(self.node_indices)) if conf is None else np.asarray(conf, np.float64)[self.node_indices]
        return pts, np.repeat(w, 2)

    @staticmethod
    def instance_points_match_cost(instance_points, instance_weights, expected_points) -> float:
        d = np.absolute(np.asarray(expected_points, np.float64) - instance_points)
        if np.all(np.isnan(d)):
            return np.nan
        return float(np.ma.average(np.ma.MaskedArray(d, mask=np.isnan(d)), weights=instance_weights))

    def get_mean_instance_distances(self, instances: list) -> dict:
        pts = {id(i): self.get_instance_points_weight(i)[0] for i in instances}

        def dist(a, b):
            d = np.absolute(pts[id(a)] - pts[id(b)])
            return float(np.nanmean(d)) if not np.all(np.isnan(d)) else np.nan

        return {(id(a), id(b)): dist(a, b) for a, b in itertools.combinations(instances, 2)}

    def get_too_close_checking_function(self, instances: list, dist_thresh: float) -> Callable:
        lookup = self.get_mean_instance_distances(instances)

        def too_close(a, b) -> bool:
            d = lookup[(id(a), id(b))] if (id(a), id(b)) in lookup else lookup[(id(b), id(a))]
            return d < dist_thresh

        return too_close

    def frame_cost_matrix(self, untracked_instances: list, filter_results: dict) -> np.ndarray:
        cost = np.full((len(untracked_instances), len(self.kalman_filters)), np.nan)
        for i, inst in enumerate(untracked_instances):
            if hasattr(inst, "score") and inst.score is not None and inst.score < self.instance_score_thresh:
                continue
            pts, w = self.get_instance_points_weight(inst)
            for j, track in enumerate(self.tracks):
                cost[i, j] = self.instance_points_match_cost(pts, w, filter_results[track]["coordinate_means"])
        return cost

    def track_frame(self, untracked_instances: list, frame_idx: int) -> list:
        filter_results = self.update_filters(only_update_matches=False)
        cost = self.frame_cost_matrix(untracked_instances, filter_results)
        if cost.size == 0 or np.all(np.isnan(cost)):
            return untracked_instances
        min_dist = float(np.nanmin(cost))
        cost = remove_second_bests_from_cost_matrix(cost, thresh=min_dist)
        too_close = self.get_too_close_checking_function(untracked_instances, dist_thresh=min_dist)
        matches = get_track_instance_matches(cost, instances=untracked_instances, tracks=self.tracks, are_too_close_function=too_close)
        self.last_results.update(self.update_filters({m.track: m.instance for m in matches}, only_update_matches=True))
        for m in matches:
            m.instance.track = m.track
            m.instance.tracking_score = float(m.score)
            self.last_frame_for_track[m.track] = frame_idx
            if m.track.spawned_on < 0:
                m.track.spawned_on = int(frame_idx)
        gap = self.tracks_with_gap(frame_idx)
        if len(gap) > 1:
            for track in gap:
                self.replace_track(track)
                self.last_frame_for_track.pop(track)
        return untracked_instances

    def tracks_with_gap(self, frame_idx: int) -> list:
        return [t for t, last in self.last_frame_for_track.items() if (frame_idx - last) > self.reset_gap_size]

    @property
    def last_frame_with_tracks(self) -> int:
        return max(self.last_frame_for_track.values(), default=0)


class KalmanInitSet:
    """tracking.py:1235-1309: contiguous well-tracked frames collected to initialise the filters."""

    def __init__(self, init_frame_count: int, instance_count: int, node_indices: List[int]):
        self.init_frame_count, self.instance_count, self.node_indices = init_frame_count, instance_count, list(node_indices)
        self.init_frames: list = []

    def add_frame_instances(self, instances: Iterable, frame_match=None):
        instances = list(instances)
        good = frame_match is None
        if frame_match is not None and frame_match.has_only_first_choice_matches:
            good = len([i for i in instances if self.is_usable_instance(i)]) >= self.instance_count
        if good:
            self.init_frames.append(instances)
        else:
            self.reset()                 # only CONTIGUOUS good frames count

    def reset(self):
        self.init_frames = []

    def is_usable_instance(self, instance) -> bool:
        if not getattr(instance, "track", None):
            return False
        return not np.any(np.isnan(np.asarray(instance.numpy(), np.float64)[self.node_indices, 0:2]))

    @property
    def is_set_ready(self) -> bool:
        return len(self.init_frames) >= self.init_frame_count

    @property
    def instances(self) -> list:
        return [i for frame in self.init_frames for i in frame if self.is_usable_instance(i)]


class KalmanTracker:
    """tracking.py:1311-1510: the regular tracker runs until ``init_frame_count`` contiguous good frames exist, the filters
    are fitted on them and take over; if they stop matching for ``re_init_after`` frames (after a cool-down) the regular
    tracker is used again to re-initialise them."""

    def __init__(self, init_tracker, init_set: KalmanInitSet, kalman_tracker: BareKalmanTracker, cull_function: Optional[Callable] = None,
                 init_frame_count: int = 10, re_init_cooldown: int = 100, re_init_after: int = 20, pre_tracked: bool = False):
        self.init_tracker, self.init_set, self.kalman_tracker = init_tracker, init_set, kalman_tracker
        self.cull_function = cull_function
        self.init_frame_count, self.re_init_cooldown, self.re_init_after = init_frame_count, re_init_cooldown, re_init_after
        self.init_done, self.pre_tracked = False, pre_tracked
        self.last_t, self.last_init_t = 0, 0

    @property
    def is_valid(self) -> bool:
        return self.pre_tracked or (self.init_tracker is not None and self.init_tracker.is_valid)

    @classmethod
    def make_tracker(cls, init_tracker, node_indices: List[int], instance_count: int, instance_iou_threshold: Optional[float] = 0.8,
                     init_frame_count: int = 10) -> "KalmanTracker":
        def cull_function(inst_list):
            cull_frame_instances(inst_list, instance_count=instance_count, iou_threshold=instance_iou_threshold)

        if getattr(init_tracker, "pre_cull_function", None) is None:
            init_tracker.pre_cull_function = cull_function
        return cls(init_tracker=init_tracker, kalman_tracker=BareKalmanTracker(node_indices=node_indices, instance_count=instance_count),
                   cull_function=cull_function, init_frame_count=init_frame_count,
                   init_set=KalmanInitSet(init_frame_count=init_frame_count, instance_count=instance_count, node_indices=node_indices))

    def track(self, untracked_instances: list, img_hw=(1, 1), img=None, t: Optional[int] = None, **kwargs) -> list:
        if t is None:
            t = self.last_t + 1
        self.last_t = t
        if self.cull_function:
            self.cull_function(untracked_instances)
        if not self.init_done:
            if self.pre_tracked:
                tracked, match_data = untracked_instances, None
            else:
                tracked = self.init_tracker.track(untracked_instances, img_hw=img_hw, img=img, t=t)
                match_data = self.init_tracker.last_matches
            self.init_set.add_frame_instances(tracked, match_data)
            if self.init_set.is_set_ready:
                self.kalman_tracker.init_filters(self.init_set.instances)
                self.init_done, self.last_init_t = True, t
        else:
            if self.pre_tracked:
                for inst in untracked_instances:
                    inst.track = None
            tracked = self.kalman_tracker.track_frame(untracked_instances, frame_idx=t)
        if self.init_done and (t - self.last_init_t) > self.re_init_cooldown:
            if self.kalman_tracker.last_frame_with_tracks < t - self.re_init_after:
                self.init_done = False
                self.init_set.reset()
                if self.init_tracker:
                    self.init_tracker.reset_candidates()
        return tracked

    def get_name(self) -> str:
        return f"kalman.{self.init_tracker.get_name()}"

    @property
    def uses_image(self) -> bool:
        return self.init_tracker.uses_image

    def final_pass(self, frames: list):
        self.init_tracker.final_pass(frames)
