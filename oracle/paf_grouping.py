"""Oracle restatement of sleap/nn/paf_grouping.py (NumPy float32 + SciPy LSAP + NetworkX).

Test infrastructure only.  Ragged tensors are represented as Python lists of per-sample
NumPy arrays.  Function names / argument order mirror the reference module.

Third-party routines (same as the reference calls, present in this image):
  * ``scipy.optimize.linear_sum_assignment`` (reference pin scipy>=1.4.1,<=1.9.0,
    environment.yml:34; call site sleap/nn/utils.py:93) -- rectangular shortest
    augmenting path (Crouse 2016).
  * ``networkx.topological_sort`` / ``bfs_edges`` (paf_grouping.py:1311-1314).
"""
from typing import Dict, List, Tuple

import numpy as np

from .tf_ops import F32, tf_linspace


def get_connection_candidates(peak_channel_inds_sample, skeleton_edges, n_nodes):
    """sleap/nn/paf_grouping.py:82-142.  Stable argsort by channel; src-major pairs."""
    ch = np.asarray(peak_channel_inds_sample, dtype=np.int32).reshape(-1)
    skeleton_edges = np.asarray(skeleton_edges, dtype=np.int32).reshape(-1, 2)
    peak_inds = np.argsort(ch, kind="stable").astype(np.int32)
    node_inds = ch[peak_inds]
    grouped = [peak_inds[node_inds == k] for k in range(n_nodes)]
    edge_inds, edge_peak_inds = [], []
    for k in range(skeleton_edges.shape[0]):
        s_list = grouped[skeleton_edges[k, 0]]
        d_list = grouped[skeleton_edges[k, 1]]
        s, d = np.meshgrid(s_list, d_list, indexing="ij")
        sd = np.stack([s, d], axis=2).reshape(-1, 2)
        edge_inds.append(np.full((sd.shape[0],), k, dtype=np.int32))
        edge_peak_inds.append(sd.astype(np.int32))
    if not edge_inds:
        return np.zeros((0,), np.int32), np.zeros((0, 2), np.int32)
    return np.concatenate(edge_inds), np.concatenate(edge_peak_inds).reshape(-1, 2)


def make_line_subs(peaks_sample, edge_peak_inds, edge_inds, n_line_points, pafs_stride):
    """sleap/nn/paf_grouping.py:145-222.  f32 linspace, /stride, round-half-even, int32."""
    peaks_sample = np.asarray(peaks_sample, dtype=F32).reshape(-1, 2)
    edge_peak_inds = np.asarray(edge_peak_inds, dtype=np.int32).reshape(-1, 2)
    edge_inds = np.asarray(edge_inds, dtype=np.int32).reshape(-1)
    src = peaks_sample[edge_peak_inds[:, 0]]
    dst = peaks_sample[edge_peak_inds[:, 1]]
    n = src.shape[0]
    XY = tf_linspace(src, dst, n_line_points)                 # (n, 2, P); dim1 = [x, y]
    XY = np.round((XY / F32(pafs_stride)).astype(F32)).astype(np.int32)   # np.round = half-even
    XY = XY[:, [1, 0], :]                                     # [row, col]
    line_subs = np.concatenate(
        [XY, np.broadcast_to(edge_inds.reshape(-1, 1, 1), (n, 1, n_line_points))], axis=1
    )
    line_subs = np.transpose(line_subs, (0, 2, 1))            # (n, P, 3)
    mul = np.array([1, 1, 2], dtype=np.int32).reshape(1, 1, 3)
    add = np.array([0, 0, 1], dtype=np.int32).reshape(1, 1, 3)
    return np.stack([line_subs * mul, line_subs * mul + add], axis=2).astype(np.int32)


def get_paf_lines(pafs_sample, peaks_sample, edge_peak_inds, edge_inds, n_line_points, pafs_stride):
    """sleap/nn/paf_grouping.py:225-275.

    Out-of-range subscripts: TF-CPU gather_nd raises, TF-GPU returns 0 (reference TODO at
    paf_grouping.py:197).  The oracle takes the GPU semantics (0), as SURVEY Appendix A.9.
    """
    pafs_sample = np.asarray(pafs_sample, dtype=F32)
    subs = make_line_subs(peaks_sample, edge_peak_inds, edge_inds, n_line_points, pafs_stride)
    H, W, Ch = pafs_sample.shape
    r, c, k = subs[..., 0], subs[..., 1], subs[..., 2]
    ok = (r >= 0) & (r < H) & (c >= 0) & (c < W) & (k >= 0) & (k < Ch)
    vals = pafs_sample[np.clip(r, 0, H - 1), np.clip(c, 0, W - 1), np.clip(k, 0, Ch - 1)]
    return np.where(ok, vals, F32(0)).astype(F32)


def compute_distance_penalty(spatial_vec_lengths, max_edge_length, dist_penalty_weight=1.0):
    """sleap/nn/paf_grouping.py:278-322."""
    l = np.asarray(spatial_vec_lengths, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        pen = np.minimum((F32(max_edge_length) / l).astype(F32) - F32(1), F32(0))
    return (pen * F32(dist_penalty_weight)).astype(F32)


def score_paf_lines(paf_lines_sample, peaks_sample, edge_peak_inds_sample, max_edge_length,
                    dist_penalty_weight=1.0):
    """sleap/nn/paf_grouping.py:325-403."""
    paf_lines_sample = np.asarray(paf_lines_sample, dtype=F32)
    peaks_sample = np.asarray(peaks_sample, dtype=F32).reshape(-1, 2)
    e = np.asarray(edge_peak_inds_sample, dtype=np.int32).reshape(-1, 2)
    src = peaks_sample[e[:, 0]]
    dst = peaks_sample[e[:, 1]]
    vec = (dst - src).astype(F32)
    length = np.sqrt((vec * vec).astype(F32).sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        unit = (vec / length).astype(F32)
    # (n, P, 2) @ (n, 2, 1): x*ux + y*uy in f32.
    line_scores = (paf_lines_sample[..., 0] * unit[:, None, 0]).astype(F32) + \
                  (paf_lines_sample[..., 1] * unit[:, None, 1]).astype(F32)
    pen = compute_distance_penalty(length, max_edge_length, dist_penalty_weight)[:, 0]
    P = paf_lines_sample.shape[1]
    mean = (line_scores.astype(F32).sum(axis=1, dtype=F32) / F32(P)).astype(F32)
    return (mean + pen).astype(F32)


def score_paf_lines_batch(pafs, peaks, peak_channel_inds, skeleton_edges, n_line_points,
                          pafs_stride, max_edge_length_ratio, dist_penalty_weight, n_nodes):
    """sleap/nn/paf_grouping.py:406-550.  peaks / peak_channel_inds: lists per sample."""
    pafs = np.asarray(pafs, dtype=F32)
    # reduce_max(tf.shape(pafs[0])) includes the channel dimension (paf_grouping.py:469-473).
    max_edge_length = F32(F32(max_edge_length_ratio) * F32(max(pafs.shape[1:])) * F32(pafs_stride))
    edge_inds, edge_peak_inds, line_scores = [], [], []
    for s in range(pafs.shape[0]):
        ei, epi = get_connection_candidates(peak_channel_inds[s], skeleton_edges, n_nodes)
        lines = get_paf_lines(pafs[s], peaks[s], epi, ei, n_line_points, pafs_stride)
        ls = score_paf_lines(lines, peaks[s], epi, max_edge_length, dist_penalty_weight)
        edge_inds.append(ei)
        edge_peak_inds.append(epi)
        line_scores.append(ls)
    return edge_inds, edge_peak_inds, line_scores


def linear_sum_assignment(cost_matrix):
    """sleap/nn/utils.py:79-98 (SciPy)."""
    from scipy.optimize import linear_sum_assignment as lsa

    r, c = lsa(np.asarray(cost_matrix))
    return r.astype(np.int32), c.astype(np.int32)


def match_candidates_sample(edge_inds_sample, edge_peak_inds_sample, line_scores_sample, n_edges):
    """sleap/nn/paf_grouping.py:553-670.  Returns edge-LOCAL src/dst indices.

    An infeasible cost matrix makes SciPy raise in the reference (it would crash); the
    oracle (and the CUDA path) define that case as "no matches for this edge".
    """
    edge_inds_sample = np.asarray(edge_inds_sample, dtype=np.int32).reshape(-1)
    edge_peak_inds_sample = np.asarray(edge_peak_inds_sample, dtype=np.int32).reshape(-1, 2)
    line_scores_sample = np.asarray(line_scores_sample, dtype=F32).reshape(-1)
    me, ms, md, msc = [], [], [], []
    for k in range(n_edges):
        sel = np.nonzero(edge_inds_sample == k)[0]
        epi = edge_peak_inds_sample[sel]
        ls = line_scores_sample[sel]
        _, first_s = np.unique(epi[:, 0], return_index=True)
        _, first_d = np.unique(epi[:, 1], return_index=True)
        n_src, n_dst = len(first_s), len(first_d)
        scores = ls.reshape(n_src, n_dst)
        cost = np.where(np.isnan(scores), F32(np.inf), -scores).astype(F32)
        if n_src == 0 or n_dst == 0:
            r = c = np.zeros((0,), np.int32)
        else:
            try:
                r, c = linear_sum_assignment(cost)
            except ValueError:
                r = c = np.zeros((0,), np.int32)
        me.append(np.full((len(r),), k, np.int32))
        ms.append(r)
        md.append(c)
        msc.append(scores[r, c].astype(F32))
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros((0,), dt))
    return cat(me, np.int32), cat(ms, np.int32), cat(md, np.int32), cat(msc, F32)


def match_candidates_batch(edge_inds, edge_peak_inds, line_scores, n_edges):
    """sleap/nn/paf_grouping.py:673-796."""
    out = ([], [], [], [])
    for s in range(len(edge_inds)):
        res = match_candidates_sample(edge_inds[s], edge_peak_inds[s], line_scores[s], n_edges)
        for o, r in zip(out, res):
            o.append(r)
    return out


def assign_connections_to_instances(connections, min_instance_peaks=0, n_nodes=None):
    """sleap/nn/paf_grouping.py:799-914.

    connections: dict {(src_node, dst_node): [(src_peak_ind, dst_peak_ind, score), ...]} in
    processing order.  Returns dict {(node_ind, peak_ind): instance_id}.
    """
    assign: Dict[Tuple[int, int], int] = dict()
    for (src_node, dst_node), conns in connections.items():
        for (sp, dp, _score) in conns:
            src_id = (int(src_node), int(sp))
            dst_id = (int(dst_node), int(dp))
            si = assign.get(src_id, None)
            di = assign.get(dst_id, None)
            if si is None and di is None:
                new = max(assign.values(), default=-1) + 1
                assign[src_id] = new
                assign[dst_id] = new
            elif si is not None and di is None:
                assign[dst_id] = si
            elif si is not None and di is not None:
                assign[dst_id] = si
                src_nodes = set(pid[0] for pid, inst in assign.items() if inst == si)
                dst_nodes = set(pid[0] for pid, inst in assign.items() if inst == di)
                if len(src_nodes.intersection(dst_nodes)) == 0:
                    for pid in assign:
                        if assign[pid] == di:
                            assign[pid] = si
    if min_instance_peaks > 0:
        if isinstance(min_instance_peaks, float):
            if n_nodes is None:
                nodes = set()
                for (a, b) in connections:
                    nodes.add(a)
                    nodes.add(b)
                n_nodes = len(nodes)
            min_instance_peaks = int(min_instance_peaks * n_nodes)
        ids, counts = np.unique(list(assign.values()), return_counts=True)
        cnt = {i: c for i, c in zip(ids, counts)}
        assign = {pid: inst for pid, inst in assign.items() if cnt[inst] >= min_instance_peaks}
    return assign


def make_predicted_instances(peaks, peak_scores, connections, instance_assignments):
    """sleap/nn/paf_grouping.py:917-981.  peaks/peak_scores: per-node lists."""
    ids, inv = np.unique(list(instance_assignments.values()), return_inverse=True)
    for pid, ind in zip(list(instance_assignments.keys()), inv):
        instance_assignments[pid] = int(ind)
    n_inst = len(ids)
    inst_scores = np.full((n_inst,), 0.0, dtype=F32)
    for (src_node, dst_node), conns in connections.items():
        for (sp, dp, score) in conns:
            sid = (int(src_node), int(sp))
            if sid in instance_assignments:
                inst_scores[instance_assignments[sid]] += F32(score)
    n_nodes = len(peaks)
    inst = np.full((n_inst, n_nodes, 2), np.nan, dtype=F32)
    inst_peak_scores = np.full((n_inst, n_nodes), np.nan, dtype=F32)
    for (node, pk), ind in instance_assignments.items():
        inst[ind, node, :] = peaks[node][pk]
        inst_peak_scores[ind, node] = peak_scores[node][pk]
    return inst, inst_peak_scores, inst_scores


def group_instances_sample(peaks_sample, peak_scores_sample, peak_channel_inds_sample,
                           match_edge_inds_sample, match_src_peak_inds_sample,
                           match_dst_peak_inds_sample, match_line_scores_sample, n_nodes,
                           sorted_edge_inds, edge_types, min_instance_peaks, min_line_scores=0.25):
    """sleap/nn/paf_grouping.py:984-1112.  edge_types: list of (src_node, dst_node)."""
    peaks_sample = np.asarray(peaks_sample, dtype=F32).reshape(-1, 2)
    peak_scores_sample = np.asarray(peak_scores_sample, dtype=F32).reshape(-1)
    ch = np.asarray(peak_channel_inds_sample, dtype=np.int32).reshape(-1)
    me = np.asarray(match_edge_inds_sample, dtype=np.int32).reshape(-1)
    ms = np.asarray(match_src_peak_inds_sample, dtype=np.int32).reshape(-1)
    md = np.asarray(match_dst_peak_inds_sample, dtype=np.int32).reshape(-1)
    msc = np.asarray(match_line_scores_sample, dtype=F32).reshape(-1)
    valid = msc >= F32(min_line_scores)
    me, ms, md, msc = me[valid], ms[valid], md[valid], msc[valid]
    peaks = [peaks_sample[ch == i] for i in range(n_nodes)]
    peak_scores = [peak_scores_sample[ch == i] for i in range(n_nodes)]
    connections = {}
    for edge_ind in sorted_edge_inds:
        sel = me == edge_ind
        et = tuple(int(v) for v in edge_types[int(edge_ind)])
        connections[et] = [(int(s), int(d), F32(sc)) for s, d, sc in zip(ms[sel], md[sel], msc[sel])]
    assign = assign_connections_to_instances(connections, min_instance_peaks=min_instance_peaks,
                                             n_nodes=n_nodes)
    return make_predicted_instances(peaks, peak_scores, connections, assign)


def group_instances_batch(peaks, peak_vals, peak_channel_inds, match_edge_inds, match_src_peak_inds,
                          match_dst_peak_inds, match_line_scores, n_nodes, sorted_edge_inds,
                          edge_types, min_instance_peaks, min_line_scores=0.25):
    """sleap/nn/paf_grouping.py:1115-1290."""
    out = ([], [], [])
    for s in range(len(peaks)):
        res = group_instances_sample(peaks[s], peak_vals[s], peak_channel_inds[s], match_edge_inds[s],
                                     match_src_peak_inds[s], match_dst_peak_inds[s],
                                     match_line_scores[s], n_nodes, sorted_edge_inds, edge_types,
                                     min_instance_peaks, min_line_scores=min_line_scores)
        for o, r in zip(out, res):
            o.append(r)
    return out


def toposort_edges(edge_types) -> Tuple[int, ...]:
    """sleap/nn/paf_grouping.py:1293-1315 (NetworkX)."""
    import networkx as nx

    edges = [(int(a), int(b)) for a, b in edge_types]
    dg = nx.DiGraph(edges)
    root = next(nx.topological_sort(dg))
    return tuple(edges.index(e) for e in nx.bfs_edges(dg, root))


class PAFScorer:
    """sleap/nn/paf_grouping.py:1318-1705 (attrs class restated as a plain class)."""

    def __init__(self, part_names, edges, pafs_stride, max_edge_length_ratio=0.25,
                 dist_penalty_weight=1.0, n_points=10, min_instance_peaks=0, min_line_scores=0.25):
        self.part_names = list(part_names)
        self.edges = [tuple(e) for e in edges]
        self.pafs_stride = pafs_stride
        self.max_edge_length_ratio = max_edge_length_ratio
        self.dist_penalty_weight = dist_penalty_weight
        self.n_points = n_points
        self.min_instance_peaks = min_instance_peaks
        self.min_line_scores = min_line_scores
        self.edge_inds = [(self.part_names.index(s), self.part_names.index(d)) for s, d in self.edges]
        self.edge_types = list(self.edge_inds)
        self.n_nodes = len(self.part_names)
        self.n_edges = len(self.edges)
        self.sorted_edge_inds = toposort_edges(self.edge_types)

    def score_paf_lines(self, pafs, peaks, peak_channel_inds):
        return score_paf_lines_batch(pafs, peaks, peak_channel_inds, self.edge_inds, self.n_points,
                                     self.pafs_stride, self.max_edge_length_ratio,
                                     self.dist_penalty_weight, self.n_nodes)

    def match_candidates(self, edge_inds, edge_peak_inds, line_scores):
        return match_candidates_batch(edge_inds, edge_peak_inds, line_scores, self.n_edges)

    def group_instances(self, peaks, peak_vals, peak_channel_inds, match_edge_inds,
                        match_src_peak_inds, match_dst_peak_inds, match_line_scores):
        return group_instances_batch(peaks, peak_vals, peak_channel_inds, match_edge_inds,
                                     match_src_peak_inds, match_dst_peak_inds, match_line_scores,
                                     self.n_nodes, self.sorted_edge_inds, self.edge_types,
                                     self.min_instance_peaks, min_line_scores=self.min_line_scores)

    def predict(self, pafs, peaks, peak_vals, peak_channel_inds):
        edge_inds, edge_peak_inds, line_scores = self.score_paf_lines(pafs, peaks, peak_channel_inds)
        me, ms, md, msc = self.match_candidates(edge_inds, edge_peak_inds, line_scores)
        inst, inst_peak_scores, inst_scores = self.group_instances(
            peaks, peak_vals, peak_channel_inds, me, ms, md, msc)
        return inst, inst_peak_scores, inst_scores, edge_inds, edge_peak_inds, line_scores
