"""Oracle restatement of the reference's own synthetic-input generators (NumPy float32).

Test infrastructure only.  Sources:
  sleap/nn/data/utils.py:41-71,74-90        make_grid_vectors, gaussian_pdf
  sleap/nn/data/confidence_maps.py:10-112   make_confmaps, make_multi_confmaps
  sleap/nn/data/edge_maps.py:15-211         distance_to_edge, make_edge_maps, make_pafs,
                                            make_multi_pafs, get_edge_points
"""
import numpy as np

F32 = np.float32

# sleap/skeletons/flies13.json (node order / body edges).
FLIES13_NODES = ["head", "thorax", "abdomen", "wingL", "wingR", "forelegL", "forelegR",
                 "midlegL", "midlegR", "hindlegL", "hindlegR", "eyeL", "eyeR"]
FLIES13_EDGES = [("thorax", "head"), ("thorax", "abdomen"), ("thorax", "wingL"),
                 ("thorax", "wingR"), ("thorax", "forelegL"), ("thorax", "forelegR"),
                 ("thorax", "midlegL"), ("thorax", "midlegR"), ("thorax", "hindlegL"),
                 ("thorax", "hindlegR"), ("head", "eyeL"), ("head", "eyeR")]


def make_grid_vectors(image_height, image_width, output_stride=1):
    xv = np.arange(0, image_width, output_stride).astype(F32)
    yv = np.arange(0, image_height, output_stride).astype(F32)
    return xv, yv


def gaussian_pdf(x, sigma):
    x = np.asarray(x, dtype=F32)
    return np.exp(-(x * x) / F32(2 * F32(sigma) * F32(sigma))).astype(F32)


def make_confmaps(points, xv, yv, sigma):
    """confidence_maps.py:10-43: (H, W, n_nodes); NaN points -> all-zero channel."""
    points = np.asarray(points, dtype=F32).reshape(-1, 2)
    x = points[:, 0].reshape(1, 1, -1)
    y = points[:, 1].reshape(1, 1, -1)
    xv = np.asarray(xv, dtype=F32).reshape(1, -1, 1)
    yv = np.asarray(yv, dtype=F32).reshape(-1, 1, 1)
    with np.errstate(invalid="ignore"):
        cm = np.exp(-((xv - x) ** 2 + (yv - y) ** 2) / F32(2 * sigma ** 2)).astype(F32)
    return np.where(np.isnan(cm), F32(0), cm).astype(F32)


def make_multi_confmaps(instances, xv, yv, sigma):
    """confidence_maps.py:46-112: element-wise max over instances inside the image."""
    instances = np.asarray(instances, dtype=F32).reshape(-1, np.asarray(instances).shape[-2], 2)
    xv = np.asarray(xv, dtype=F32)
    yv = np.asarray(yv, dtype=F32)
    cms = np.zeros((len(yv), len(xv), instances.shape[1]), F32)
    lim = np.array([xv[-1], yv[-1]], F32).reshape(1, 1, 2)
    with np.errstate(invalid="ignore"):
        in_img = (instances > 0) & (instances < lim)
    in_img = np.any(np.all(in_img, axis=-1), axis=1)
    for pts in instances[in_img]:
        cms = np.maximum(cms, make_confmaps(pts, xv, yv, sigma))
    return cms


def distance_to_edge(points, edge_source, edge_destination):
    """edge_maps.py:15-80.  Returns the SQUARED distance (reference quirk).

    points: (..., 2); edge_source/destination: (E, 2) -> (..., E)
    """
    points = np.asarray(points, dtype=F32)
    es = np.asarray(edge_source, dtype=F32).reshape(-1, 2)
    ed = np.asarray(edge_destination, dtype=F32).reshape(-1, 2)
    direction = (ed - es).astype(F32)                                   # (E, 2)
    edge_length = np.maximum((direction * direction).sum(axis=1, dtype=F32), F32(1))
    rel = points[..., None, :] - es                                       # (..., E, 2)
    with np.errstate(invalid="ignore"):
        proj = (rel * direction).sum(axis=-1, dtype=F32) / edge_length   # (..., E)
        proj = np.clip(proj, F32(0), F32(1))
        d = (proj[..., None] * direction - rel).astype(F32)
        return (d * d).sum(axis=-1, dtype=F32).astype(F32)


def make_edge_maps(xv, yv, edge_source, edge_destination, sigma):
    """edge_maps.py:83-116: gaussian_pdf(squared distance) -> exp(-d^4 / (2 sigma^2))."""
    gx, gy = np.meshgrid(np.asarray(xv, F32), np.asarray(yv, F32))
    grid = np.stack([gx, gy], axis=-1)
    return gaussian_pdf(distance_to_edge(grid, edge_source, edge_destination), sigma)


def make_pafs(xv, yv, edge_source, edge_destination, sigma):
    """edge_maps.py:119-167: (H, W, E, 2)."""
    es = np.asarray(edge_source, dtype=F32).reshape(-1, 2)
    ed = np.asarray(edge_destination, dtype=F32).reshape(-1, 2)
    uv = (ed - es).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        uv = uv / np.sqrt((uv * uv).sum(axis=-1, keepdims=True, dtype=F32))
    em = make_edge_maps(xv, yv, es, ed, sigma)
    return (em[..., None] * uv.reshape(1, 1, -1, 2)).astype(F32)


def make_multi_pafs(xv, yv, edge_sources, edge_destinations, sigma):
    """edge_maps.py:170-211: SUM over instances (NaN -> 0)."""
    edge_sources = np.asarray(edge_sources, dtype=F32)
    edge_destinations = np.asarray(edge_destinations, dtype=F32)
    n_inst, n_edges = edge_sources.shape[0], edge_sources.shape[1]
    pafs = np.zeros((len(yv), len(xv), n_edges, 2), F32)
    for i in range(n_inst):
        paf = make_pafs(xv, yv, edge_sources[i], edge_destinations[i], sigma)
        pafs = pafs + np.where(np.isnan(paf), F32(0), paf)
    return pafs.astype(F32)


def get_edge_points(instances, edge_inds):
    """edge_maps.py:214-240."""
    instances = np.asarray(instances, dtype=F32)
    edge_inds = np.asarray(edge_inds, dtype=np.int32).reshape(-1, 2)
    return instances[:, edge_inds[:, 0]], instances[:, edge_inds[:, 1]]


def flies13_edge_inds():
    return [(FLIES13_NODES.index(a), FLIES13_NODES.index(b)) for a, b in FLIES13_EDGES]


def make_bottomup_frame(seed, height=1024, width=1024, n_instances=5, cm_stride=4, paf_stride=8,
                        cm_sigma=2.5, paf_sigma=75.0, nodes=None, edge_inds=None, noise=0.0,
                        centroid_margin=64.0, spread=40.0):
    """analytic multi-instance cms + pafs for post-processing parity.

    Returns (points (I, N, 2) f32 image px, cms (H/cs, W/cs, N), pafs (H/ps, W/ps, 2E)).
    """
    rng = np.random.default_rng(seed)
    n_nodes = len(nodes) if nodes is not None else len(FLIES13_NODES)
    if edge_inds is None:
        edge_inds = flies13_edge_inds()
    cent = rng.uniform(centroid_margin, min(height, width) - centroid_margin, size=(n_instances, 1, 2))
    pts = (cent + rng.normal(0.0, spread, size=(n_instances, n_nodes, 2))).astype(F32)
    pts[..., 0] = np.clip(pts[..., 0], 4, width - 5)
    pts[..., 1] = np.clip(pts[..., 1], 4, height - 5)
    xv, yv = make_grid_vectors(height, width, cm_stride)
    cms = make_multi_confmaps(pts, xv, yv, sigma=cm_sigma * cm_stride)
    xv8, yv8 = make_grid_vectors(height, width, paf_stride)
    es, ed = get_edge_points(pts, edge_inds)
    pafs = make_multi_pafs(xv8, yv8, es, ed, sigma=paf_sigma)
    pafs = pafs.reshape(pafs.shape[0], pafs.shape[1], -1)
    if noise > 0:
        cms = (cms + rng.normal(0, noise, size=cms.shape)).astype(F32)
        pafs = (pafs + rng.normal(0, noise, size=pafs.shape)).astype(F32)
    return pts, cms.astype(F32), pafs.astype(F32)
