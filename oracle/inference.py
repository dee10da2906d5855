"""Oracle of the reference's inference *layers* (test infrastructure only; never imported by the product).

Restates, on top of the other oracle modules (torch-CPU fp32 network + NumPy post-processing):
  sleap/nn/inference.py:1319-1380   SingleInstanceInferenceLayer.call
  sleap/nn/inference.py:1747-1966   CentroidCrop.call (incl. the per-sample top_k of :1879-1916)
  sleap/nn/inference.py:2059-2200   FindInstancePeaks.call
  sleap/nn/inference.py:2864-3003   BottomUpInferenceLayer.forward_pass / find_peaks / call
  sleap/nn/inference.py:940-967     InferenceLayer.preprocess (ensure_grayscale is decided by the
                                    network's input channel count, :905-911)
Pinned by the reference's own predictor tests on its trained fixture models
(tests/nn/test_inference.py:585-800; see tests/test_reference_models.py).
"""
import numpy as np

from oracle import convnet, paf_grouping as opg, peak_finding as opf, preprocess as opre, tf_ops

F = np.float32


def _heads(spec, weights, imgs, in_ch, input_scale, pad_stride, resize_img=True):
    x = opre.preprocess(imgs, ensure_gray=(in_ch == 1), input_scale=input_scale, pad_stride=pad_stride,
                        resize_img=resize_img)
    if spec.get("backbone") == "identity":       # the reference's layer tests wrap ``Lambda(lambda x: x)`` (test_inference.py:218-220)
        return {spec["heads"][0]["name"]: x}
    outs = convnet.model_forward(x, spec, weights)
    return {h["name"]: o for h, o in zip(spec["heads"], outs)}


def _scale_fix(points, stride, input_scale):
    points = (points * F(stride)).astype(F)
    if input_scale != 1.0:
        points = (points / F(input_scale) + F(0.5)).astype(F)     # :1368-1372
    return points


def single_instance_layer(imgs, spec, weights, in_ch=1, input_scale=1.0, pad_stride=1, peak_threshold=0.2,
                          refinement="integral", integral_patch_size=5, head="SingleInstanceConfmapsHead"):
    h = _heads(spec, weights, imgs, in_ch, input_scale, pad_stride)
    cms, offs = h[head], h.get("OffsetRefinementHead")
    stride = next(x["output_stride"] for x in spec["heads"] if x["name"] == head)
    if offs is None:
        pts, vals = opf.find_global_peaks(cms, peak_threshold, refinement, integral_patch_size)
    else:
        pts, vals = opf.find_global_peaks_with_offsets(cms, offs, peak_threshold)
    return {"instance_peaks": _scale_fix(pts, stride, input_scale)[:, None], "instance_peak_vals": vals[:, None],
            "confmaps": cms}


def centroid_crop_layer(imgs, spec, weights, crop_size, in_ch=1, input_scale=1.0, pad_stride=1, peak_threshold=0.2,
                        refinement="integral", integral_patch_size=5, max_instances=None):
    h = _heads(spec, weights, imgs, in_ch, input_scale, pad_stride)
    cms, offs = h["CentroidConfmapsHead"], h.get("OffsetRefinementHead")
    stride = spec["heads"][0]["output_stride"]
    if offs is None:
        pts, vals, sinds, _ = opf.find_local_peaks(cms, peak_threshold, refinement, integral_patch_size)
    else:
        pts, vals, sinds, _ = opf.find_local_peaks_with_offsets(cms, offs, peak_threshold)
    pts = _scale_fix(pts, stride, input_scale)
    if max_instances is not None and len(pts):
        keep = []
        for s in range(len(imgs)):
            idx = np.nonzero(sinds == s)[0]
            if max_instances < len(idx):
                idx = idx[np.argsort(-vals[idx], kind="stable")[:max_instances]]     # tf.math.top_k: ties -> lower index
            keep.append(idx)
        keep = np.concatenate(keep)
        pts, vals, sinds = pts[keep], vals[keep], sinds[keep]
    crop_offsets = (pts - F(crop_size / 2)).astype(F)
    if len(pts):
        bboxes = tf_ops.make_centered_bboxes(pts, crop_size, crop_size)
        crops = tf_ops.crop_bboxes(imgs, bboxes, sinds)
    else:
        crops = np.zeros((0, crop_size, crop_size, imgs.shape[3]), imgs.dtype)
    return {"centroids": pts, "centroid_vals": vals, "crop_sample_inds": sinds, "crops": crops,
            "crop_offsets": crop_offsets, "centroid_confmaps": cms}


def find_instance_peaks_layer(crops, crop_offsets, spec, weights, in_ch=1, input_scale=1.0, pad_stride=1,
                              peak_threshold=0.2, refinement="integral", integral_patch_size=5, resize_input_image=True):
    """``resize_input_image=False`` is how the top-down predictor builds the layer (:2405-2413): the crops were cut
    from full frames already resized by ``input_scale`` (CentroidCrop.precrop_resize / CentroidCropGroundTruth.input_scale),
    so only the coordinate fix-ups use ``input_scale``."""
    if len(crops) == 0:
        n = next(x["channels"] for x in spec["heads"] if x["name"] == "CenteredInstanceConfmapsHead")
        return np.zeros((0, n, 2), F), np.zeros((0, n), F)
    h = _heads(spec, weights, crops, in_ch, input_scale, pad_stride, resize_img=resize_input_image)
    cms, offs = h["CenteredInstanceConfmapsHead"], h.get("OffsetRefinementHead")
    stride = spec["heads"][0]["output_stride"]
    if offs is None:
        pts, vals = opf.find_global_peaks(cms, peak_threshold, refinement, integral_patch_size)
    else:
        pts, vals = opf.find_global_peaks_with_offsets(cms, offs, peak_threshold)
    pts = _scale_fix(pts, stride, input_scale)
    if crop_offsets is not None:
        pts = (pts + (crop_offsets[:, None, :] / F(input_scale)).astype(F)).astype(F)     # :2172-2177
    return pts, vals


def topdown_model(imgs, cspec, cweights, ispec, iweights, crop_size, c_in_ch=1, i_in_ch=1, c_input_scale=1.0,
                  i_input_scale=1.0, c_pad=1, i_pad=1, peak_threshold=0.2, refinement="integral",
                  integral_patch_size=5, max_instances=None):
    cc = centroid_crop_layer(imgs, cspec, cweights, crop_size, c_in_ch, c_input_scale, c_pad, peak_threshold,
                             refinement, integral_patch_size, max_instances)
    pts, vals = find_instance_peaks_layer(cc["crops"], cc["crop_offsets"], ispec, iweights, i_in_ch, i_input_scale,
                                          i_pad, peak_threshold, refinement, integral_patch_size)
    B = len(imgs)
    s = cc["crop_sample_inds"]
    return {"instance_peaks": [pts[s == b] for b in range(B)], "instance_peak_vals": [vals[s == b] for b in range(B)],
            "centroids": [cc["centroids"][s == b] for b in range(B)],
            "centroid_vals": [cc["centroid_vals"][s == b] for b in range(B)]}


def bottomup_layer(imgs, spec, weights, in_ch=1, input_scale=1.0, pad_stride=1, peak_threshold=0.2,
                   refinement="integral", integral_patch_size=5, **scorer_kwargs):
    h = _heads(spec, weights, imgs, in_ch, input_scale, pad_stride)
    cms, pafs, offs = h["MultiInstanceConfmapsHead"], h["PartAffinityFieldsHead"], h.get("OffsetRefinementHead")
    cm_stride = next(x["output_stride"] for x in spec["heads"] if x["name"] == "MultiInstanceConfmapsHead")
    paf_stride = next(x["output_stride"] for x in spec["heads"] if x["name"] == "PartAffinityFieldsHead")
    if offs is None:
        p, v, si, ci = opf.find_local_peaks(cms, peak_threshold, refinement, integral_patch_size)
    else:
        p, v, si, ci = opf.find_local_peaks_with_offsets(cms, offs, peak_threshold)
    p = (p * F(cm_stride)).astype(F)                                                  # :2922
    B = len(imgs)
    scorer = opg.PAFScorer(spec["part_names"], spec["edges"], paf_stride, **scorer_kwargs)
    inst, ivals, iscores, *_ = scorer.predict(pafs, [p[si == b] for b in range(B)], [v[si == b] for b in range(B)],
                                              [ci[si == b] for b in range(B)])
    if input_scale != 1.0:
        inst = [(x / F(input_scale) + F(0.5)).astype(F) for x in inst]                # :2980-2984
    return {"instance_peaks": inst, "instance_peak_vals": ivals, "instance_scores": iscores, "confmaps": cms,
            "part_affinity_fields": pafs}


def centroid_crop_ground_truth_layer(imgs, centroids, crop_size, input_scale=1.0):
    """sleap/nn/inference.py:743-809 CentroidCropGroundTruth.call: ``centroids`` = one (n, 2) array per sample."""
    full = imgs
    cents = [np.asarray(c, F).reshape(-1, 2) for c in centroids]
    if input_scale != 1.0:
        full = opre.resize_image(full, input_scale)                                   # :768-770
        cents = [(c * F(input_scale)).astype(F) for c in cents]
    sinds = np.concatenate([np.full(len(c), s, np.int32) for s, c in enumerate(cents)])
    pts = np.concatenate(cents)
    crop_offsets = (pts - F(crop_size / 2)).astype(F)
    crops = tf_ops.crop_bboxes(full, tf_ops.make_centered_bboxes(pts, crop_size, crop_size), sinds)
    return {"crops": crops, "crop_offsets": crop_offsets, "crop_sample_inds": sinds, "centroids": cents}


def match_points(points_gt, points_pr):
    """sleap/nn/utils.py:101-128 ``match_points``: LSAP on pairwise Euclidean distances."""
    from scipy.optimize import linear_sum_assignment
    d = np.linalg.norm(points_gt[:, None, :] - points_pr[None, :, :], axis=-1)
    d = np.where(np.isnan(d), np.inf, d)
    return linear_sum_assignment(np.where(np.isinf(d), 1e12, d))
