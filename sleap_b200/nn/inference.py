"""Drop-in for the inference side of ``sleap.nn.inference`` (reference: sleap/nn/inference.py).

Kept surface (same names / attributes / dict contract, NumPy in and out):
  Predictor (:158-590), InferenceModel.predict / predict_on_batch (:989-1090),
  SingleInstanceInferenceLayer/Model/Predictor (:1229-1636),
  CentroidCrop (:1638-1967), FindInstancePeaks (:1969-2201), TopDownInferenceModel (:2246-2311),
  TopDownPredictor (:2314-2735), BottomUpInferenceLayer/Model (:2737-3053),
  BottomUpPredictor (:3055-3349), load_model (:4865).
All tensor work runs on the GPU through libsleapb200 (C-ABI); this file is orchestration only.
"""
import ctypes
import json
import os
import queue
import threading
from ctypes import c_int32, byref
from typing import Dict, List, Optional

import numpy as np

from sleap_b200 import _lib
from sleap_b200._lib import BottomUpParams, CentroidParams, GlobalParams, TopdownParams, f32, i32, ptr
from sleap_b200.nn import architectures as arch
from sleap_b200.nn import paf_grouping, peak_finding
from sleap_b200.nn.model import DeviceModel, PRECISION_FP16, PRECISION_FP32, load_weights, load_weights_npz

REFINE = peak_finding.REFINE


def _images_of(data):
    if isinstance(data, dict):
        return data["image"]
    return data


def _ragged_to_dense(rows: List[np.ndarray], inner_shape, dtype=np.float32):
    """RaggedTensor.to_tensor(default=NaN) to the bounding shape + row lengths (data/utils.py:118-146)."""
    n = max([len(r) for r in rows] + [0])
    out = np.full((len(rows), n) + tuple(inner_shape), np.nan, dtype)
    for i, r in enumerate(rows):
        if len(r):
            out[i, :len(r)] = r
    return out, np.asarray([len(r) for r in rows], np.int64)


class InferenceModel:
    """sleap/nn/inference.py:969-1171 (predict / predict_on_batch contract)."""

    def call(self, data):
        raise NotImplementedError

    def __call__(self, data):
        return self.call(data)

    def predict_on_batch(self, data, numpy: bool = True, **kwargs):
        """:1047-1090.  Ragged outputs come back NaN-padded with an ``n_valid`` key."""
        outs = self.call(data)
        if isinstance(data, dict):
            for k in ("video_ind", "frame_ind", "scale", "offset_x", "offset_y"):
                if k in data and k not in outs:
                    outs[k] = data[k]
        return outs

    def predict(self, data, numpy: bool = True, batch_size: int = 4, **kwargs):
        """:989-1045: iterate batches and concatenate (NaN-padding to the widest batch)."""
        imgs = np.asarray(_images_of(data))
        chunks = [self.predict_on_batch(imgs[i:i + batch_size]) for i in range(0, len(imgs), batch_size)]
        return _merge_batches(chunks)


def _merge_batches(chunks):
    out = {}
    if not chunks:
        return out
    for k in chunks[0]:
        if k.startswith("gathered_"):             # per-step arrays of the multi-GPU exchange: kept as a list
            out[k] = [c[k] for c in chunks]
            continue
        arrs = [np.asarray(c[k]) for c in chunks]
        if arrs[0].ndim >= 2 and np.issubdtype(arrs[0].dtype, np.floating):
            n = max(a.shape[1] for a in arrs)
            padded = []
            for a in arrs:
                if a.shape[1] < n:
                    pad = np.full((a.shape[0], n - a.shape[1]) + a.shape[2:], np.nan, a.dtype)
                    a = np.concatenate([a, pad], axis=1)
                padded.append(a)
            out[k] = np.concatenate(padded, axis=0)
        else:
            out[k] = np.concatenate(arrs, axis=0)
    return out


class InferenceLayer:
    """sleap/nn/inference.py:897-967: owns the device model and the preprocessing parameters.
    (uint8 -> float, gray/rgb, resize by input_scale and pad_to_stride run inside the device op-list.)"""

    def __init__(self, keras_model: DeviceModel, input_scale: float = 1.0, pad_to_stride: int = 1,
                 ensure_grayscale: Optional[bool] = None, ensure_float: bool = True):
        self.keras_model = keras_model      # attribute name kept from the reference
        self.input_scale = input_scale
        self.pad_to_stride = pad_to_stride
        if ensure_grayscale is None:
            ensure_grayscale = keras_model.cm.input_channels == 1
        self.ensure_grayscale = ensure_grayscale
        self.ensure_float = ensure_float

    @staticmethod
    def _prep(imgs):
        imgs = np.ascontiguousarray(imgs)
        if imgs.ndim == 3:
            imgs = imgs[..., None]
        if imgs.dtype != np.uint8:
            imgs = np.ascontiguousarray(imgs, dtype=np.float32)
        return imgs


def _find_head(model: DeviceModel, name: str):
    if name not in model.cm.head_buffers:
        return None
    return model.cm.head_buffers[name]


# ------------------------------------------------------------------------------------------
class SingleInstanceInferenceLayer(InferenceLayer):
    """sleap/nn/inference.py:1229-1380."""

    HEAD = "SingleInstanceConfmapsHead"

    def __init__(self, keras_model, input_scale=1.0, pad_to_stride=1, output_stride=None, peak_threshold=0.2,
                 refinement="local", integral_patch_size=5, return_confmaps=False, confmaps_ind=None,
                 offsets_ind=None, **kwargs):
        super().__init__(keras_model, input_scale=input_scale, pad_to_stride=pad_to_stride, **kwargs)
        self.confmaps_buffer = _find_head(keras_model, self.HEAD)
        if self.confmaps_buffer is None:
            raise ValueError(f"Index of the confidence maps output tensor must be specified if not named '{self.HEAD}'.")
        self.offsets_buffer = _find_head(keras_model, "OffsetRefinementHead")
        if output_stride is None:
            output_stride = keras_model.cm.head_strides[self.HEAD]
        self.output_stride = output_stride
        self.peak_threshold = peak_threshold
        self.refinement = refinement
        self.integral_patch_size = integral_patch_size
        self.return_confmaps = return_confmaps
        self._cfg_key = None

    def _configure(self, B, H, W, C):
        m = self.keras_model
        if not (m.configured_for and m.configured_for[0] >= B and m.configured_for[1:] == (H, W, C)):
            m.configure(B, H, W, C)
            self._cfg_key = None
        key = (m.configured_for, self.peak_threshold, self.refinement, self.integral_patch_size, self.input_scale)
        if self._cfg_key != key:
            p = GlobalParams(self.confmaps_buffer, -1 if self.offsets_buffer is None else self.offsets_buffer,
                             int(self.output_stride), float(self.peak_threshold), REFINE.get(self.refinement, 0),
                             int(self.integral_patch_size), float(self.input_scale))
            m.handle.call("sb_global_configure", m.model_id, byref(p))
            self._cfg_key = key

    def call(self, data, crop_offsets=None):
        imgs = self._prep(_images_of(data))
        B, H, W, C = imgs.shape
        self._configure(B, H, W, C)
        m = self.keras_model
        n_nodes = next(h["channels"] for h in m.spec["heads"] if h["name"] == self.HEAD)
        pts = np.zeros((B, n_nodes, 2), np.float32)
        vals = np.zeros((B, n_nodes), np.float32)
        co = None if crop_offsets is None else f32(crop_offsets).reshape(B, 2)
        m.handle.call("sb_infer_global", m.model_id, ptr(imgs), int(imgs.dtype == np.uint8), B, ptr(co), ptr(pts), ptr(vals))
        out = {"instance_peaks": pts[:, None], "instance_peak_vals": vals[:, None]}
        if self.return_confmaps:
            out["confmaps"] = m.forward(imgs, [self.HEAD])[0]
        return out


class SingleInstanceInferenceModel(InferenceModel):
    """sleap/nn/inference.py:1383-1415."""

    def __init__(self, single_instance_layer: SingleInstanceInferenceLayer):
        self.single_instance_layer = single_instance_layer

    def call(self, example):
        return self.single_instance_layer.call(example)


# ------------------------------------------------------------------------------------------
class CentroidCrop(InferenceLayer):
    """sleap/nn/inference.py:1638-1967: centroid net -> local peaks -> crops of the raw frames."""

    HEAD = "CentroidConfmapsHead"

    def __init__(self, keras_model, crop_size, input_scale=1.0, pad_to_stride=1, output_stride=None,
                 peak_threshold=0.2, refinement="local", integral_patch_size=5, return_confmaps=False,
                 return_crops=True, confmaps_ind=None, offsets_ind=None, max_instances=None,
                 precrop_resize=1.0, max_peaks_per_sample=256, **kwargs):
        super().__init__(keras_model, input_scale=input_scale, pad_to_stride=pad_to_stride, **kwargs)
        self.crop_size = crop_size
        self.confmaps_buffer = _find_head(keras_model, self.HEAD)
        if self.confmaps_buffer is None:
            raise ValueError(f"Index of the confidence maps output tensor must be specified if not named '{self.HEAD}'.")
        self.offsets_buffer = _find_head(keras_model, "OffsetRefinementHead")
        self.output_stride = output_stride or keras_model.cm.head_strides[self.HEAD]
        self.peak_threshold = peak_threshold
        self.refinement = refinement
        self.integral_patch_size = integral_patch_size
        self.return_confmaps = return_confmaps
        self.return_crops = return_crops
        self.max_instances = max_instances
        self.precrop_resize = precrop_resize
        self.max_peaks_per_sample = max_peaks_per_sample
        self._cfg_key = None
        self._resizer = None

    def call(self, inputs):
        full_imgs = self._prep(_images_of(inputs))
        B, H, W, C = full_imgs.shape
        m = self.keras_model
        m.configure(B, H, W, C)
        key = (m.configured_for, self.peak_threshold, self.refinement, self.integral_patch_size, self.input_scale)
        if self._cfg_key != key:
            p = CentroidParams(self.confmaps_buffer, -1 if self.offsets_buffer is None else self.offsets_buffer,
                               int(self.output_stride), float(self.peak_threshold), REFINE.get(self.refinement, 0),
                               int(self.integral_patch_size), float(self.input_scale), int(self.max_peaks_per_sample))
            m.handle.call("sb_centroid_configure", m.model_id, byref(p))
            self._cfg_key = key
        cap = B * self.max_peaks_per_sample
        pts = np.zeros((cap, 2), np.float32)
        vals = np.zeros((cap,), np.float32)
        sinds = np.zeros((cap,), np.int32)
        n = c_int32(0)
        flags = np.zeros((B,), np.int32)
        m.handle.call("sb_infer_centroids", m.model_id, ptr(full_imgs), int(full_imgs.dtype == np.uint8), B, ptr(pts),
                      ptr(vals), ptr(sinds), byref(n), ptr(flags))
        k = n.value
        pts, vals, sinds = pts[:k].copy(), vals[:k].copy(), sinds[:k].copy()
        if self.precrop_resize != 1.0:                       # :1836-1841 (the returned centroids stay in the resized frame, as in the reference)
            if self._resizer is None:
                from sleap_b200.nn.model import FrameResizer
                self._resizer = FrameResizer(m.handle)
            full_imgs = self._resizer(full_imgs, self.precrop_resize)
            H, W = full_imgs.shape[1:3]
            pts = (pts * np.float32(self.precrop_resize)).astype(np.float32)
        if k > 0 and self.max_instances is not None:
            keep = []
            for s in range(B):   # tf.math.top_k: descending score, ties keep the lower index (:1879-1894)
                idx = np.nonzero(sinds == s)[0]
                if self.max_instances < len(idx):
                    order = np.argsort(-vals[idx], kind="stable")[: self.max_instances]
                    idx = idx[order]
                keep.append(idx)
            keep = np.concatenate(keep) if keep else np.zeros((0,), np.int64)
            pts, vals, sinds = pts[keep], vals[keep], sinds[keep]
            k = len(keep)
        crop_offsets = (pts - np.float32(self.crop_size / 2)).astype(np.float32)
        out = dict(centroids=[pts[sinds == s] for s in range(B)], centroid_vals=[vals[sinds == s] for s in range(B)],
                   flags=flags)
        if self.return_crops:
            if k > 0:
                crops = np.zeros((k, self.crop_size, self.crop_size, C), full_imgs.dtype)
                m.handle.call("sb_crop_centered", ptr(full_imgs), int(full_imgs.dtype == np.uint8), B, H, W, C,
                              ptr(f32(pts)), ptr(i32(sinds)), k, self.crop_size, self.crop_size, ptr(crops))
            else:
                crops = np.zeros((0, self.crop_size, self.crop_size, C), full_imgs.dtype)
            out["crops"] = crops
            out["crop_offsets"] = crop_offsets
            out["crop_sample_inds"] = sinds
            out["samples"] = B
        return out


class FindInstancePeaks(SingleInstanceInferenceLayer):
    """sleap/nn/inference.py:1969-2201: centered-instance net on crops -> global peaks (+ crop offsets)."""

    HEAD = "CenteredInstanceConfmapsHead"

    def __init__(self, keras_model, max_crops_per_call=64, **kwargs):
        super().__init__(keras_model, **kwargs)
        self.max_crops_per_call = max_crops_per_call

    def call(self, inputs):
        if isinstance(inputs, dict):
            crops = inputs["crops"]
        else:
            crops, inputs = inputs, {}
        crops = self._prep(crops)
        n = crops.shape[0]
        if "crop_sample_inds" in inputs:
            samples, sinds = inputs["samples"], np.asarray(inputs["crop_sample_inds"])
        else:
            samples, sinds = n, np.arange(n)
        co = inputs.get("crop_offsets")
        m = self.keras_model
        n_nodes = next(h["channels"] for h in m.spec["heads"] if h["name"] == self.HEAD)
        pts = np.zeros((n, n_nodes, 2), np.float32)
        vals = np.zeros((n, n_nodes), np.float32)
        for i in range(0, n, self.max_crops_per_call):
            sl = slice(i, min(n, i + self.max_crops_per_call))
            sub = np.ascontiguousarray(crops[sl])
            # keep one device configuration for all chunk sizes
            self._configure(self.max_crops_per_call, *sub.shape[1:])
            o = super().call(sub, crop_offsets=None if co is None else np.asarray(co)[sl])
            pts[sl], vals[sl] = o["instance_peaks"][:, 0], o["instance_peak_vals"][:, 0]
        out = {"instance_peaks": [pts[sinds == s] for s in range(samples)],
               "instance_peak_vals": [vals[sinds == s] for s in range(samples)]}
        for k in ("centroids", "centroid_vals"):
            if k in inputs:
                out[k] = inputs[k]
        return out


class CentroidCropGroundTruth:
    """sleap/nn/inference.py:723-809: stands in for a centroid model -- crops of ``crop_size`` around the
    ground-truth centroids of a labels example (``example_gt["centroids"]``: one (n, 2) array per sample, made by
    ``LabelsReader(with_centroids=True)`` = InstanceCentroidFinder).  The crop itself is the device kernel."""

    def __init__(self, crop_size: int, input_scale: float = 1.0, handle=None):
        self.crop_size = crop_size
        self.input_scale = input_scale
        self.handle = handle
        self._resizer = None

    def call(self, example_gt):
        from sleap_b200 import _lib
        full_imgs = np.ascontiguousarray(example_gt["image"])
        cents = [f32(c).reshape(-1, 2) for c in example_gt["centroids"]]
        if self.input_scale != 1.0:                          # :768-770: resized frames, centroids scaled with them
            if self._resizer is None:
                from sleap_b200.nn.model import FrameResizer
                self._resizer = FrameResizer(self.handle or _lib.default_handle())
            full_imgs = self._resizer(full_imgs, self.input_scale)
            cents = [(c * np.float32(self.input_scale)).astype(np.float32) for c in cents]
        B, H, W, C = full_imgs.shape
        sinds = np.concatenate([np.full(len(c), s, np.int32) for s, c in enumerate(cents)]) if cents else np.zeros(0, np.int32)
        pts = np.concatenate(cents) if cents else np.zeros((0, 2), np.float32)
        k = len(pts)
        crop_offsets = (pts - np.float32(self.crop_size / 2)).astype(np.float32)          # :777
        crops = np.zeros((k, self.crop_size, self.crop_size, C), full_imgs.dtype)
        if k > 0:
            h = self.handle or _lib.default_handle()
            h.call("sb_crop_centered", ptr(full_imgs), int(full_imgs.dtype == np.uint8), B, H, W, C, ptr(f32(pts)), ptr(i32(sinds)), k,
                   self.crop_size, self.crop_size, ptr(crops))
        return dict(crops=crops, crop_offsets=crop_offsets, crop_sample_inds=sinds, samples=B, centroids=cents,
                    centroid_vals=[np.ones(len(c), np.float32) for c in cents])


class FindInstancePeaksGroundTruth:
    """sleap/nn/inference.py:812-893: stands in for a centered-instance model -- every centroid gets the
    ground-truth instance whose nearest node is closest to it (``example_gt["instances"]``: one (n, nodes, 2)
    array per sample); peak values are 1."""

    def call(self, example_gt, crop_output):
        peaks, vals = [], []
        for inst, cent in zip(example_gt["instances"], crop_output["centroids"]):
            inst, cent = f32(inst), f32(cent).reshape(-1, 2)
            n_nodes = inst.shape[1] if inst.ndim == 3 else 0
            rows = []
            if len(inst) and len(cent):
                import warnings
                with np.errstate(invalid="ignore"), warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    d = np.sqrt(((inst[None] - cent[:, None, None, :]) ** 2).sum(-1))      # (n_centroids, n_insts, n_nodes)
                    # reduce_min over nodes (:860): Eigen's scalar min keeps the accumulator when the new value is NaN,
                    # so invisible nodes are skipped and only an all-NaN instance yields NaN
                    # (tests/nn/test_inference.py:168-209 "GT instances have NaNs")
                    d = np.nanmin(d, axis=-1)
                for c in range(len(cent)):
                    if np.all(np.isnan(d[c])):
                        continue                                                           # :866-868 all-NaN rows are dropped
                    best = 0                                                               # tf.argmin: first index; NaN never wins a "<"
                    for j in range(1, d.shape[1]):
                        if d[c, j] < d[c, best]:
                            best = j
                    rows.append(inst[best])
            peaks.append(np.stack(rows) if rows else np.zeros((0, n_nodes, 2), np.float32))
            vals.append(np.ones(peaks[-1].shape[:2], np.float32))
        return dict(centroids=crop_output["centroids"], centroid_vals=crop_output["centroid_vals"], instance_peaks=peaks,
                    instance_peak_vals=vals)


class CentroidInferenceModel(InferenceModel):
    """sleap/nn/inference.py:2203-2244: the first stage of the top-down path on its own (centroids only)."""

    def __init__(self, centroid_crop):
        self.centroid_crop = centroid_crop

    def call(self, example):
        if isinstance(example, np.ndarray):
            example = dict(image=example)
        out = self.centroid_crop.call(example)
        ce, nv = _ragged_to_dense(out["centroids"], (2,))
        cv, _ = _ragged_to_dense(out["centroid_vals"], ())
        res = {"centroids": ce, "centroid_vals": cv, "n_valid": nv}
        for k in ("crops", "crop_offsets", "crop_sample_inds", "flags"):
            if k in out:
                res[k] = out[k]
        return res


class TopDownInferenceModel(InferenceModel):
    """sleap/nn/inference.py:2246-2311."""

    def __init__(self, centroid_crop, instance_peaks):
        self.centroid_crop = centroid_crop
        self.instance_peaks = instance_peaks
        self.fused = True            # one device pipeline (sb_infer_topdown) when both stages are device models
        self._td_key = None

    def _can_fuse(self):
        cc, fp = self.centroid_crop, self.instance_peaks
        return (self.fused and type(cc) is CentroidCrop and type(fp) is FindInstancePeaks and cc.precrop_resize == 1.0 and
                cc.return_crops and not cc.return_confmaps and not fp.return_confmaps and cc.keras_model.handle is fp.keras_model.handle
                and fp.keras_model.input_scale == 1.0)

    def _call_fused(self, imgs):
        """sb_infer_topdown: frames up once, centroid peaks / top-k / crops / instance network / peaks on the device, one
        dense record back (include/sleap_b200.h)."""
        cc, fp = self.centroid_crop, self.instance_peaks
        imgs = cc._prep(imgs)
        B, H, W, C = imgs.shape
        mc, mi = cc.keras_model, fp.keras_model
        K = int(cc.max_instances) if cc.max_instances else int(cc.max_peaks_per_sample)
        cap = max(B, self._td_key[0]) if self._td_key else B
        key = (cap, H, W, C, K, cc.peak_threshold, cc.refinement, cc.integral_patch_size, cc.input_scale, cc.max_peaks_per_sample,
               cc.max_instances, cc.crop_size, fp.peak_threshold, fp.refinement, fp.integral_patch_size, fp.input_scale,
               fp.max_crops_per_call)
        if self._td_key != key or mc.configured_for != (cap, H, W, C) or mi.configured_for != (fp.max_crops_per_call, cc.crop_size, cc.crop_size, C):
            p = TopdownParams(
                mc.model_id, mi.model_id,
                CentroidParams(cc.confmaps_buffer, -1 if cc.offsets_buffer is None else cc.offsets_buffer, int(cc.output_stride),
                               float(cc.peak_threshold), REFINE.get(cc.refinement, 0), int(cc.integral_patch_size), float(cc.input_scale),
                               int(cc.max_peaks_per_sample)),
                GlobalParams(fp.confmaps_buffer, -1 if fp.offsets_buffer is None else fp.offsets_buffer, int(fp.output_stride),
                             float(fp.peak_threshold), REFINE.get(fp.refinement, 0), int(fp.integral_patch_size), float(fp.input_scale)),
                int(cc.crop_size), int(cc.max_instances or 0), K, int(fp.max_crops_per_call))
            mc.handle.call("sb_topdown_configure", byref(p), cap, H, W, C)
            mc.configured_for = (cap, H, W, C)
            mi.configured_for = (fp.max_crops_per_call, cc.crop_size, cc.crop_size, C)
            cc._cfg_key = fp._cfg_key = None
            self._td_key = key
        n_nodes = next(h["channels"] for h in mi.spec["heads"] if h["name"] == fp.HEAD)
        ce = np.zeros((B, K, 2), np.float32); cv = np.zeros((B, K), np.float32)
        ip = np.zeros((B, K, n_nodes, 2), np.float32); iv = np.zeros((B, K, n_nodes), np.float32)
        nv = np.zeros((B,), np.int32); fl = np.zeros((B,), np.int32)
        mc.handle.call("sb_infer_topdown", mc.model_id, ptr(imgs), int(imgs.dtype == np.uint8), B, ptr(ce), ptr(cv), ptr(ip), ptr(iv),
                       ptr(nv), ptr(fl))
        n = int(nv.max()) if B else 0
        return {"centroids": ce[:, :n].copy(), "centroid_vals": cv[:, :n].copy(), "instance_peaks": ip[:, :n].copy(),
                "instance_peak_vals": iv[:, :n].copy(), "n_valid": nv.astype(np.int64), "flags": fl}

    def call(self, example):
        if isinstance(example, np.ndarray):
            example = dict(image=example)
        if self._can_fuse():
            return self._call_fused(_images_of(example))
        crop_out = self.centroid_crop.call(example)
        if isinstance(self.instance_peaks, FindInstancePeaksGroundTruth):                 # :2300-2304
            peaks_out = self.instance_peaks.call(example, crop_out)
        else:
            peaks_out = self.instance_peaks.call(crop_out)
        if isinstance(self.instance_peaks, FindInstancePeaksGroundTruth):
            n_nodes = max([p.shape[1] for p in peaks_out["instance_peaks"] if p.ndim == 3] + [0])
        else:
            n_nodes = next(h["channels"] for h in self.instance_peaks.keras_model.spec["heads"] if h["name"] == self.instance_peaks.HEAD)
        ip, nv = _ragged_to_dense(peaks_out["instance_peaks"], (n_nodes, 2))
        iv, _ = _ragged_to_dense(peaks_out["instance_peak_vals"], (n_nodes,))
        ce, _ = _ragged_to_dense(peaks_out["centroids"], (2,))
        cv, _ = _ragged_to_dense(peaks_out["centroid_vals"], ())
        out = {"centroids": ce, "centroid_vals": cv, "instance_peaks": ip, "instance_peak_vals": iv, "n_valid": nv}
        if "flags" in crop_out:
            out["flags"] = crop_out["flags"]
        return out


# ------------------------------------------------------------------------------------------
class BottomUpInferenceLayer(InferenceLayer):
    """sleap/nn/inference.py:2737-3003: net -> local peaks -> PAF scoring -> matching -> grouping,
    executed as one device pipeline (``sb_infer_bottomup``)."""

    def __init__(self, keras_model, paf_scorer, input_scale=1.0, pad_to_stride=1, cm_output_stride=None,
                 paf_output_stride=None, peak_threshold=0.2, refinement="local", integral_patch_size=5,
                 return_confmaps=False, return_pafs=False, return_paf_graph=False, confmaps_ind=None,
                 pafs_ind=None, offsets_ind=None, max_peaks_per_sample=1024, max_node_peaks=32,
                 max_instances=64, **kwargs):
        super().__init__(keras_model, input_scale=input_scale, pad_to_stride=pad_to_stride, **kwargs)
        self.paf_scorer = paf_scorer
        self.confmaps_buffer = _find_head(keras_model, "MultiInstanceConfmapsHead")
        self.pafs_buffer = _find_head(keras_model, "PartAffinityFieldsHead")
        self.offsets_buffer = _find_head(keras_model, "OffsetRefinementHead")
        if self.confmaps_buffer is None:
            raise ValueError("Index of the confidence maps output tensor must be specified if not named 'MultiInstanceConfmapsHead'.")
        if self.pafs_buffer is None:
            raise ValueError("Index of the part affinity fields output tensor must be specified if not named 'PartAffinityFieldsHead'.")
        self.cm_output_stride = cm_output_stride or keras_model.cm.head_strides["MultiInstanceConfmapsHead"]
        self.paf_output_stride = paf_output_stride or keras_model.cm.head_strides["PartAffinityFieldsHead"]
        self.peak_threshold = peak_threshold
        self.refinement = refinement
        self.integral_patch_size = integral_patch_size
        self.return_confmaps = return_confmaps
        self.return_pafs = return_pafs
        self.return_paf_graph = return_paf_graph
        self.max_peaks_per_sample = max_peaks_per_sample
        self.max_node_peaks = max_node_peaks
        self.max_instances = max_instances
        self._cfg_key = None
        self._keep = None

    def params(self) -> BottomUpParams:
        ps = self.paf_scorer
        edges = i32(ps.edge_inds).reshape(-1, 2)
        sorted_e = i32(list(ps.sorted_edge_inds))
        mip = ps.min_instance_peaks
        if isinstance(mip, float):
            mip = int(mip * ps.n_nodes)
        self._keep = (edges, sorted_e)
        return BottomUpParams(
            self.confmaps_buffer, self.pafs_buffer, -1 if self.offsets_buffer is None else self.offsets_buffer,
            int(self.cm_output_stride), int(self.paf_output_stride), float(self.peak_threshold),
            REFINE.get(self.refinement, 0), int(self.integral_patch_size), ps.n_nodes, ps.n_edges,
            edges.ctypes.data, sorted_e.ctypes.data, len(sorted_e), int(ps.n_points), float(ps.max_edge_length_ratio),
            float(ps.dist_penalty_weight), float(ps.min_line_scores), int(mip), float(self.input_scale),
            int(self.max_peaks_per_sample), int(self.max_node_peaks), int(self.max_instances))

    def _configure(self, B, H, W, C):
        m = self.keras_model
        if not (m.configured_for and m.configured_for[0] >= B and m.configured_for[1:] == (H, W, C)):
            m.configure(B, H, W, C)
            self._cfg_key = None
        ps = self.paf_scorer
        key = (m.configured_for, self.peak_threshold, self.refinement, self.integral_patch_size, self.input_scale,
               ps.n_points, ps.max_edge_length_ratio, ps.dist_penalty_weight, ps.min_line_scores, ps.min_instance_peaks,
               self.max_peaks_per_sample, self.max_node_peaks, self.max_instances)
        if self._cfg_key != key:
            p = self.params()
            m.handle.call("sb_bottomup_configure", m.model_id, byref(p))
            self._cfg_key = key

    def call(self, data):
        imgs = self._prep(_images_of(data))
        if imgs.dtype != np.uint8:
            raise ValueError("BottomUpInferenceLayer expects uint8 frames (the fused path reads raw frames).")
        B, H, W, C = imgs.shape
        self._configure(B, H, W, C)
        m = self.keras_model
        I, N = self.max_instances, self.paf_scorer.n_nodes
        ip = np.zeros((B, I, N, 2), np.float32)
        iv = np.zeros((B, I, N), np.float32)
        isc = np.zeros((B, I), np.float32)
        nv = np.zeros((B,), np.int32)
        fl = np.zeros((B,), np.int32)
        m.handle.call("sb_infer_bottomup", m.model_id, ptr(imgs), B, ptr(ip), ptr(iv), ptr(isc), ptr(nv), ptr(fl))
        n = int(nv.max()) if B else 0
        out = {"instance_peaks": ip[:, :n].copy(), "instance_peak_vals": iv[:, :n].copy(),
               "instance_scores": isc[:, :n].copy(), "n_valid": nv.astype(np.int64), "flags": fl}
        pg = getattr(m, "peer_gather", None)
        if pg is not None:          # multi-GPU: this call was one exchange step; the window came over with the result copy
            out["gathered_records"], out["gathered_counts"] = pg.gathered(-1, B, I, N)
        if self.return_confmaps or self.return_pafs:
            cms, pafs = m.forward(imgs, ["MultiInstanceConfmapsHead", "PartAffinityFieldsHead"])
            if self.return_confmaps:
                out["confmaps"] = cms
            if self.return_pafs:
                out["part_affinity_fields"] = pafs
        if self.return_paf_graph:
            out.update(self.fetch_graph(B))
        return out

    def fetch_graph(self, B):
        m = self.keras_model
        cp = B * self.max_peaks_per_sample
        cc = B * self.paf_scorer.n_edges * self.max_node_peaks ** 2
        peaks = np.zeros((cp, 2), np.float32); pv = np.zeros((cp,), np.float32); pc = np.zeros((cp,), np.int32)
        po = np.zeros((B + 1,), np.int32)
        ei = np.zeros((cc,), np.int32); epi = np.zeros((cc, 2), np.int32); ls = np.zeros((cc,), np.float32)
        co = np.zeros((B + 1,), np.int32)
        m.handle.call("sb_bottomup_fetch_graph", m.model_id, B, cp, ptr(peaks), ptr(pv), ptr(pc), ptr(po), cc, ptr(ei),
                      ptr(epi), ptr(ls), ptr(co))
        rows = lambda a, o: [a[o[b]:o[b + 1]].copy() for b in range(B)]
        return {"peaks": rows(peaks, po), "peak_vals": rows(pv, po), "peak_channel_inds": rows(pc, po),
                "edge_inds": rows(ei, co), "edge_peak_inds": rows(epi, co), "line_scores": rows(ls, co)}


def bottomup_from_maps(cms, pafs, paf_scorer, cm_output_stride, peak_threshold=0.2, refinement="integral",
                       integral_patch_size=5, offsets=None, input_scale=1.0, max_peaks_per_sample=1024,
                       max_node_peaks=32, max_instances=64, return_paf_graph=True, handle=None):
    """The BottomUpInferenceLayer post-processing chain (inference.py:2892-3003) on caller-supplied
    confidence maps / PAFs -- the parity entry point (identical maps in, bit-exact instances out)."""
    h = handle or _lib.default_handle()
    cms, pafs = f32(cms), f32(pafs)
    B, H, W, C = cms.shape
    _, Hp, Wp, C2 = pafs.shape
    ps = paf_scorer
    edges = i32(ps.edge_inds).reshape(-1, 2)
    sorted_e = i32(list(ps.sorted_edge_inds))
    mip = ps.min_instance_peaks
    if isinstance(mip, float):
        mip = int(mip * ps.n_nodes)
    p = BottomUpParams(-1, -1, -1, int(cm_output_stride), int(ps.pafs_stride), float(peak_threshold),
                       REFINE.get(refinement, 0), int(integral_patch_size), ps.n_nodes, ps.n_edges, edges.ctypes.data,
                       sorted_e.ctypes.data, len(sorted_e), int(ps.n_points), float(ps.max_edge_length_ratio),
                       float(ps.dist_penalty_weight), float(ps.min_line_scores), int(mip), float(input_scale),
                       int(max_peaks_per_sample), int(max_node_peaks), int(max_instances))
    I, N = max_instances, ps.n_nodes
    ip = np.zeros((B, I, N, 2), np.float32); iv = np.zeros((B, I, N), np.float32); isc = np.zeros((B, I), np.float32)
    nv = np.zeros((B,), np.int32); fl = np.zeros((B,), np.int32)
    cp = B * max_peaks_per_sample
    cc = B * ps.n_edges * max_node_peaks ** 2
    peaks = np.zeros((cp, 2), np.float32); pv = np.zeros((cp,), np.float32); pc = np.zeros((cp,), np.int32)
    po = np.zeros((B + 1,), np.int32)
    ei = np.zeros((cc,), np.int32); epi = np.zeros((cc, 2), np.int32); ls = np.zeros((cc,), np.float32)
    co = np.zeros((B + 1,), np.int32)
    off = None if offsets is None else f32(offsets)
    h.call("sb_bottomup_from_maps", byref(p), ptr(cms), B, H, W, ptr(pafs), Hp, Wp, ptr(off), ptr(ip), ptr(iv), ptr(isc),
           ptr(nv), ptr(fl), cp, ptr(peaks) if return_paf_graph else None, ptr(pv), ptr(pc), ptr(po), cc, ptr(ei),
           ptr(epi), ptr(ls), ptr(co))
    out = {"instance_peaks": [ip[b, :nv[b]].copy() for b in range(B)],
           "instance_peak_vals": [iv[b, :nv[b]].copy() for b in range(B)],
           "instance_scores": [isc[b, :nv[b]].copy() for b in range(B)], "n_valid": nv, "flags": fl}
    if return_paf_graph:
        rows = lambda a, o: [a[o[b]:o[b + 1]].copy() for b in range(B)]
        out.update({"peaks": rows(peaks, po), "peak_vals": rows(pv, po), "peak_channel_inds": rows(pc, po),
                    "edge_inds": rows(ei, co), "edge_peak_inds": rows(epi, co), "line_scores": rows(ls, co)})
    return out


class BottomUpInferenceModel(InferenceModel):
    """sleap/nn/inference.py:3006-3052."""

    def __init__(self, bottomup_layer: BottomUpInferenceLayer):
        self.bottomup_layer = bottomup_layer

    def call(self, example):
        return self.bottomup_layer.call(example)

    def predict_batches(self, data, batch_size: int = 4):
        """Generator over per-batch result dicts for a stack of uint8 frames (the Predictor batch loop,
        inference.py:377-420), double-buffered: the upload of batch i+1 overlaps the compute of batch i."""
        layer = self.bottomup_layer
        imgs = _images_of(data)
        n = len(imgs)
        if n == 0:
            return
        first = layer._prep(np.asarray(imgs[0:min(n, batch_size)]))
        if first.dtype != np.uint8 or layer.return_confmaps or layer.return_pafs or layer.return_paf_graph:
            for i in range(0, n, batch_size):        # generic (synchronous) path
                yield self.predict_on_batch(np.asarray(imgs[i:i + batch_size]))
            return
        _, H, W, C = first.shape
        layer._configure(batch_size, H, W, C)
        m = layer.keras_model
        I, N = layer.max_instances, layer.paf_scorer.n_nodes
        starts = list(range(0, n, batch_size))
        keep = {}

        def submit(k):
            batch = layer._prep(np.asarray(imgs[starts[k]:starts[k] + batch_size]))
            keep[k % 2] = batch                       # the async copy reads this host buffer until collect()
            m.handle.call("sb_bottomup_submit", m.model_id, ptr(batch), batch.shape[0], k % 2)
            return batch.shape[0]

        sizes = {0: submit(0)}
        for k in range(len(starts)):
            if k + 1 < len(starts):
                sizes[k + 1] = submit(k + 1)
            B = sizes.pop(k)
            ip = np.zeros((B, I, N, 2), np.float32); iv = np.zeros((B, I, N), np.float32)
            isc = np.zeros((B, I), np.float32); nv = np.zeros((B,), np.int32); fl = np.zeros((B,), np.int32)
            m.handle.call("sb_bottomup_collect", m.model_id, k % 2, B, ptr(ip), ptr(iv), ptr(isc), ptr(nv), ptr(fl))
            w = int(nv.max()) if B else 0
            out = {"instance_peaks": ip[:, :w], "instance_peak_vals": iv[:, :w], "instance_scores": isc[:, :w],
                   "n_valid": nv.astype(np.int64), "flags": fl}
            pg = getattr(m, "peer_gather", None)
            if pg is not None:      # multi-GPU: every rank's records of this step came over with the result copy (sb_gather_*)
                out["gathered_records"], out["gathered_counts"] = pg.gathered(k % 2, B, I, N)
            yield out

    def predict(self, data, numpy: bool = True, batch_size: int = 4, **kwargs):
        """sleap/nn/inference.py:989-1045 with the pipelined batch loop."""
        return _merge_batches(list(self.predict_batches(data, batch_size)))


# ------------------------------------------------------------------------------------------
class BottomUpMultiClassInferenceLayer(InferenceLayer):
    """sleap/nn/inference.py:3351-3589: network (confidence maps + class maps) -> local peaks -> identity grouping by the
    class-map probability at each peak (sleap_b200.nn.identity.classify_peaks_from_maps).  The network and the peak
    finder run on the device; the grouping is host logic in the reference too (TensorFlow ops around a SciPy callback)."""

    def __init__(self, keras_model, input_scale=1.0, pad_to_stride=1, cm_output_stride=None, class_maps_output_stride=None,
                 peak_threshold=0.2, refinement="integral", integral_patch_size=5, return_confmaps=False,
                 return_class_maps=False, **kwargs):
        super().__init__(keras_model, input_scale=input_scale, pad_to_stride=pad_to_stride, **kwargs)
        heads = keras_model.cm.head_buffers
        if "MultiInstanceConfmapsHead" not in heads:
            raise ValueError("Index of the confidence maps output tensor must be specified if not named 'MultiInstanceConfmapsHead'.")
        if "ClassMapsHead" not in heads:
            raise ValueError("Index of the class maps output tensor must be specified if not named 'ClassMapsHead'.")
        self.has_offsets = "OffsetRefinementHead" in heads
        self.cm_output_stride = cm_output_stride or keras_model.cm.head_strides["MultiInstanceConfmapsHead"]
        self.class_maps_output_stride = class_maps_output_stride or keras_model.cm.head_strides["ClassMapsHead"]
        self.peak_threshold = peak_threshold
        self.refinement = refinement
        self.integral_patch_size = integral_patch_size
        self.return_confmaps = return_confmaps
        self.return_class_maps = return_class_maps

    def forward_pass(self, data):
        """:3462-3490 (the class maps come out of a linear 1x1 head; their sigmoid, heads.py:336-338, is applied here)."""
        imgs = self._prep(_images_of(data))
        names = ["MultiInstanceConfmapsHead", "ClassMapsHead"] + (["OffsetRefinementHead"] if self.has_offsets else [])
        outs = self.keras_model.forward(imgs, names)
        cms, class_maps = outs[0], outs[1]
        with np.errstate(over="ignore"):         # exp(+large) = inf -> 1 / inf = 0: the limit the sigmoid has there
            class_maps = (np.float32(1) / (np.float32(1) + np.exp(-class_maps, dtype=np.float32))).astype(np.float32)
        return cms, class_maps, (outs[2] if self.has_offsets else None)

    def find_peaks(self, cms, offsets):
        """:3492-3528."""
        h = self.keras_model.handle
        if offsets is None:
            peaks, vals, s_inds, c_inds = peak_finding.find_local_peaks(cms, threshold=self.peak_threshold, refinement=self.refinement,
                                                                        integral_patch_size=self.integral_patch_size, handle=h)
        else:
            peaks, vals, s_inds, c_inds = peak_finding.find_local_peaks_with_offsets(cms, offsets, threshold=self.peak_threshold, handle=h)
        return (peaks * np.float32(self.cm_output_stride)).astype(np.float32), vals, s_inds, c_inds

    def call(self, data):
        """:3530-3589."""
        from sleap_b200.nn import identity
        cms, class_maps, offsets = self.forward_pass(data)
        peaks, peak_vals, s_inds, c_inds = self.find_peaks(cms, offsets)
        peaks = (peaks / np.float32(self.class_maps_output_stride)).astype(np.float32)
        inst, inst_vals, inst_scores = identity.classify_peaks_from_maps(class_maps, peaks, peak_vals, s_inds, c_inds,
                                                                         n_channels=cms.shape[3])
        inst = (inst * np.float32(self.class_maps_output_stride)).astype(np.float32)
        if self.input_scale != 1.0:
            inst = (inst / np.float32(self.input_scale) + np.float32(0.5)).astype(np.float32)
        out = {"instance_peaks": inst, "instance_peak_vals": inst_vals, "instance_scores": inst_scores}
        if self.return_confmaps:
            out["confmaps"] = cms
        if self.return_class_maps:
            out["class_maps"] = class_maps
        return out


class BottomUpMultiClassInferenceModel(InferenceModel):
    """sleap/nn/inference.py:3592-3635."""

    def __init__(self, inference_layer: BottomUpMultiClassInferenceLayer):
        self.inference_layer = inference_layer

    def call(self, example):
        return self.inference_layer.call(example)


class TopDownMultiClassFindPeaks(InferenceLayer):
    """sleap/nn/inference.py:3863-4136: centered-instance confidence maps + class vectors on crops -> global peaks ->
    one instance per class and sample (``identity.classify_peaks_from_vectors``).  Confidence maps and the class-vector
    head's feature map come from one device pass; the head's dense layers and the grouping run on the host."""

    HEAD = "CenteredInstanceConfmapsHead"

    def __init__(self, keras_model, input_scale=1.0, output_stride=None, peak_threshold=0.2, refinement="local",
                 integral_patch_size=5, return_confmaps=False, return_class_vectors=False, optimal_grouping=True,
                 max_crops_per_call=64, **kwargs):
        super().__init__(keras_model, input_scale=input_scale, pad_to_stride=1, **kwargs)
        if self.HEAD not in keras_model.cm.head_buffers:
            raise ValueError(f"Index of the confidence maps output tensor must be specified if not named '{self.HEAD}'.")
        if "ClassVectorsHead" not in keras_model.cm.vector_taps:
            raise ValueError("Index of the classifier output tensor must be specified if not named 'ClassVectorsHead'.")
        self.has_offsets = "OffsetRefinementHead" in keras_model.cm.head_buffers
        self.output_stride = output_stride or keras_model.cm.head_strides[self.HEAD]
        self.peak_threshold = peak_threshold
        self.refinement = refinement
        self.integral_patch_size = integral_patch_size
        self.return_confmaps = return_confmaps
        self.return_class_vectors = return_class_vectors
        self.optimal_grouping = optimal_grouping
        self.max_crops_per_call = max_crops_per_call

    def call(self, inputs):
        from sleap_b200.nn import identity
        if isinstance(inputs, dict):
            crops = inputs["crops"]
        else:
            crops, inputs = inputs, {}
        crops = self._prep(crops)
        n = crops.shape[0]
        if "crop_sample_inds" in inputs:
            samples, sinds = int(inputs["samples"]), np.asarray(inputs["crop_sample_inds"], np.int32)
        else:
            samples, sinds = n, np.arange(n, dtype=np.int32)
        m = self.keras_model
        n_nodes = next(h["channels"] for h in m.spec["heads"] if h["name"] == self.HEAD)
        n_classes = next(h["channels"] for h in m.spec["heads"] if h["name"] == "ClassVectorsHead")
        pts = np.zeros((n, n_nodes, 2), np.float32)
        vals = np.zeros((n, n_nodes), np.float32)
        probs = np.zeros((n, n_classes), np.float32)
        names = [self.HEAD, "ClassVectorsHead"] + (["OffsetRefinementHead"] if self.has_offsets else [])
        all_cms = []
        for i in range(0, n, self.max_crops_per_call):
            sl = slice(i, min(n, i + self.max_crops_per_call))
            outs = m.forward(np.ascontiguousarray(crops[sl]), names)
            cms, probs[sl] = outs[0], outs[1]
            if self.has_offsets:
                pts[sl], vals[sl] = peak_finding.find_global_peaks_with_offsets(cms, outs[2], threshold=self.peak_threshold, handle=m.handle)
            else:
                pts[sl], vals[sl] = peak_finding.find_global_peaks(cms, threshold=self.peak_threshold, refinement=self.refinement,
                                                                   integral_patch_size=self.integral_patch_size, handle=m.handle)
            if self.return_confmaps:
                all_cms.append(cms)
        pts = (pts * np.float32(self.output_stride)).astype(np.float32)                      # :4076-4082
        if self.input_scale != 1.0:
            pts = (pts / np.float32(self.input_scale) + np.float32(0.5)).astype(np.float32)
        if inputs.get("crop_offsets") is not None:
            pts = (pts + f32(inputs["crop_offsets"]).reshape(n, 1, 2)).astype(np.float32)     # :4084-4088
        if self.optimal_grouping:
            points, point_vals, class_probs = identity.classify_peaks_from_vectors(pts, vals, probs, sinds, samples)
            out = {"instance_peaks": points, "instance_peak_vals": point_vals, "instance_scores": class_probs}
        else:
            out = {"instance_peaks": pts, "instance_peak_vals": vals, "instance_scores": probs}
        for k in ("centroids", "centroid_vals"):
            if k in inputs:
                out[k] = inputs[k]
        if self.return_confmaps:
            cms = np.concatenate(all_cms) if all_cms else np.zeros((0,), np.float32)
            out["instance_confmaps"] = [cms[sinds == s_] for s_ in range(samples)]
        if self.return_class_vectors and self.optimal_grouping:
            out["class_vectors"] = probs
        return out


class TopDownMultiClassInferenceModel(InferenceModel):
    """sleap/nn/inference.py:4139-4210: centroid stage (model or ground truth) -> TopDownMultiClassFindPeaks."""

    def __init__(self, centroid_crop, instance_peaks: TopDownMultiClassFindPeaks):
        self.centroid_crop = centroid_crop
        self.instance_peaks = instance_peaks

    def call(self, example):
        if isinstance(example, np.ndarray):
            example = dict(image=example)
        crop_out = self.centroid_crop.call(example)
        out = self.instance_peaks.call(crop_out)
        res = {k: out[k] for k in ("instance_peaks", "instance_peak_vals", "instance_scores")}
        if "centroids" in out:
            res["centroids"], _ = _ragged_to_dense(out["centroids"], (2,))
            res["centroid_vals"], _ = _ragged_to_dense(out["centroid_vals"], ())
        return res


class PredictedInstance:
    """Array contract of ``sleap.PredictedInstance.from_numpy`` (sleap/instance.py:1164)."""

    def __init__(self, points, point_confidences, instance_score, skeleton=None, track=None, tracking_score=0.0):
        self.points = np.asarray(points)
        self.point_confidences = np.asarray(point_confidences)
        self.score = float(instance_score)
        self.skeleton = skeleton
        self.track = track
        self.tracking_score = float(tracking_score)

    @classmethod
    def from_numpy(cls, points, point_confidences, instance_score, skeleton=None, track=None, tracking_score=0.0):
        return cls(points, point_confidences, instance_score, skeleton, track, tracking_score)

    def numpy(self):
        return self.points

    @property
    def n_visible_points(self):
        return int(np.sum(~np.isnan(self.points).any(axis=1)))


class LabeledFrame:
    def __init__(self, video, frame_idx, instances):
        self.video, self.frame_idx, self.instances = video, frame_idx, instances


class Predictor:
    """sleap/nn/inference.py:158-590."""

    verbosity = "none"
    report_rate = 2.0
    model_paths: List[str] = []
    tracker = None          # optional sleap_b200.nn.tracking.Tracker applied frame by frame (:3306-3313)

    def __init__(self, batch_size=4):
        self.batch_size = batch_size
        self.inference_model = None

    @classmethod
    def from_model_paths(cls, model_paths, peak_threshold=0.2, integral_refinement=True, integral_patch_size=5,
                         batch_size=4, resize_input_layer=True, max_instances=None, precision=PRECISION_FP16,
                         handle=None, **caps):
        """:176-311: dispatch on the head type found in each model's training_config.json."""
        if isinstance(model_paths, str):
            model_paths = [model_paths]
        if not model_paths:
            raise ValueError("Must specify at least one model path.")   # :2479
        cfgs = {}
        for p in model_paths:
            cfg, d = cls._read_config(p)
            heads = {k: v for k, v in cfg["model"]["heads"].items() if v is not None}
            cfgs[next(iter(heads))] = (cfg, d)
        kw = dict(peak_threshold=peak_threshold, integral_refinement=integral_refinement,
                  integral_patch_size=integral_patch_size, batch_size=batch_size, precision=precision, handle=handle)
        kw.update(caps)
        if "single_instance" in cfgs:
            return SingleInstancePredictor.from_trained_models(cfgs["single_instance"], **kw)
        if "multi_class_topdown" in cfgs:
            mk = {k: kw[k] for k in ("peak_threshold", "integral_refinement", "integral_patch_size", "batch_size", "precision", "handle")}
            return TopDownMultiClassPredictor.from_trained_models(centroid_model_path=cfgs.get("centroid"),
                                                                  confmap_model_path=cfgs["multi_class_topdown"], **mk)
        if "centroid" in cfgs or "centered_instance" in cfgs:
            return TopDownPredictor.from_trained_models(centroid_model_path=cfgs.get("centroid"),
                                                        confmap_model_path=cfgs.get("centered_instance"),
                                                        max_instances=max_instances, **kw)
        if "multi_instance" in cfgs:
            return BottomUpPredictor.from_trained_models(cfgs["multi_instance"], max_instances=max_instances, **kw)
        if "multi_class_bottomup" in cfgs:
            mk = {k: kw[k] for k in ("peak_threshold", "integral_refinement", "integral_patch_size", "batch_size", "precision", "handle")}
            return BottomUpMultiClassPredictor.from_trained_models(cfgs["multi_class_bottomup"], **mk)
        raise ValueError("Could not create predictor from model paths:" + "\n".join(model_paths))

    # -- shared helpers ---------------------------------------------------------------------
    @staticmethod
    def _read_config(model_path):
        """``TrainingJobConfig.load_json`` (config/training_job.py:93-124) on a model folder or a json inside it."""
        cfg_path = model_path if model_path.endswith(".json") else os.path.join(model_path, "training_config.json")
        with open(cfg_path) as f:
            return json.load(f), os.path.dirname(cfg_path)

    @staticmethod
    def _load(cfg_and_dir, precision, handle, resize_in_graph=True):
        """``resize_in_graph=False``: the network graph gets no resize op (top-down instance models: their crops come
        from frames that were already resized, FindInstancePeaks(resize_input_image=False), :2405-2413); the
        configured ``input_scaling`` is still returned through ``model.input_scale``."""
        if isinstance(cfg_and_dir, (str, os.PathLike)):        # the reference passes model paths here
            cfg_and_dir = Predictor._read_config(os.fspath(cfg_and_dir))
        cfg, d = cfg_and_dir
        spec = arch.spec_from_config(cfg["model"], *_skeleton_from_cfg(cfg))
        pre = cfg["data"]["preprocessing"]
        weights = load_weights(d)
        # The reference reads the channel count off the loaded Keras model's input (:905-911,
        # ``is_grayscale``); here it is the C_in of the first convolution's kernel.
        first = arch.compile_model(spec, 1).layers[0]["name"]
        in_ch = int(np.asarray(weights[first]["kernel"]).shape[2])
        scale = float(pre.get("input_scaling", 1.0) or 1.0)
        model = DeviceModel(spec, weights, input_channels=in_ch, input_scale=scale if resize_in_graph else 1.0,
                            pad_to_stride=pre.get("pad_to_stride"), precision=precision, handle=handle)
        model.config_input_scale = scale
        return cfg, spec, model

    def _as_frames(self, data):
        """``predict`` accepts a video path, ``Video`` or ``VideoReader`` provider (:496-531, make_pipeline :329-371):
        those are read through the threaded ``FrameFeeder`` (decode ahead into pinned batch buffers)."""
        from sleap_b200.io.video import FrameFeeder, Video, VideoReader
        if isinstance(data, (str, os.PathLike, Video, VideoReader)):
            return FrameFeeder(data, batch_size=self.batch_size)
        return data

    def _label_examples(self, reader):
        """Batches of labels examples (make_pipeline with a LabelsReader, :329-371): stacked frames plus the
        per-sample ground-truth ``instances`` / ``centroids`` the stand-in layers read."""
        idx = reader.indices()
        for i in range(0, len(idx), self.batch_size):
            exs = [reader.example(j) for j in idx[i:i + self.batch_size]]
            batch = {"image": np.stack([e["image"] for e in exs]), "instances": [e["instances"] for e in exs],
                     "frame_ind": np.asarray([e["frame_ind"] for e in exs]), "video_ind": np.asarray([e["video_ind"] for e in exs])}
            if "centroids" in exs[0]:
                batch["centroids"] = [e["centroids"] for e in exs]
            yield batch

    def _batches(self, data):
        from sleap_b200.io.video import FrameFeeder
        imgs = _images_of(data)
        if isinstance(imgs, FrameFeeder):
            i = 0
            for _, batch in imgs.batches():
                yield i, batch
                i += len(batch)
            return
        n = len(imgs)
        for i in range(0, n, self.batch_size):
            batch = np.stack([np.asarray(imgs[j]) for j in range(i, min(n, i + self.batch_size))])
            yield i, batch

    def _predict_generator(self, data):
        """:377-420: one predict_on_batch per batch (+ frame indices)."""
        from sleap_b200.io.labels import Labels, LabelsReader
        from sleap_b200.io.video import FrameFeeder
        if isinstance(data, Labels):
            data = LabelsReader(data, with_centroids=True, center_on_part=getattr(self, "anchor_part", None))
        if isinstance(data, LabelsReader):
            for batch in self._label_examples(data):
                use_gt = getattr(self, "uses_ground_truth", False)
                ex = self.inference_model.predict_on_batch(batch if use_gt else batch["image"])
                ex["frame_ind"], ex["video_ind"] = batch["frame_ind"], batch["video_ind"]
                ex["image_hw"] = tuple(batch["image"].shape[1:3])
                if self.tracker is not None and getattr(self.tracker, "uses_image", False):
                    ex["image"] = batch["image"]
                self._check_flags(ex)
                yield ex
            return
        data = self._as_frames(data)
        feeder = data if isinstance(data, FrameFeeder) else None
        frame_inds = (lambda a, b: np.asarray(feeder.inds[a:b])) if feeder is not None else (lambda a, b: np.arange(a, b))
        try:
            if hasattr(self.inference_model, "predict_batches"):      # pipelined device loop (upload i+1 || compute i)
                i0 = 0
                want_img = self.tracker is not None and getattr(self.tracker, "uses_image", False)   # flow trackers (:2664-2671)
                imgs_all = _images_of(data)
                hw = tuple(np.asarray(imgs_all[0]).shape[:2]) if len(imgs_all) else (1, 1)
                for ex in self.inference_model.predict_batches(imgs_all, self.batch_size):
                    n = len(ex["n_valid"])
                    ex["frame_ind"] = frame_inds(i0, i0 + n)
                    ex["video_ind"] = np.zeros(n, np.int64)
                    ex["image_hw"] = hw
                    if want_img:
                        ex["image"] = np.stack([np.asarray(imgs_all[j]) for j in range(i0, i0 + n)])
                    i0 += n
                    self._check_flags(ex)
                    yield ex
                return
            for i0, batch in self._batches(data):
                ex = self.inference_model.predict_on_batch(batch)
                ex["frame_ind"] = frame_inds(i0, i0 + len(batch))
                ex["video_ind"] = np.zeros(len(batch), np.int64)
                ex["image_hw"] = tuple(batch.shape[1:3])
                if self.tracker is not None and getattr(self.tracker, "uses_image", False):
                    ex["image"] = batch
                self._check_flags(ex)
                yield ex
        finally:
            if feeder is not None:
                feeder.close()

    # Device workspaces are capacity bounded (the reference's ragged tensors are not): a frame that hits a cap comes
    # back truncated with SB_FLAG_* bits set.  ``on_overflow``: "warn" (default), "raise" or "ignore".
    on_overflow = "warn"
    _FLAG_NAMES = {1: "max_peaks_per_sample", 2: "max_node_peaks", 4: "max_instances_per_frame"}   # SB_FLAG_* (include/sleap_b200.h)

    def _check_flags(self, ex):
        fl = ex.get("flags")
        if fl is None or self.on_overflow == "ignore":
            return
        fl = np.asarray(fl)
        if not fl.any():
            return
        bits = int(np.bitwise_or.reduce(fl.astype(np.int64)))
        names = [n for b, n in self._FLAG_NAMES.items() if bits & b] or [f"flags=0x{bits:x}"]
        frames = np.asarray(ex.get("frame_ind", np.arange(len(fl))))[fl != 0].tolist()
        msg = (f"device capacity reached ({', '.join(names)}) on frame(s) {frames[:8]}{'...' if len(frames) > 8 else ''}: "
               "detections were truncated; raise the corresponding cap (max_peaks_per_sample / max_node_peaks / "
               "max_instances_per_frame) when constructing the predictor")
        if self.on_overflow == "raise":
            raise OverflowError(msg)
        import warnings
        warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def _with_progress(self, gen, n_total):
        """:422-491: ``verbosity`` = "none" | "json" (one JSON line per ``report_period`` seconds: n_processed, n_total, elapsed,
        rate over the last 30 batches, eta) | "rich" (a rich progress bar with the same rate estimate)."""
        import time
        from collections import deque
        if self.verbosity not in ("json", "rich"):
            yield from gen
            return
        n_processed, n_recent, el_recent = 0, deque(maxlen=30), deque(maxlen=30)
        t0_all = t0_batch = last_report = time.time()
        period = 1.0 / self.report_rate
        progress = task = None
        if self.verbosity == "rich":
            import rich.progress
            progress = rich.progress.Progress("{task.description}", rich.progress.BarColumn(), "[progress.percentage]{task.percentage:>3.0f}%",
                                              "ETA:", rich.progress.TimeRemainingColumn(), auto_refresh=False, speed_estimate_period=5)
            progress.start()
            task = progress.add_task("Predicting...", total=n_total)
        try:
            for ex in gen:
                now = time.time()
                n_batch = len(ex["frame_ind"])
                n_processed += n_batch
                n_recent.append(n_batch); el_recent.append(now - t0_batch)
                t0_batch = now
                rate = sum(n_recent) / max(sum(el_recent), 1e-9)
                if progress is not None:
                    progress.update(task, advance=n_batch)
                if now - last_report > period:
                    if progress is not None:
                        progress.refresh()
                    else:
                        print(json.dumps({"n_processed": n_processed, "n_total": n_total, "elapsed": now - t0_all, "rate": rate,
                                          "eta": (n_total - n_processed) / rate if n_total else None}), flush=True)
                    last_report = now
                yield ex
        finally:
            if progress is not None:
                progress.refresh()
                progress.stop()

    @staticmethod
    def _n_total(data):
        try:
            return len(_images_of(data))
        except TypeError:
            return len(getattr(data, "inds", [])) or None

    def predict(self, data, make_labels: bool = True):
        """:496-531."""
        gen = self._with_progress(self._predict_generator(data), self._n_total(data))
        if make_labels:
            return self._make_labeled_frames_from_generator(gen, data)
        return list(gen)

    def skeleton(self):
        """Skeleton of the loaded model(s): node names (+ edges for bottom-up models), as the reference takes them
        from the training config (:1547-1560, :2562-2580, :3230-3240)."""
        from sleap_b200.io.labels import Skeleton
        for m in (getattr(self, "bottomup_model", None), getattr(self, "confmap_model", None), getattr(self, "centroid_model", None),
                  getattr(self, "model", None)):
            if m is not None and m.spec.get("part_names"):
                return Skeleton(m.spec["part_names"], m.spec.get("edges") or [])
        raise ValueError("the loaded model carries no part names")

    def to_labels(self, frames, video_filename: str = "", video_spec=None):
        """``predict(..., make_labels=True)`` output -> ``sleap_b200.io.labels.Labels`` (``.save("out.slp")`` writes the
        reference's HDF5 labels container, sleap/io/format/hdf5.py:265-575)."""
        from sleap_b200.io.labels import labels_from_predictions
        return labels_from_predictions(frames, self.skeleton(), video_spec, video_filename)

    def _make_labeled_frames_from_generator(self, generator, data):
        """:3230-3343 pattern: a consumer thread builds the objects while the batch loop runs."""
        q: "queue.Queue" = queue.Queue()
        frames: List[LabeledFrame] = []

        def worker():
            while True:
                ex = q.get()
                if ex is None:
                    return
                new = self._frames_from_example(ex)
                if self.tracker is not None:                     # sequential by nature; runs on the consumer thread
                    hw = tuple(ex.get("image_hw") or (1, 1))
                    for k, lf in enumerate(new):
                        img = ex["image"][k] if "image" in ex else None
                        lf.instances = self.tracker.track(lf.instances, img_hw=hw, img=img, t=lf.frame_idx)
                frames.extend(new)

        t = threading.Thread(target=worker)
        t.start()
        try:
            for ex in generator:
                q.put(ex)
        finally:
            q.put(None)
            t.join()
        if self.tracker is not None:                                 # :2702-2703, :3345-3346
            self.tracker.final_pass(frames)
        return frames

    def _frames_from_example(self, ex):
        out = []
        scores = ex.get("instance_scores")
        topdown = scores is None and "centroid_vals" in ex
        if topdown:                       # top-down: the instance score is the centroid confidence (:2640-2660)
            scores = ex["centroid_vals"]
        for i in range(len(ex["instance_peaks"])):
            insts = []
            for j in range(ex["instance_peaks"].shape[1]):
                pts = ex["instance_peaks"][i, j]
                if np.all(np.isnan(pts)):
                    continue   # :3285
                sc = float(scores[i, j]) if scores is not None else float(np.nansum(ex["instance_peak_vals"][i, j]))
                insts.append(PredictedInstance.from_numpy(pts, ex["instance_peak_vals"][i, j], sc))
            mi = None if topdown else getattr(self, "max_instances", None)   # top-down caps centroids (:1879-1894) only
            if mi is not None and len(insts) > mi:   # :3297
                insts = sorted(insts, key=lambda x: x.score, reverse=True)[:mi]
            out.append(LabeledFrame(int(ex["video_ind"][i]), int(ex["frame_ind"][i]), insts))
        return out


def _skeleton_from_cfg(cfg):
    sk = (cfg.get("data", {}).get("labels", {}).get("skeletons") or [None])[0]
    if not sk or "nodes" not in sk:
        return None, None
    try:
        names = [n["id"]["py/state"]["py/tuple"][0] if "py/state" in n["id"] else None for n in sk["nodes"]]
        if all(names):
            return names, None
    except Exception:
        pass
    return None, None


class SingleInstancePredictor(Predictor):
    """sleap/nn/inference.py:1418-1636."""

    def __init__(self, confmap_model, peak_threshold=0.2, integral_refinement=True, integral_patch_size=5,
                 batch_size=4):
        super().__init__(batch_size)
        self.confmap_model = confmap_model
        self.peak_threshold = peak_threshold
        self.integral_refinement = integral_refinement
        self.integral_patch_size = integral_patch_size
        self._initialize_inference_model()

    def _initialize_inference_model(self):
        """:1456-1478."""
        m = self.confmap_model
        self.inference_model = SingleInstanceInferenceModel(SingleInstanceInferenceLayer(
            keras_model=m, input_scale=m.input_scale, pad_to_stride=m.cm.max_stride,
            peak_threshold=self.peak_threshold, refinement="integral" if self.integral_refinement else "local",
            integral_patch_size=self.integral_patch_size))

    @classmethod
    def from_trained_models(cls, model_path, inference_object=None, peak_threshold=0.2, integral_refinement=True,
                            integral_patch_size=5, batch_size=4, resize_input_layer=True, precision=PRECISION_FP16,
                            handle=None, **_):
        """:1480-1545 (same argument names; ``model_path`` may also be a loaded (config, folder) pair)."""
        _, _, model = cls._load(model_path, precision, handle)
        return cls(model, peak_threshold, integral_refinement, integral_patch_size, batch_size)


class TopDownPredictor(Predictor):
    """sleap/nn/inference.py:2314-2735."""

    def __init__(self, centroid_model=None, confmap_model=None, crop_size=160, peak_threshold=0.2,
                 integral_refinement=True, integral_patch_size=5, batch_size=4, max_instances=None,
                 max_peaks_per_sample=256):
        super().__init__(batch_size)
        self.max_peaks_per_sample = max_peaks_per_sample
        if centroid_model is None and confmap_model is None:
            raise ValueError("Either the centroid or topdown confidence map model must be provided.")   # :2479
        self.centroid_model, self.confmap_model = centroid_model, confmap_model
        self.anchor_part = None          # instance_cropping.center_on_part of the model config (ground-truth centroids)
        self.crop_size = crop_size
        self.peak_threshold = peak_threshold
        self.integral_refinement = integral_refinement
        self.integral_patch_size = integral_patch_size
        self.max_instances = max_instances
        self._initialize_inference_model()

    def _initialize_inference_model(self):
        """:2373-2433."""
        ref = "integral" if self.integral_refinement else "local"
        cm, im = self.centroid_model, self.confmap_model
        if cm is None:                                   # ground-truth centroids stand in for the centroid model
            cc = CentroidCropGroundTruth(crop_size=self.crop_size, handle=im.handle)
        else:
            cc = CentroidCrop(keras_model=cm, crop_size=self.crop_size if im is not None else 1, input_scale=cm.input_scale,
                              pad_to_stride=cm.cm.max_stride, peak_threshold=self.peak_threshold, refinement=ref,
                              integral_patch_size=self.integral_patch_size, max_instances=self.max_instances,
                              return_crops=im is not None, max_peaks_per_sample=self.max_peaks_per_sample)
        if im is None:                                   # ground-truth instances stand in for the instance model
            fp = FindInstancePeaksGroundTruth()
        else:
            iscale = float(getattr(im, "config_input_scale", im.input_scale))
            if im.input_scale != 1.0:
                raise ValueError("top-down instance models must be built without a resize op (resize_input_image=False)")
            fp = FindInstancePeaks(keras_model=im, input_scale=iscale, peak_threshold=self.peak_threshold,
                                   refinement=ref, integral_patch_size=self.integral_patch_size)
            if cm is None:
                cc.input_scale = iscale                  # :2414-2415
            else:
                cc.precrop_resize = iscale               # :2416-2419
        self.inference_model = TopDownInferenceModel(cc, fp)

    @property
    def uses_ground_truth(self):
        return self.centroid_model is None or self.confmap_model is None

    @classmethod
    def from_trained_models(cls, centroid_model_path=None, confmap_model_path=None, batch_size=4, peak_threshold=0.2,
                            integral_refinement=True, integral_patch_size=5, resize_input_layer=True,
                            max_instances=None, precision=PRECISION_FP16, handle=None, max_peaks_per_sample=256, **_):
        """:2435-2560 (same argument names; paths may also be loaded (config, folder) pairs)."""
        centroid_cfg, confmap_cfg = centroid_model_path, confmap_model_path
        if centroid_cfg is None and confmap_cfg is None:
            raise ValueError("Either the centroid or topdown confidence map model must be provided.")  # :2479
        cmodel = imodel = None
        crop, anchor = 1, None
        if centroid_cfg is not None:
            ccfg, _, cmodel = cls._load(centroid_cfg, precision, handle)
            anchor = ccfg["data"]["instance_cropping"].get("center_on_part")
        if confmap_cfg is not None:
            icfg, _, imodel = cls._load(confmap_cfg, precision, handle, resize_in_graph=False)
            crop = icfg["data"]["instance_cropping"]["crop_size"]
            anchor = icfg["data"]["instance_cropping"].get("center_on_part") if anchor is None else anchor
        obj = cls(cmodel, imodel, crop, peak_threshold, integral_refinement, integral_patch_size, batch_size, max_instances,
                  max_peaks_per_sample)
        obj.anchor_part = anchor
        return obj


class BottomUpPredictor(Predictor):
    """sleap/nn/inference.py:3055-3349 (attrs :3104-3117)."""

    def __init__(self, bottomup_model, part_names, edges, peak_threshold=0.2, batch_size=4, max_edge_length_ratio=0.25,
                 dist_penalty_weight=1.0, paf_line_points=10, min_line_scores=0.25, integral_refinement=True,
                 integral_patch_size=5, max_instances=None, max_peaks_per_sample=1024, max_node_peaks=32,
                 max_instances_per_frame=64):
        super().__init__(batch_size)
        self.bottomup_model = bottomup_model
        self.part_names, self.edges = list(part_names), [tuple(e) for e in edges]
        self.peak_threshold = peak_threshold
        self.max_edge_length_ratio = max_edge_length_ratio
        self.dist_penalty_weight = dist_penalty_weight
        self.paf_line_points = paf_line_points
        self.min_line_scores = min_line_scores
        self.integral_refinement = integral_refinement
        self.integral_patch_size = integral_patch_size
        self.max_instances = max_instances
        self._caps = (max_peaks_per_sample, max_node_peaks, max_instances_per_frame)
        self._initialize_inference_model()

    def _initialize_inference_model(self):
        """:3119-3150."""
        m = self.bottomup_model
        scorer = paf_grouping.PAFScorer(
            part_names=self.part_names, edges=self.edges, pafs_stride=m.cm.head_strides["PartAffinityFieldsHead"],
            max_edge_length_ratio=self.max_edge_length_ratio, dist_penalty_weight=self.dist_penalty_weight,
            n_points=self.paf_line_points, min_instance_peaks=0, min_line_scores=self.min_line_scores)
        self.inference_model = BottomUpInferenceModel(BottomUpInferenceLayer(
            keras_model=m, paf_scorer=scorer, input_scale=m.input_scale, pad_to_stride=m.cm.max_stride,
            peak_threshold=self.peak_threshold, refinement="integral" if self.integral_refinement else "local",
            integral_patch_size=self.integral_patch_size, max_peaks_per_sample=self._caps[0],
            max_node_peaks=self._caps[1], max_instances=self._caps[2]))

    @classmethod
    def from_trained_models(cls, model_path, batch_size=4, peak_threshold=0.2, integral_refinement=True,
                            integral_patch_size=5, max_edge_length_ratio=0.25, dist_penalty_weight=1.0,
                            paf_line_points=10, min_line_scores=0.25, resize_input_layer=True, max_instances=None,
                            precision=PRECISION_FP16, handle=None, max_peaks_per_sample=1024, max_node_peaks=32,
                            max_instances_per_frame=64, **_):
        """:3150-3228 (same argument names; ``model_path`` may also be an already loaded (config, folder) pair).
        ``max_peaks_per_sample`` / ``max_node_peaks`` / ``max_instances_per_frame`` size the device workspaces (the
        reference's ragged tensors are unbounded); a frame that reaches one is reported through ``on_overflow``."""
        _, spec, model = cls._load(model_path, precision, handle)
        return cls(model, spec["part_names"], spec["edges"], peak_threshold=peak_threshold, batch_size=batch_size,
                   max_edge_length_ratio=max_edge_length_ratio, dist_penalty_weight=dist_penalty_weight,
                   paf_line_points=paf_line_points, min_line_scores=min_line_scores,
                   integral_refinement=integral_refinement, integral_patch_size=integral_patch_size,
                   max_instances=max_instances, max_peaks_per_sample=max_peaks_per_sample, max_node_peaks=max_node_peaks,
                   max_instances_per_frame=max_instances_per_frame)


class BottomUpMultiClassPredictor(Predictor):
    """sleap/nn/inference.py:3638-3860: bottom-up identity models (confidence maps + class maps).  Instances come out one per
    class, in class order; each gets the ``Track`` named after its class (:3781-3790), ``score`` = mean point confidence and
    ``tracking_score`` = mean class probability (:3818-3826)."""

    def __init__(self, model, classes, peak_threshold=0.2, batch_size=4, integral_refinement=True, integral_patch_size=5, tracks=None):
        super().__init__(batch_size)
        self.model = model
        self.classes = list(classes)
        self.peak_threshold = peak_threshold
        self.integral_refinement = integral_refinement
        self.integral_patch_size = integral_patch_size
        self.tracks = tracks
        self._initialize_inference_model()

    def _initialize_inference_model(self):
        """:3682-3697."""
        m = self.model
        self.inference_model = BottomUpMultiClassInferenceModel(BottomUpMultiClassInferenceLayer(
            keras_model=m, input_scale=m.input_scale, pad_to_stride=m.cm.max_stride, peak_threshold=self.peak_threshold,
            refinement="integral" if self.integral_refinement else "local", integral_patch_size=self.integral_patch_size,
            cm_output_stride=m.cm.head_strides["MultiInstanceConfmapsHead"],
            class_maps_output_stride=m.cm.head_strides["ClassMapsHead"]))

    @classmethod
    def from_trained_models(cls, model_path, batch_size=4, peak_threshold=0.2, integral_refinement=True, integral_patch_size=5,
                            resize_input_layer=True, precision=PRECISION_FP16, handle=None, **_):
        """:3699-3745."""
        _, spec, model = cls._load(model_path, precision, handle)
        return cls(model, spec["classes"], peak_threshold=peak_threshold, batch_size=batch_size,
                   integral_refinement=integral_refinement, integral_patch_size=integral_patch_size)

    def _frames_from_example(self, ex):
        """:3781-3838."""
        return _multiclass_frames(self, ex)


def _multiclass_frames(predictor, ex):
    """Shared by the two identity predictors (:3781-3838, :4506-4566): instance j of a frame belongs to class j and gets
    the ``Track`` named after it; ``score`` = mean point confidence, ``tracking_score`` = mean class probability."""
    from sleap_b200.nn.tracking import Track
    tracks = predictor.tracks
    if tracks is None:
        tracks = predictor.tracks = [Track(spawned_on=0, name=n) for n in predictor.classes]
    out = []
    for i in range(len(ex["instance_peaks"])):
        insts = []
        for j in range(ex["instance_peaks"].shape[1]):
            pts, confs = ex["instance_peaks"][i, j], ex["instance_peak_vals"][i, j]
            if np.all(np.isnan(pts)):
                continue
            insts.append(PredictedInstance.from_numpy(pts, confs, float(np.nanmean(confs)), track=tracks[j] if j < len(tracks) else None,
                                                      tracking_score=float(np.nanmean(ex["instance_scores"][i, j]))))
        out.append(LabeledFrame(int(ex["video_ind"][i]), int(ex["frame_ind"][i]), insts))
    return out


class TopDownMultiClassPredictor(Predictor):
    """sleap/nn/inference.py:4213-4605: top-down identity models -- a centroid model (or ground-truth centroids from a
    labels provider) and a centered-instance model with a class-vector head; one instance per class and frame."""

    def __init__(self, centroid_model=None, confmap_model=None, crop_size=160, peak_threshold=0.2, integral_refinement=True,
                 integral_patch_size=5, batch_size=4, max_instances=None, max_peaks_per_sample=256, tracks=None):
        super().__init__(batch_size)
        if confmap_model is None:
            raise ValueError("The topdown multi-class confidence map model must be provided.")
        self.centroid_model, self.confmap_model = centroid_model, confmap_model
        self.classes = list(confmap_model.spec["classes"])
        self.anchor_part = None
        self.crop_size = crop_size
        self.peak_threshold = peak_threshold
        self.integral_refinement = integral_refinement
        self.integral_patch_size = integral_patch_size
        self.max_instances = max_instances
        self.max_peaks_per_sample = max_peaks_per_sample
        self.tracks = tracks
        self._initialize_inference_model()

    def _initialize_inference_model(self):
        """:4263-4318."""
        ref = "integral" if self.integral_refinement else "local"
        cm, im = self.centroid_model, self.confmap_model
        iscale = float(getattr(im, "config_input_scale", im.input_scale))
        if im.input_scale != 1.0:
            raise ValueError("top-down instance models must be built without a resize op (resize_input_image=False)")
        if cm is None:
            cc = CentroidCropGroundTruth(crop_size=self.crop_size, handle=im.handle)
            cc.input_scale = iscale
        else:
            cc = CentroidCrop(keras_model=cm, crop_size=self.crop_size, input_scale=cm.input_scale, pad_to_stride=cm.cm.max_stride,
                              peak_threshold=self.peak_threshold, refinement=ref, integral_patch_size=self.integral_patch_size,
                              max_instances=self.max_instances, return_crops=True, max_peaks_per_sample=self.max_peaks_per_sample)
            cc.precrop_resize = iscale
        fp = TopDownMultiClassFindPeaks(keras_model=im, input_scale=iscale, peak_threshold=self.peak_threshold, refinement=ref,
                                        integral_patch_size=self.integral_patch_size)
        self.inference_model = TopDownMultiClassInferenceModel(cc, fp)

    @property
    def uses_ground_truth(self):
        return self.centroid_model is None

    @classmethod
    def from_trained_models(cls, centroid_model_path=None, confmap_model_path=None, batch_size=4, peak_threshold=0.2,
                            integral_refinement=True, integral_patch_size=5, resize_input_layer=True, max_instances=None,
                            precision=PRECISION_FP16, handle=None, max_peaks_per_sample=256, **_):
        """:4320-4401."""
        if confmap_model_path is None:
            raise ValueError("The topdown multi-class confidence map model must be provided.")
        cmodel, anchor = None, None
        if centroid_model_path is not None:
            ccfg, _, cmodel = cls._load(centroid_model_path, precision, handle)
            anchor = ccfg["data"]["instance_cropping"].get("center_on_part")
        icfg, _, imodel = cls._load(confmap_model_path, precision, handle, resize_in_graph=False)
        crop = icfg["data"]["instance_cropping"]["crop_size"]
        anchor = icfg["data"]["instance_cropping"].get("center_on_part") if anchor is None else anchor
        obj = cls(cmodel, imodel, crop, peak_threshold, integral_refinement, integral_patch_size, batch_size, max_instances,
                  max_peaks_per_sample)
        obj.anchor_part = anchor
        return obj

    def _frames_from_example(self, ex):
        """:4506-4566."""
        return _multiclass_frames(self, ex)


def load_model(model_path, batch_size=4, peak_threshold=0.2, refinement="integral", **kwargs):
    """sleap/nn/inference.py:4865-5005 (``sleap.load_model``)."""
    return Predictor.from_model_paths(model_path, peak_threshold=peak_threshold,
                                      integral_refinement=(refinement == "integral"), batch_size=batch_size, **kwargs)
