"""Secondary configurations of BASELINE.json (C1, C2, C3, C5) through the public predictors, one
JSON line each (frames/s end to end with host frames, CUDA-event timed device loop where available).
Not the driver's bench (that is bench.py = C4); results are copied into profiles/.

  python tools/bench_configs.py [c1] [c2] [c3] [c5] [--steps K] [--c5-batch B]   (C5 default: 16 frames per GPU and step)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleap_b200.nn import architectures as A
from sleap_b200.nn.inference import (BottomUpPredictor, SingleInstancePredictor, TopDownPredictor)
from sleap_b200.nn.model import DeviceModel

FLIES13 = ["head", "thorax", "abdomen", "wingL", "wingR", "forelegL", "forelegR", "midlegL", "midlegR", "hindlegL",
           "hindlegR", "eyeL", "eyeR"]


def frames(n, h, w, c, seed):
    """Page-locked frame stack (what FrameFeeder hands the predictors): the upload is a true asynchronous DMA."""
    import torch
    a = np.random.default_rng(seed).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    return torch.from_numpy(a).pin_memory().numpy()


def model_for(spec, in_ch, seed, input_scale=1.0):
    cm = A.compile_model(spec, in_ch, input_scale)
    w = A.make_synthetic_weights(cm, seed)
    return DeviceModel(spec, w, input_channels=in_ch, input_scale=input_scale, precision=0), cm, w


def timed(fn, steps, warmup=3):
    """Seconds per call (wall clock, the call returns host results) + nvidia-smi clock samples taken meanwhile."""
    global LAST_CLOCKS
    import bench
    for _ in range(warmup):
        fn()
    sampler = bench.ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    LAST_CLOCKS = sampler.stop()
    return dt


LAST_CLOCKS = None


def roofline(gflop_per_frame, fps):
    """Tensor roofline of a whole config: algorithmic conv FLOPs per second / measured cuBLAS bf16 burst peak."""
    import bench
    peaks, src = bench.peaks_file()
    tf = gflop_per_frame * fps / 1e3
    return {"bound": "tensor", "achieved": tf, "peak": float(peaks["bf16_tflops"]), "unit": "TFLOP/s",
            "frac": tf / float(peaks["bf16_tflops"]), "peak_source": f"{src} bf16_tflops", "note": "end-to-end time incl. H2D / post-processing / D2H"}


def unet(filters, max_stride, output_stride):
    return dict(filters=filters, filters_rate=2, max_stride=max_stride, output_stride=output_stride, middle_block=True,
                up_interpolate=True, stacks=1)


def single(name, size, nodes, B, steps):
    spec = dict(backbone="unet", backbone_cfg=unet(16, 16, 2), head_type="single_instance", part_names=nodes, edges=None,
                heads=[dict(name="SingleInstanceConfmapsHead", channels=len(nodes), output_stride=2)])
    m, cm, _ = model_for(spec, 1, 1001)
    pred = SingleInstancePredictor(m, peak_threshold=0.2, integral_refinement=True, batch_size=B)
    fr = frames(B, size, size, 1, 1)
    dt = timed(lambda: pred.inference_model.predict_on_batch(fr), steps)
    return {"config": name, "metric": "frames/s (predict_on_batch, host frames)", "value": B / dt, "ms_per_step": dt * 1e3,
            "batch": B, "gflop_per_frame": cm.flops_per_pixel * size * size / 1e9, "dtype": "f16", "clocks": LAST_CLOCKS,
            "roofline": roofline(cm.flops_per_pixel * size * size / 1e9, B / dt)}


def topdown(steps):
    cspec = dict(backbone="unet", backbone_cfg=unet(16, 16, 2), head_type="centroid", part_names=None, edges=None,
                 heads=[dict(name="CentroidConfmapsHead", channels=1, output_stride=2)])
    ispec = dict(backbone="unet", backbone_cfg=dict(unet(24, 16, 4), up_interpolate=False), head_type="centered_instance",
                 part_names=FLIES13, edges=None, heads=[dict(name="CenteredInstanceConfmapsHead", channels=13, output_stride=4)])
    cm_model, ccm, cw = model_for(cspec, 1, 1003, input_scale=0.5)
    B = 16
    fr = frames(B, 1024, 1024, 1, 3)
    # calibrate the centroid head so that ~5 animals per frame pass the threshold (random weights otherwise give thousands)
    cms = cm_model.forward(fr[:2])[0]
    thr = float(np.sort(cms.reshape(-1))[-(5 * 2 * 6)])
    im_model, icm, _ = model_for(ispec, 1, 1004)
    pred = TopDownPredictor(cm_model, im_model, crop_size=160, peak_threshold=thr, integral_refinement=True, batch_size=B,
                            max_instances=5)
    pred.inference_model.instance_peaks.peak_threshold = 0.0
    out = pred.inference_model.predict_on_batch(fr)
    dt = timed(lambda: pred.inference_model.predict_on_batch(fr), steps)
    return {"config": "C3 top-down centroid(512^2 after 0.5 scale)+centered-instance(160^2 crops), 1024x1024, max 5 animals, B=16",
            "metric": "frames/s (predict_on_batch, host frames)", "value": B / dt, "ms_per_step": dt * 1e3, "batch": B,
            "mean_instances_per_frame": float(np.mean(out.get("n_valid", [0]))), "dtype": "f16", "clocks": LAST_CLOCKS,
            "gflop_per_frame": ccm.flops_per_pixel * 512 * 512 / 1e9 + 5 * icm.flops_per_pixel * 160 * 160 / 1e9,
            "roofline": roofline(ccm.flops_per_pixel * 512 * 512 / 1e9 + 5 * icm.flops_per_pixel * 160 * 160 / 1e9, B / dt)}


def hourglass(steps, B=16):
    nodes = [f"n{i}" for i in range(24)]
    edges = [(f"n{i}", f"n{i + 1}") for i in range(23)]
    spec = dict(backbone="hourglass", backbone_cfg=dict(stem_stride=4, max_stride=64, output_stride=4, stem_filters=128,
                                                         filters=256, filter_increase=128, stacks=3),
                head_type="multi_instance", part_names=nodes, edges=edges,
                heads=[dict(name="MultiInstanceConfmapsHead", channels=24, output_stride=4),
                       dict(name="PartAffinityFieldsHead", channels=46, output_stride=4)])
    m, cm, w = model_for(spec, 3, 1005)
    fr = frames(B, 1536, 1536, 3, 5)
    cms, pafs = m.forward(fr[:1])
    thr = float(np.quantile(cms, 1 - 5.0 / (384 * 384)))   # ~5 peaks per channel
    pred = BottomUpPredictor(m, nodes, edges, peak_threshold=thr, batch_size=B, max_peaks_per_sample=4096,
                             max_node_peaks=64, max_instances_per_frame=64)
    out = pred.inference_model.predict_on_batch(fr)
    dt = timed(lambda: pred.inference_model.predict_on_batch(fr), steps, warmup=2)
    gf = cm.flops_per_pixel * 1536 * 1536 / 1e9
    return {"config": f"C5 stacked hourglass x3 bottom-up 1536x1536x3, 24 nodes / 23 edges, B={B} per GPU and step",
            "metric": "frames/s (predict_on_batch, host frames)", "value": B / dt, "ms_per_step": dt * 1e3, "batch": B,
            "gflop_per_frame": gf, "tflops": gf * B / dt / 1e3, "flags": [int(f) for f in out.get("flags", [])], "dtype": "f16",
            "clocks": LAST_CLOCKS, "roofline": roofline(gf, B / dt)}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c3", "c5"]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
    for c in which:
        if c == "c1":
            r = single("C1 single-instance UNet 256x256x1, 5 nodes, B=1", 256, list("abcde"), 1, steps)
        elif c == "c2":
            r = single("C2 single-instance UNet 512x512x1, 13 nodes, B=32", 512, FLIES13, 32, steps)
        elif c == "c3":
            r = topdown(steps)
        else:
            b5 = int(sys.argv[sys.argv.index("--c5-batch") + 1]) if "--c5-batch" in sys.argv else 16
            r = hourglass(max(3, steps // 3), b5)
        print(json.dumps(r), flush=True)
