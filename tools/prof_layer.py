"""Runs ONE conv layer of the tensor-core path (for ncu): python tools/prof_layer.py CIN COUT K H W B [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ctypes import c_int, c_void_p, byref
from sleap_b200 import _lib
from sleap_b200.nn import oplist as ol

cin, cout, k, H, W, B = [int(a) for a in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
rng = np.random.default_rng(0)
recs = [ol.buffer_record(0, 1, 1, 0, 1), ol.buffer_record(1, 1, cin, 0, 0), ol.buffer_record(2, 1, cout, 0, 0),
        ol.preprocess_record(0, 1, 1.0, 1)]
w0 = (rng.standard_normal((3, 3, 1, cin)) * 0.5).astype(np.float32); b0 = np.zeros(cin, np.float32)
w1 = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32); b1 = np.zeros(cout, np.float32)
blob = np.concatenate([w0.reshape(-1), b0, w1.reshape(-1), b1]).astype(np.float32)
o0, o1 = 0, w0.size + cin
recs.append(ol.conv_record(0, 0, 1, 1, 0, cin, 3, 1, True, o0, o0 + w0.size))
recs.append(ol.conv_record(1, 0, cin, 2, 0, cout, k, 1, True, o1, o1 + w1.size))
ops = np.ascontiguousarray(np.stack(recs).astype(np.int32))
h = _lib.default_handle()
mid = c_int(-1)
h.call("sb_load_model", _lib.ptr(ops), ops.shape[0], _lib.ptr(blob), int(blob.size), 0, byref(mid))
h.call("sb_model_configure", mid.value, B, H, W, 1)
imgs = rng.integers(0, 256, size=(B, H, W, 1), dtype=np.uint8)
out = np.zeros((B, H, W, cout), np.float32)
ids = np.asarray([2], np.int32)
ptrs = (c_void_p * 1)(out.ctypes.data)
h.call("sb_model_forward", mid.value, _lib.ptr(imgs), 1, B, 1, _lib.ptr(ids), ptrs)       # warm
import ctypes
try:
    rt = ctypes.CDLL("libcudart.so.12")
except OSError:
    rt = None
if rt: rt.cudaProfilerStart()             # `ncu --profile-from-start off` captures only the launches below
for _ in range(reps):
    h.call("sb_model_forward", mid.value, _lib.ptr(imgs), 1, B, 1, _lib.ptr(ids), ptrs)
if rt: rt.cudaProfilerStop()
print("ok", float(np.abs(out).mean()))
