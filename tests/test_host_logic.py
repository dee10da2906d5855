"""Host-side logic (no GPU): graph compiler structural known answers and edge ordering."""
import numpy as np
import pytest

import cases_paf
from oracle import paf_grouping as opg
from oracle import synth
from sleap_b200.nn import architectures as A
from sleap_b200.nn import paf_grouping as pg


def _unet_spec(cfg, heads):
    return dict(backbone="unet", backbone_cfg=cfg, head_type="x", heads=heads, part_names=None, edges=None)


def test_unet_param_counts():
    # reference tests/nn/architectures/test_unet.py:29-84 (34,512,128) and :86-119 (16,320)
    h = [dict(name="H", channels=1, output_stride=1)]
    cm = A.compile_model(_unet_spec(dict(filters=64, filters_rate=2, max_stride=16, output_stride=1, middle_block=True,
                                         up_interpolate=False), h), 1)
    assert A.count_params(cm) - (64 + 1) == 34512128
    cm = A.compile_model(_unet_spec(dict(filters=8, filters_rate=2, max_stride=4, output_stride=1, middle_block=False,
                                         up_interpolate=False), h), 1)
    assert A.count_params(cm) - (8 + 1) == 16320


def test_stacked_unet_param_count():
    # reference tests/nn/architectures/test_unet.py:121-157: 3 stacks, f16, 5 down / 5 up, interp -> 23,590,608
    h = [dict(name="H", channels=1, output_stride=1)]
    cm = A.compile_model(_unet_spec(dict(filters=16, filters_rate=2, max_stride=32, output_stride=1, middle_block=True,
                                         up_interpolate=True, stacks=3), h), 1)
    assert A.count_params(cm) - (16 + 1) == 23590608


def test_hourglass_param_count():
    # reference tests/nn/architectures/test_hourglass.py:31-48: 66,002,944 total
    spec = dict(backbone="hourglass", backbone_cfg=dict(), head_type="x", heads=[dict(name="H", channels=1, output_stride=4)],
                part_names=None, edges=None)
    cm = A.compile_model(spec, 1)
    assert A.count_params(cm) - (256 + 1) == 66002944


def test_c4_flops():
    # 92.32 GF (tconv) / 99.56 GF (interp) per 1024x1024 frame
    heads = [dict(name="MultiInstanceConfmapsHead", channels=13, output_stride=4),
             dict(name="PartAffinityFieldsHead", channels=24, output_stride=8)]
    c4 = dict(filters=16, filters_rate=2, max_stride=32, output_stride=4, middle_block=True, up_interpolate=False)
    assert abs(A.compile_model(_unet_spec(c4, heads), 1).flops_per_pixel * 1024 * 1024 / 1e9 - 92.32) < 0.01
    c4["up_interpolate"] = True
    assert abs(A.compile_model(_unet_spec(c4, heads), 1).flops_per_pixel * 1024 * 1024 / 1e9 - 99.56) < 0.01


def test_head_stride_error():
    heads = [dict(name="H", channels=1, output_stride=64)]
    with pytest.raises(ValueError):
        A.compile_model(_unet_spec(dict(filters=8, max_stride=16, output_stride=2), heads), 1)


def test_toposort_matches_networkx():
    cases_paf.check_toposort(pg)
    cases_paf.check_connection_candidates(pg)
    for edges in (synth.flies13_edge_inds(), cases_paf.TOPO_A, cases_paf.TOPO_B, [(0, 1), (1, 2), (2, 3)]):
        assert pg.toposort_edges(edges) == opg.toposort_edges(edges)


def test_spec_from_config_order():
    cfg = {"backbone": {"unet": dict(filters=16, filters_rate=2, max_stride=32, output_stride=4, middle_block=True,
                                     up_interpolate=False, stacks=1, stem_stride=None), "hourglass": None},
           "heads": {"single_instance": None, "multi_instance": {
               "confmaps": {"part_names": ["a", "b"], "sigma": 2.5, "output_stride": 4, "offset_refinement": True},
               "pafs": {"edges": [["a", "b"]], "sigma": 75, "output_stride": 8}}}}
    spec = A.spec_from_config(cfg)
    assert [h["name"] for h in spec["heads"]] == ["MultiInstanceConfmapsHead", "PartAffinityFieldsHead", "OffsetRefinementHead"]
    assert [h["channels"] for h in spec["heads"]] == [2, 2, 4]


def test_topdown_instance_score_is_centroid_confidence():
    """sleap/nn/inference.py:2640-2660: a top-down PredictedInstance's score is the centroid confidence (not the sum of
    its peak values), and no post-hoc max_instances cut is applied (the cap acts on centroids, :1879-1894)."""
    import numpy as np
    from sleap_b200.nn.inference import Predictor
    p = Predictor()
    p.max_instances = 1
    ex = {"instance_peaks": np.asarray([[[[1, 2], [3, 4]], [[5, 6], [7, 8]]]], np.float32),
          "instance_peak_vals": np.asarray([[[0.9, 0.8], [0.7, 0.6]]], np.float32),
          "centroid_vals": np.asarray([[0.25, 0.5]], np.float32), "video_ind": np.zeros(1, int), "frame_ind": np.zeros(1, int)}
    lf = p._frames_from_example(ex)[0]
    assert [i.score for i in lf.instances] == [0.25, 0.5]
    # bottom-up keeps its own scores and the max_instances cut (:3297)
    ex2 = dict(ex, instance_scores=np.asarray([[1.0, 2.0]], np.float32))
    lf2 = p._frames_from_example(ex2)[0]
    assert [i.score for i in lf2.instances] == [2.0]


def test_overflow_flags_are_reported():
    import numpy as np
    import pytest
    from sleap_b200.nn.inference import Predictor
    p = Predictor()
    ex = {"flags": np.asarray([0, 2, 0], np.int32), "frame_ind": np.asarray([10, 11, 12])}
    with pytest.warns(RuntimeWarning, match="max_node_peaks.*11"):
        p._check_flags(ex)
    p.on_overflow = "raise"
    with pytest.raises(OverflowError):
        p._check_flags(ex)
    p.on_overflow = "ignore"
    p._check_flags(ex)
    p.on_overflow = "raise"
    p._check_flags({"flags": np.zeros(3, np.int32)})


def test_progress_reporting_json_and_rich(capsys):
    """Predictor.verbosity (sleap/nn/inference.py:422-491): json lines carry n_processed / n_total / rate / eta."""
    import json
    import numpy as np
    from sleap_b200.nn.inference import Predictor
    p = Predictor()
    p.verbosity, p.report_rate = "json", 1e9
    out = list(p._with_progress(({"frame_ind": np.arange(i, i + 4)} for i in range(0, 12, 4)), 12))
    lines = [json.loads(l) for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 3 and lines[-1]["n_processed"] == 12 and lines[-1]["n_total"] == 12 and lines[-1]["eta"] == 0
    assert [l["n_processed"] for l in lines] == [4, 8, 12]
    p.verbosity = "rich"
    assert len(list(p._with_progress(({"frame_ind": np.arange(4)} for _ in range(3)), 12))) == 3
    p.verbosity = "none"
    assert len(list(p._with_progress(iter([{"frame_ind": np.arange(2)}]), 2))) == 1


def test_split_precision_layout_and_weight_expansion():
    """Precision 2 (compile_model(split=True)): every fp16 tensor takes 3C physical channels [lo | hi | hi], concat
    buffers interleave the triples of their parts, conv records carry physical in_C but logical out_C, and the expanded
    weight rows [Wh | Wl | Wh] reproduce x @ W to ~1e-6 from the split activations (numpy emulation of the three products)."""
    from oracle import synth
    from sleap_b200.nn import architectures as A, oplist as ol
    spec = dict(backbone="unet", head_type="multi_instance", part_names=synth.FLIES13_NODES, edges=synth.FLIES13_EDGES,
                backbone_cfg=dict(filters=16, filters_rate=2, max_stride=32, output_stride=4, middle_block=True, up_interpolate=False),
                heads=[dict(name="MultiInstanceConfmapsHead", channels=13, output_stride=4),
                       dict(name="PartAffinityFieldsHead", channels=24, output_stride=8)])
    cm1, cm2 = A.compile_model(spec, 1), A.compile_model(spec, 1, split=True)
    assert cm1.n_buffers == cm2.n_buffers and len(cm1.records) == len(cm2.records)
    r1, r2 = np.stack(cm1.records), np.stack(cm2.records)
    for a, b in zip(r1, r2):
        assert a[0] == b[0]
        if a[0] == ol.BUFFER:
            if a[1] == 0:
                assert b[3] == a[3] and b[4] == 1               # the preprocessed frame: same channels, fp32
            else:
                assert b[3] == (a[3] if a[4] else 3 * a[3])     # fp32 head buffers keep their size
        elif a[0] in (ol.CONV, ol.TCONV):
            first = a[1] == 0
            assert b[3] == (a[3] if first else 3 * a[3]) and b[8] == a[8]       # physical in_C, logical out_C
            assert b[2] == 3 * a[2] and b[7] == (a[7] if r2[int(b[6])][4] else 3 * a[7])
    w = A.make_synthetic_weights(cm1, 3)
    blob1, blob2 = cm1.pack_weights(w), cm2.pack_weights(w)
    rng = np.random.default_rng(0)
    # a decoder conv that reads a concat buffer (skip | upsampled): physical order [Xl Xh Xh | Ul Uh Uh]
    name = "stack0_dec1_s16_to_s8_refine_conv0"
    L1 = next(L for L in cm1.layers if L["name"] == name)
    L2 = next(L for L in cm2.layers if L["name"] == name)
    cin, cout, k = L1["cin"], L1["cout"], L1["k"]
    src, part = L2["expand"]
    assert len(src) == 3 * cin and sorted(src.tolist()) == sorted(list(range(cin)) * 3)
    half = cin // 2
    assert part[:half].tolist() == [0] * half and part[half:2 * half].tolist() == [1] * half and part[2 * half:3 * half].tolist() == [2] * half
    assert src[3 * half:4 * half].tolist() == list(range(half, cin))           # the second part's triple follows the first's
    W = blob1[cm1._w_slots[name]["w"]:][:k * k * cin * cout].reshape(k * k, cin, cout)
    We = blob2[cm2._w_slots[name]["w"]:][:k * k * 3 * cin * cout].reshape(k * k, 3 * cin, cout)
    assert np.array_equal(We.astype(np.float16).astype(np.float32), We)       # every expanded weight is exactly fp16
    x = np.maximum(rng.standard_normal((64, cin)), 0).astype(np.float32)
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    planes = np.stack([lo, hi, hi])                                             # plane p of logical channel c
    xe = planes[part, :, src].T                                                 # (64, 3*cin) in physical channel order
    got = xe.astype(np.float64) @ We[4].astype(np.float64)
    want = x.astype(np.float64) @ W[4].astype(np.float64)
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
