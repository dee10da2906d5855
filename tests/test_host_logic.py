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
