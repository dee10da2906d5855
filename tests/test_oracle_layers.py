"""The oracle's inference *layers* against the reference's layer tests, which wrap an identity Keras model so that
analytic confidence maps go straight into the layer (tests/nn/test_inference.py:213-254 test_centroid_crop_layer,
:257-379 test_instance_peaks_layer, :542-589 test_single_instance_inference, :1091-1150 test_centroid_inference)."""
import numpy as np
from numpy.testing import assert_allclose, assert_array_equal

from oracle import inference as oinf
from oracle import synth


def _ident(head, channels, stride=1):
    return dict(backbone="identity", heads=[dict(name=head, channels=channels, output_stride=stride)])


def test_centroid_crop_layer():
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    points = np.asarray([[[1.75, 2.75]], [[3.75, 4.75]], [[5.75, 6.75]]], np.float32)
    cms = synth.make_multi_confmaps(points, xv, yv, 1.5)[None]
    out = oinf.centroid_crop_layer(cms, _ident("CentroidConfmapsHead", 1), None, crop_size=3, in_ch=1, refinement="local")
    assert out["centroids"].shape == (3, 2) and out["crops"].shape == (3, 3, 3, 1) and out["crop_offsets"].shape == (3, 2)
    assert_allclose(out["centroids"], points[:, 0])
    assert_allclose(out["centroid_vals"], [1, 1, 1], atol=0.1)
    assert_array_equal(out["crop_sample_inds"], [0, 0, 0])
    # test_centroid_inference: max_instances >= / < the number of peaks (:1137-1150)
    for k, n in ((3, 3), (2, 2), (1, 1)):
        o = oinf.centroid_crop_layer(cms, _ident("CentroidConfmapsHead", 1), None, crop_size=3, in_ch=1, refinement="local",
                                     max_instances=k)
        assert len(o["centroids"]) == n


def test_instance_peaks_layer():
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    points = np.asarray([[1.5, 2.5], [3.5, 4.5], [5.5, 6.5]], np.float32)
    cms = np.stack([synth.make_confmaps(points, xv, yv, 1.0), synth.make_confmaps(points + 1, xv, yv, 1.0)])
    spec = _ident("CenteredInstanceConfmapsHead", 3)
    pts, vals = oinf.find_instance_peaks_layer(cms, None, spec, None, in_ch=3, refinement="integral")
    assert pts.shape == (2, 3, 2) and vals.shape == (2, 3)
    assert_allclose(pts[0], points, atol=0.1)
    assert_allclose(pts[1], points + 1, atol=0.1)
    assert_allclose(vals, 1.0, atol=0.3)
    # offset adjustment (:318-345)
    off = np.asarray([[1, 2], [3, 4]], np.float32)
    pts, _ = oinf.find_instance_peaks_layer(cms, off, spec, None, in_ch=3, refinement="integral")
    assert_allclose(pts[0], points + [[1, 2]], atol=0.1)
    assert_allclose(pts[1], points + 1 + [[3, 4]], atol=0.1)
    # input scaling (:347-379): maps are made at 1/scale resolution, the layer resizes them by `scale`
    scale = 0.5
    xv, yv = synth.make_grid_vectors(int(12 / scale), int(12 / scale), 1)
    cms2 = np.stack([synth.make_confmaps(points / scale, xv, yv, 1.0 / scale),
                     synth.make_confmaps((points + 1) / scale, xv, yv, 1.0 / scale)])
    pts, _ = oinf.find_instance_peaks_layer(cms2, None, spec, None, in_ch=3, input_scale=scale, refinement="integral")
    assert_allclose(pts[0], points / scale, atol=0.15)
    assert_allclose(pts[1], (points + 1) / scale, atol=0.15)


def test_single_instance_inference():
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    p0 = np.asarray([[1.75, 2.75], [3.75, 4.75], [5.75, 6.75]], np.float32)
    points = np.stack([p0, p0 + 1])
    cms = np.stack([synth.make_confmaps(points[0], xv, yv, 1.0), synth.make_confmaps(points[1], xv, yv, 1.0)])
    out = oinf.single_instance_layer(cms, _ident("SingleInstanceConfmapsHead", 3), None, in_ch=3, refinement="local")
    assert out["instance_peaks"].shape == (2, 1, 3, 2) and out["instance_peak_vals"].shape == (2, 1, 3)
    assert_array_equal(out["instance_peaks"][:, 0], points)          # quarter-pixel local refinement is exact here
    assert_allclose(out["instance_peak_vals"], 1.0, atol=0.1)
    assert_array_equal(out["confmaps"], cms)
