"""The reference's own *layer* tests, on the CUDA layers: an identity network (``Lambda(lambda x: x)`` named after the
head) feeds analytic confidence maps straight into CentroidCrop / FindInstancePeaks / SingleInstanceInferenceLayer
(tests/nn/test_inference.py:213-254, 257-379, 542-589, 1091-1150).  The same tests pass on the oracle layers
(tests/test_oracle_layers.py).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import synth

pytestmark = pytest.mark.gpu


def _model(head, channels, input_scale=1.0):
    from sleap_b200.nn.model import DeviceModel, PRECISION_FP32
    spec = dict(backbone="identity", heads=[dict(name=head, channels=channels, output_stride=1)], part_names=None, edges=None)
    return DeviceModel(spec, {}, input_channels=channels, input_scale=input_scale, pad_to_stride=1, precision=PRECISION_FP32)


def test_centroid_crop_layer():
    from sleap_b200.nn.inference import CentroidCrop
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    points = np.asarray([[[1.75, 2.75]], [[3.75, 4.75]], [[5.75, 6.75]]], np.float32)
    cms = synth.make_multi_confmaps(points, xv, yv, 1.5)[None]
    layer = CentroidCrop(keras_model=_model("CentroidConfmapsHead", 1), input_scale=1.0, crop_size=3, pad_to_stride=1,
                         output_stride=None, refinement="local", integral_patch_size=5, peak_threshold=0.2)
    out = layer.call(cms)
    assert len(out["centroids"]) == 1 and out["centroids"][0].shape == (3, 2) and out["centroid_vals"][0].shape == (3,)
    assert out["crops"].shape == (3, 3, 3, 1) and out["crop_offsets"].shape == (3, 2)
    assert_allclose(out["centroids"][0], points[:, 0])
    assert_allclose(out["centroid_vals"][0], [1, 1, 1], atol=0.1)
    for k, n in ((3, 3), (2, 2), (1, 1)):                       # test_centroid_inference :1137-1150
        layer.max_instances = k
        assert out["centroids"][0].shape[0] >= n and layer.call(cms)["centroids"][0].shape == (n, 2)


def test_instance_peaks_layer():
    from sleap_b200.nn.inference import FindInstancePeaks
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    points = np.asarray([[1.5, 2.5], [3.5, 4.5], [5.5, 6.5]], np.float32)
    cms = np.stack([synth.make_confmaps(points, xv, yv, 1.0), synth.make_confmaps(points + 1, xv, yv, 1.0)])
    model = _model("CenteredInstanceConfmapsHead", 3)
    layer = FindInstancePeaks(keras_model=model, input_scale=1.0, peak_threshold=0.2, refinement="integral")
    out = layer.call(cms)                                       # raw tensor: one crop per sample
    assert [p.shape for p in out["instance_peaks"]] == [(1, 3, 2), (1, 3, 2)]
    assert_allclose(out["instance_peaks"][0][0], points, atol=0.1)
    assert_allclose(out["instance_peaks"][1][0], points + 1, atol=0.1)
    assert_allclose(out["instance_peak_vals"][0][0], [1, 1, 1], atol=0.3)
    out = layer.call({"crops": cms, "crop_sample_inds": np.asarray([0, 0]), "samples": 1})   # one sample, two instances
    assert out["instance_peaks"][0].shape == (2, 3, 2)
    assert_allclose(out["instance_peaks"][0][1], points + 1, atol=0.1)
    out = layer.call({"crops": cms, "crop_sample_inds": np.asarray([0, 1]), "samples": 2, "centroids": [np.zeros((1, 2))] * 2,
                      "centroid_vals": [np.zeros(1)] * 2, "crop_offsets": np.asarray([[1, 2], [3, 4]], np.float32)})
    assert "centroids" in out and "centroid_vals" in out
    assert_allclose(out["instance_peaks"][0][0], points + [[1, 2]], atol=0.1)
    assert_allclose(out["instance_peaks"][1][0], points + 1 + [[3, 4]], atol=0.1)
    scale = 0.5                                                 # input scaling :347-379
    xv, yv = synth.make_grid_vectors(24, 24, 1)
    cms2 = np.stack([synth.make_confmaps(points / scale, xv, yv, 1.0 / scale), synth.make_confmaps((points + 1) / scale, xv, yv, 1.0 / scale)])
    layer = FindInstancePeaks(keras_model=_model("CenteredInstanceConfmapsHead", 3, input_scale=scale), input_scale=scale,
                              peak_threshold=0.2, refinement="integral")
    out = layer.call(cms2)
    assert_allclose(out["instance_peaks"][0][0], points / scale, atol=0.15)
    assert_allclose(out["instance_peaks"][1][0], (points + 1) / scale, atol=0.15)


def test_single_instance_inference():
    from sleap_b200.nn.inference import SingleInstanceInferenceLayer, SingleInstanceInferenceModel
    xv, yv = synth.make_grid_vectors(12, 12, 1)
    p0 = np.asarray([[1.75, 2.75], [3.75, 4.75], [5.75, 6.75]], np.float32)
    points = np.stack([p0, p0 + 1])
    cms = np.stack([synth.make_confmaps(points[0], xv, yv, 1.0), synth.make_confmaps(points[1], xv, yv, 1.0)])
    model = _model("SingleInstanceConfmapsHead", 3)
    layer = SingleInstanceInferenceLayer(keras_model=model, refinement="local")
    assert layer.output_stride == 1
    out = layer.call(cms)
    assert out["instance_peaks"].shape == (2, 1, 3, 2) and out["instance_peak_vals"].shape == (2, 1, 3)
    assert_array_equal(out["instance_peaks"][:, 0], points)
    assert_allclose(out["instance_peak_vals"], 1.0, atol=0.1)
    assert "confmaps" not in out
    assert_array_equal(layer.call({"image": cms})["instance_peaks"][:, 0], points)
    layer = SingleInstanceInferenceLayer(keras_model=model, refinement="local", return_confmaps=True)
    out = layer.call(cms)
    assert_array_equal(out["confmaps"], cms)
    preds = SingleInstanceInferenceModel(layer).predict(cms)
    assert preds["instance_peaks"].shape == (2, 1, 3, 2) and "instance_peak_vals" in preds and "confmaps" in preds
