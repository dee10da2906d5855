"""CUDA path on the reference's trained fixture models (tests/golden/, made from the reference's own test data):
the public ``Predictor.from_model_paths(...).predict(...)`` call must (1) satisfy the assertions of the
reference's predictor tests against the ground-truth labels (tests/nn/test_inference.py:585-800) and
(2) agree with the CPU oracle on the same frames (fp32 path: sub-pixel identical; fp16 tensor-core path: 0.15 px)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import reference_models as rm
from oracle import inference as oinf

pytestmark = pytest.mark.gpu

TOL = {1: 5e-3, 0: 0.15, 2: 5e-3}     # px, vs the fp32 oracle (2 = split fp16 pairs on the tensor cores: the fp32 bar)


def _matched(a, b, atol):
    i1, i2 = oinf.match_points(a, b)
    assert len(i1) == len(a)
    assert_allclose(a[i1], b[i2], atol=atol)


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_bottomup_trained_model(precision):
    from sleap_b200.nn.inference import BottomUpPredictor, Predictor
    imgs, gt = rm.frames("minimal_instance")
    pred = Predictor.from_model_paths([rm.model_dir("minimal_instance.bottomup")], precision=precision)
    assert isinstance(pred, BottomUpPredictor)
    frames = pred.predict(imgs)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    pts = np.concatenate([i.numpy() for i in frames[0].instances])
    _matched(gt[0].reshape(-1, 2), pts, 1.75)
    cfg, spec, w, in_ch = rm.load_fixture_model("minimal_instance.bottomup")
    want = oinf.bottomup_layer(imgs, spec, w, in_ch, 1.0, 8)
    out = pred.inference_model.predict_on_batch(imgs)
    assert int(out["n_valid"][0]) == len(want["instance_peaks"][0]) == 2
    assert_allclose(out["instance_peaks"][0, :2], want["instance_peaks"][0], atol=TOL[precision])
    assert_allclose(out["instance_scores"][0, :2], want["instance_scores"][0], atol={0: 2e-2, 1: 1e-4, 2: 5e-4}[precision])
    # the same frames through the provider path: Video -> threaded FrameFeeder -> pipelined submit/collect
    from sleap_b200.io.video import Video
    rep = np.concatenate([imgs] * 9)
    via_feeder = pred.predict(Video.from_numpy(rep))
    assert len(via_feeder) == 9 and [f.frame_idx for f in via_feeder] == list(range(9))
    for f in via_feeder:
        assert len(f.instances) == 2
        assert_allclose(np.concatenate([i.numpy() for i in f.instances]), pts, atol=1e-5)
    hi = BottomUpPredictor.from_trained_models(model_path=rm.model_dir("minimal_instance.bottomup"), min_line_scores=1.1,
                                               precision=precision)
    assert len(hi.predict(imgs)[0].instances) == 0


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_topdown_trained_models(precision):
    from sleap_b200.nn.inference import Predictor, TopDownPredictor
    imgs, gt = rm.frames("minimal_instance")
    paths = [rm.model_dir("minimal_instance.centroid"), rm.model_dir("minimal_instance.centered_instance")]
    pred = Predictor.from_model_paths(paths, precision=precision)
    assert isinstance(pred, TopDownPredictor) and pred.crop_size == 96
    out = pred.inference_model.predict_on_batch(imgs)
    assert int(out["n_valid"][0]) == 2
    ccfg, cspec, cw, cin = rm.load_fixture_model("minimal_instance.centroid")
    icfg, ispec, iw, iin = rm.load_fixture_model("minimal_instance.centered_instance")
    want = oinf.topdown_model(imgs, cspec, cw, ispec, iw, 96, cin, iin, 1.0, 1.0, 8, 8)
    assert_allclose(out["centroids"][0, :2], want["centroids"][0], atol=TOL[precision])
    # a centroid that moves by d px moves the bilinear crop, so the instance stage is compared more loosely in fp16
    assert_allclose(out["instance_peaks"][0, :2], want["instance_peaks"][0], atol=TOL[precision] * (3 if precision == 0 else 1))
    _matched(gt[0].reshape(-1, 2), out["instance_peaks"][0, :2].reshape(-1, 2), 2.0)
    for k in (1, 2, 3):                                      # test_topdown_predictor_centroid_max_instances
        p = Predictor.from_model_paths(paths, precision=precision, max_instances=k)
        assert int(p.inference_model.predict_on_batch(imgs)["n_valid"][0]) == min(k, 2)
    p = Predictor.from_model_paths(paths, precision=precision, peak_threshold=1.5)
    assert len(p.predict(imgs)[0].instances) == 0


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_single_instance_trained_model(precision):
    from sleap_b200.nn.inference import Predictor, SingleInstancePredictor
    imgs, gt = rm.frames("robot")
    d = rm.model_dir("minimal_robot.single_instance")
    pred = Predictor.from_model_paths([d], precision=precision)
    assert isinstance(pred, SingleInstancePredictor)
    frames = pred.predict(imgs)
    assert len(frames) == 2 and len(frames[0].instances) == 1
    pts = np.stack([f.instances[0].numpy() for f in frames])
    assert_allclose(pts, gt[:, 0], atol=10.0)
    cfg, spec, w, in_ch = rm.load_fixture_model("minimal_robot.single_instance")
    want = oinf.single_instance_layer(imgs, spec, w, in_ch, 0.5, 4)
    assert_allclose(pts, want["instance_peaks"][:, 0], atol=TOL[precision] * 2)   # x2: coordinates are /input_scale
    lo = Predictor.from_model_paths([d], precision=precision, peak_threshold=0.0).predict(imgs)
    assert all(np.isfinite(f.instances[0].numpy()).all() for f in lo)
    hi = Predictor.from_model_paths([d], precision=precision, peak_threshold=1.5).predict(imgs)
    assert all(len(f.instances) == 0 for f in hi)


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_topdown_single_model_modes(precision):
    """test_topdown_predictor_centroid (:638-656) and test_topdown_predictor_centered_instance (:728-757): a top-down
    predictor built from ONE model, the other stage replaced by its ground-truth stand-in layer, fed the labels."""
    from sleap_b200.nn.inference import TopDownPredictor
    labels = rm.labels_minimal_instance()
    gt = np.concatenate([i.numpy() for i in labels[0].instances])
    # centroid model only: FindInstancePeaksGroundTruth hands back the labelled instance nearest to each centroid
    pred = TopDownPredictor.from_trained_models(centroid_model_path=rm.model_dir("minimal_instance.centroid"), precision=precision)
    frames = pred.predict(labels)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    _matched(gt, np.concatenate([i.numpy() for i in frames[0].instances]), 1.5)
    for k in (1, 2, 3):                                                    # :659-671
        p = TopDownPredictor.from_trained_models(centroid_model_path=rm.model_dir("minimal_instance.centroid"), precision=precision,
                                                 max_instances=k)
        assert len(p.predict(labels)[0].instances) == min(k, 2)
    hi = TopDownPredictor.from_trained_models(centroid_model_path=rm.model_dir("minimal_instance.centroid"), precision=precision,
                                              peak_threshold=1.5)
    assert len(hi.predict(labels)[0].instances) == 0                       # :674-683
    # centered-instance model only: crops at the ground-truth centroids (CentroidCropGroundTruth)
    pred = TopDownPredictor.from_trained_models(confmap_model_path=rm.model_dir("minimal_instance.centered_instance"), precision=precision)
    assert pred.crop_size == 96
    frames = pred.predict(labels)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    _matched(gt, np.concatenate([i.numpy() for i in frames[0].instances]), 1.5)
    hi = TopDownPredictor.from_trained_models(confmap_model_path=rm.model_dir("minimal_instance.centered_instance"), precision=precision,
                                              peak_threshold=1.5)
    assert len(hi.predict(labels)[0].instances) == 0                       # :760-769
    with pytest.raises(ValueError):
        TopDownPredictor.from_trained_models()


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_topdown_centered_instance_with_scaling(precision):
    """test_topdown_predictor_centered_instance_with_scaling (:708-729) and
    test_topdown_predictor_centroid_centered_instance_with_scaling (:732-755): instance model trained at
    input_scaling 0.5 -- frames are resized before cropping, the crops are not resized again."""
    from sleap_b200.nn.inference import TopDownPredictor
    labels = rm.labels_minimal_instance()
    imgs, gt4 = rm.frames("minimal_instance")
    gt = gt4[0].reshape(-1, 2)
    d = rm.model_dir("minimal_instance.centered_instance_with_scaling")
    pred = TopDownPredictor.from_trained_models(confmap_model_path=d, precision=precision)
    assert pred.crop_size == 56
    frames = pred.predict(labels)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    pts = np.concatenate([i.numpy() for i in frames[0].instances])
    _matched(gt, pts, 1.5)
    # against the oracle restatement of the same variant
    cfg, spec, w, in_ch = rm.load_fixture_model("minimal_instance.centered_instance_with_scaling")
    cent_gt = np.stack([(g.min(0) + g.max(0)) * 0.5 for g in gt4[0]]).astype(np.float32)
    cc = oinf.centroid_crop_ground_truth_layer(imgs, [cent_gt], 56, input_scale=0.5)
    want, _ = oinf.find_instance_peaks_layer(cc["crops"], cc["crop_offsets"], spec, w, in_ch, input_scale=0.5, resize_input_image=False)
    i1, i2 = oinf.match_points(want.reshape(-1, 2), pts)
    assert_allclose(pts[i2], want.reshape(-1, 2)[i1], atol=TOL[precision] * 4)
    # full chain: centroid model + scaled instance model
    both = TopDownPredictor.from_trained_models(centroid_model_path=rm.model_dir("minimal_instance.centroid"), confmap_model_path=d,
                                                precision=precision)
    out = both.predict(imgs)
    assert len(out) == 1 and len(out[0].instances) == 2
    _matched(gt, np.concatenate([i.numpy() for i in out[0].instances]), 2.0)


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_bottomup_multiclass_trained_model(precision):
    """tests/nn/test_inference.py:809-852 through ``Predictor.from_model_paths`` -> ``BottomUpMultiClassPredictor``: the
    reference's trained identity model; two instances on the tracks named after their classes, points within 2 % of the
    ground truth; nothing above a 1.5 threshold.  Device maps / peaks, host-side identity grouping (as in the reference)."""
    import os
    from sleap_b200.nn.inference import BottomUpMultiClassPredictor, Predictor
    z = np.load(os.path.join(rm.GOLDEN, "frames_tracks_2node.npz"))
    imgs, gt, names = z["images"], z["points_gt"][0], [str(n) for n in z["track_names"][0]]
    pred = Predictor.from_model_paths([rm.model_dir("min_tracks_2node.bottomup_multiclass")], peak_threshold=0.7,
                                      integral_refinement=False, precision=precision)
    assert isinstance(pred, BottomUpMultiClassPredictor)
    out = pred.inference_model.predict_on_batch(imgs)
    assert out["instance_peaks"].shape == (1, 2, 2, 2) and out["instance_scores"].shape == (1, 2, 2)
    frames = pred.predict(imgs)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    got = sorted(frames[0].instances, key=lambda i: i.track.name)
    assert [i.track.name for i in got] == sorted(names)
    for inst in got:
        assert_allclose(inst.numpy(), gt[names.index(inst.track.name)], rtol=0.02)
        assert 0.9 < inst.tracking_score <= 1.0 and 0.7 < inst.score < 1.2
    hi = Predictor.from_model_paths([rm.model_dir("min_tracks_2node.bottomup_multiclass")], peak_threshold=1.5,
                                    integral_refinement=False, precision=precision).predict(imgs)
    assert len(hi) == 1 and len(hi[0].instances) == 0


@pytest.mark.parametrize("precision", [1, 0, 2])
def test_topdown_multiclass_trained_model(precision):
    """tests/nn/test_inference.py:855-894 through ``TopDownMultiClassPredictor.from_trained_models(confmap_model_path=...)``
    on a ``Labels`` input (ground-truth centroids, device crops): confidence maps and the class-vector head's feature map
    from one device pass, dense layers + grouping on the host; two instances on the tracks of their classes, points within
    2 % of the ground truth; nothing above a 1.5 threshold; class probabilities agree with the oracle network."""
    import os
    from oracle import convnet, preprocess as opre
    from sleap_b200.nn.inference import Predictor, TopDownMultiClassPredictor
    z = np.load(os.path.join(rm.GOLDEN, "frames_tracks_2node.npz"))
    gt, names = z["points_gt"][0], [str(n) for n in z["track_names"][0]]
    labels = rm.labels_tracks_2node()
    d = rm.model_dir("min_tracks_2node.topdown_multiclass")
    pred = TopDownMultiClassPredictor.from_trained_models(confmap_model_path=d, peak_threshold=0.7, integral_refinement=False,
                                                          precision=precision)
    assert isinstance(Predictor.from_model_paths([d], precision=precision), TopDownMultiClassPredictor)
    frames = pred.predict(labels)
    assert len(frames) == 1 and len(frames[0].instances) == 2
    got = sorted(frames[0].instances, key=lambda i: i.track.name)
    assert [i.track.name for i in got] == sorted(names)
    for inst in got:
        assert_allclose(inst.numpy(), gt[names.index(inst.track.name)], rtol=0.02)
        assert inst.tracking_score > 0.99
    # class probabilities of the device path (feature map tapped on the device, dense layers on the host) vs the oracle network
    cfg, spec, w, in_ch = rm.load_fixture_model("min_tracks_2node.topdown_multiclass")
    cc = oinf.centroid_crop_ground_truth_layer(z["images"], [gt[:, 1, :]], cfg["data"]["instance_cropping"]["crop_size"], 1.0)
    want = convnet.model_forward(opre.preprocess(cc["crops"], ensure_gray=True, input_scale=1.0, pad_stride=16), spec, w)[1]
    have = pred.confmap_model.forward(cc["crops"], ["ClassVectorsHead"])[0]
    assert_allclose(have, want, atol={0: 2e-2, 1: 1e-4, 2: 1e-4}[precision])
    hi = TopDownMultiClassPredictor.from_trained_models(confmap_model_path=d, peak_threshold=1.5, integral_refinement=False,
                                                        precision=precision).predict(labels)
    assert len(hi) == 1 and len(hi[0].instances) == 0
