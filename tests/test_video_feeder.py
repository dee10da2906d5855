"""Frame feeder: threaded decode into batch buffers must deliver exactly the frames
``MediaVideo.get_frame`` semantics give (sleap/io/video.py:486-507), in order, for any worker count."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from sleap_b200.io.video import FrameFeeder, Video, VideoReader


def _write_video(path, n=37, hw=(48, 64), color=False):
    w = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), 30, (hw[1], hw[0]), True)
    assert w.isOpened()
    rng = np.random.default_rng(0)
    for i in range(n):
        f = np.full(hw + (3,), 20 + 5 * i, np.uint8)
        f[5:15, i:i + 12] = 250
        if color:
            f[..., 2] = rng.integers(0, 255)
        w.write(f)
    w.release()
    return str(path)


def test_video_api_and_grayscale_detection(tmp_path):
    p = _write_video(tmp_path / "g.avi")
    v = Video.from_filename(p)
    assert v.grayscale is True and v.shape == (37, 48, 64, 1) and len(v) == 37
    f3 = v.get_frame(3)
    assert f3.shape == (48, 64, 1) and f3.dtype == np.uint8
    assert np.array_equal(v[3], f3) and v[2:5].shape == (3, 48, 64, 1)
    with pytest.raises(KeyError):
        v.get_frame(1000)                                     # "Unable to load frame" (video.py:495-496)
    with pytest.raises(FileNotFoundError):
        Video.from_filename(str(tmp_path / "missing.mp4"))
    c = Video.from_filename(_write_video(tmp_path / "c.avi", color=True))
    assert c.grayscale is False and c.shape[-1] == 3
    raw = cv2.VideoCapture(c.filename).read()[1]
    assert np.array_equal(c.get_frame(0), raw[..., ::-1])     # BGR -> RGB (video.py:504-505)
    r = VideoReader.from_filepath(p, example_indices=[4, 9])
    ex = list(r)
    assert len(r) == 2 and [int(e["frame_ind"]) for e in ex] == [4, 9]
    assert set(ex[0]) == set(r.output_keys) and np.array_equal(ex[1]["image"], v.get_frame(9))


@pytest.mark.parametrize("workers,chunk", [(1, 1), (3, 2), (4, 8)])
def test_feeder_matches_sequential_decode(tmp_path, workers, chunk):
    p = _write_video(tmp_path / "g.avi")
    v = Video.from_filename(p)
    want = v.get_frames(range(len(v)))
    with FrameFeeder(p, batch_size=5, n_workers=workers, chunk_batches=chunk, pinned=False) as f:
        assert len(f) == 37 and f.n_batches == 8
        got, inds = [], []
        for ids, batch in f.batches():
            got.append(batch.copy())
            inds.extend(ids.tolist())
    assert inds == list(range(37))
    assert np.array_equal(np.concatenate(got), want)


def test_feeder_sequence_protocol_and_hold(tmp_path):
    """The slices predict_batches asks for: batch 0 twice (shape probe, then submit), then k+1 while k is in flight."""
    p = _write_video(tmp_path / "g.avi")
    want = Video.from_filename(p).get_frames(range(37))
    f = FrameFeeder(VideoReader.from_filepath(p, example_indices=range(3, 37)), batch_size=4, n_workers=2, chunk_batches=1,
                    pinned=False)
    assert len(f) == 34
    a = f[0:4]
    assert np.array_equal(f[0:4], a)
    b = f[4:8]
    c = f[8:12]
    assert np.array_equal(b, want[7:11]) and np.array_equal(c, want[11:15])
    assert np.array_equal(a, want[3:7])                      # still valid two batches later (hold = 2)
    assert np.array_equal(f[20:23], want[23:26])             # random access falls back to a synchronous decode
    f.close()


def test_feeder_over_array_video_and_error_propagation(tmp_path):
    arr = np.random.default_rng(1).integers(0, 255, size=(11, 8, 8, 1), dtype=np.uint8)
    f = FrameFeeder(Video.from_numpy(arr), batch_size=4, pinned=False)
    assert np.array_equal(np.concatenate([b.copy() for _, b in f.batches()]), arr)
    p = _write_video(tmp_path / "g.avi", n=10)
    bad = FrameFeeder(VideoReader.from_filepath(p, example_indices=[1, 2, 500]), batch_size=2, pinned=False)
    with pytest.raises(KeyError):
        list(bad.batches())
