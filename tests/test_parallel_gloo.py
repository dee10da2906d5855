"""N > 1 host logic on CPU: world_size-2 gloo run of the shard + gather step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sleap_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames, I, C, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sl = parallel.frame_shard(n_frames, rank, world)
    idx = torch.arange(sl.start, sl.stop, dtype=torch.float32)
    B = len(idx)
    peaks = idx.view(B, 1, 1, 1).expand(B, I, C, 2).clone() + 0.25
    peaks[:, 1:] = float("nan")
    vals = idx.view(B, 1, 1).expand(B, I, C).clone() * 2
    scores = idx.view(B, 1).expand(B, I).clone() * 3
    nv = torch.ones(B, dtype=torch.int32)
    rec = parallel.pack_records(peaks, vals, scores, nv)
    assert rec.shape == (B, parallel.record_width(I, C))
    allrec = parallel.all_gather_records(rec)
    gp, gv, gs, gn = parallel.unpack_records(allrec, I, C)
    ok = (allrec.shape[0] == n_frames and torch.equal(gp[:, 0, 0, 0], torch.arange(n_frames, dtype=torch.float32) + 0.25)
          and torch.isnan(gp[:, 1:]).all().item() and torch.equal(gs[:, 0], torch.arange(n_frames, dtype=torch.float32) * 3)
          and torch.equal(gn, torch.ones(n_frames, dtype=torch.int64)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shard_covers_all_frames():
    for n, w in [(64, 8), (10, 4), (3, 8), (16, 2)]:
        seen = []
        for r in range(w):
            sl = parallel.frame_shard(n, r, w)
            seen += list(range(sl.start, sl.stop))
        assert seen == list(range(n))


@pytest.mark.timeout(120)
def test_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, 4, 13, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


def _fake_predict(frames):
    """Stand-in for inference_model.predict_on_batch: the frame's first pixel value decides how many instances it
    has (0..2) and is written into their coordinates, so the gathered result identifies frame and order."""
    frames = np.asarray(frames)
    b = len(frames)
    ids = frames.reshape(b, -1)[:, 0].astype(np.float32)
    peaks = np.full((b, 2, 3, 2), np.nan, np.float32); vals = np.full((b, 2, 3), np.nan, np.float32)
    scores = np.full((b, 2), np.nan, np.float32); nv = (ids.astype(np.int64) % 3)
    for i in range(b):
        for j in range(int(nv[i])):
            peaks[i, j] = ids[i] + 0.5 * j
            vals[i, j] = 1.0
            scores[i, j] = ids[i]
    return {"instance_peaks": peaks, "instance_peak_vals": vals, "instance_scores": scores, "n_valid": nv}


def _worker_sharded(rank, world, port, n_frames, gb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.arange(n_frames, dtype=np.uint8).reshape(n_frames, 1, 1, 1) * np.ones((1, 2, 2, 1), np.uint8)
    outs = list(parallel.predict_sharded(_fake_predict, frames, gb, max_instances=4, n_nodes=3))
    want = _fake_predict(frames)
    ok = sum(len(o["n_valid"]) for o in outs) == n_frames
    got_nv = np.concatenate([o["n_valid"] for o in outs]); got_p = np.concatenate([o["instance_peaks"] for o in outs])
    ok = ok and np.array_equal(got_nv, want["n_valid"]) and np.array_equal(np.concatenate([o["frame_ind"] for o in outs]), np.arange(n_frames))
    ok = ok and np.array_equal(np.nan_to_num(got_p[:, :2], nan=-1), np.nan_to_num(want["instance_peaks"], nan=-1)) and np.isnan(got_p[:, 2:]).all()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n_frames,gb", [(16, 8), (11, 4), (3, 8)])
def test_predict_sharded_world2(n_frames, gb):
    """Whole shard -> predict -> gather loop on 2 gloo ranks, including ragged last batches and a batch smaller than
    the world (one rank idle)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, n_frames, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


def test_predict_sharded_single_process():
    frames = np.arange(5, dtype=np.uint8).reshape(5, 1, 1, 1) * np.ones((1, 2, 2, 1), np.uint8)
    outs = list(parallel.predict_sharded(_fake_predict, frames, 2, max_instances=2, n_nodes=3))
    assert [len(o["n_valid"]) for o in outs] == [2, 2, 1]
    assert np.array_equal(np.concatenate([o["n_valid"] for o in outs]), np.arange(5) % 3)
