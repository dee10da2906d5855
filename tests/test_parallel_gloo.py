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
