"""torchrun --nproc-per-node N tools/gather_check.py: the peer-memory record exchange against an NCCL all-gather of the same
records, host consumer and device consumer, more steps than generations.  Prints GATHER_CHECK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ctypes import byref, c_void_p
    from sleap_b200 import parallel
    import test_gpu_gather as T
    model, pred = T._predictor(batch_size=3, seed=31)                      # same weights on every rank
    B, I, C = 3, 16, 13
    w = parallel.record_width(I, C)
    frames = np.random.default_rng(100 + rank).integers(0, 256, size=(B * 13, 128, 128, 1), dtype=np.uint8)
    pred.inference_model.predict_on_batch(frames[:B])                      # configures the device pipeline
    pg = parallel.PeerGather(model, generations=4)
    # (1) host consumer through the public predictor: 13 steps > 4 generations
    outs = pred.predict(frames, make_labels=False)
    assert pg.pushed() == 13 and pg.consumed == 13
    for o in outs:
        mine = torch.from_numpy(T._records_of(o, I, C)).cuda()
        ref = torch.empty((world * B, w), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(ref, mine)
        a, b = np.nan_to_num(o["gathered_records"], nan=-7.0), np.nan_to_num(ref.cpu().numpy(), nan=-7.0)
        assert np.array_equal(a, b), f"rank {rank}: peer-memory records differ from the NCCL all-gather"
        assert list(o["gathered_counts"]) == [B] * world
    # (2) device consumer with lag 2 (the bench loop): windows hold what the ranks' device records held
    h = model.handle
    dev = torch.from_numpy(frames[:B]).cuda()
    rec_ptr = c_void_p()
    h.call("sb_bottomup_device_records", model.model_id, byref(rec_ptr))
    for i in range(9):
        h.call("sb_infer_bottomup_dev", model.model_id, c_void_p(dev.data_ptr()), B)
        if pg.pushed() - pg.consumed > 2:
            pg.consume_next_dev()
    while pg.consumed < pg.pushed():
        pg.consume_next_dev()
    h.synchronize()
    torch.cuda.synchronize()
    assert pg.pushed() == 22 and pg.consumed == 22
    dist.barrier()
    pg.close()
    dist.barrier()
    if rank == 0:
        print("GATHER_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
