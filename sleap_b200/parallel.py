"""Frame sharding across GPUs (one process per GPU) and the single exchange step of the path:
an all-gather of fixed-size instance records.  The reference is single-GPU
(sleap/nn/system.py:29-46 raises when more than one GPU is visible); frames are independent, so
ranks own contiguous chunks of every global batch and weights are replicated.

Backend: ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def frame_shard(n_frames: int, rank: int, world: int) -> slice:
    """Contiguous chunk of a global batch owned by ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def record_width(max_instances: int, n_nodes: int) -> int:
    return max_instances * n_nodes * 3 + max_instances + 1


def pack_records(instance_peaks, instance_peak_vals, instance_scores, n_valid) -> torch.Tensor:
    """(B,I,C,2), (B,I,C), (B,I), (B,) -> (B, I*C*3 + I + 1) float32 records (tensors, any device)."""
    B = instance_peaks.shape[0]
    return torch.cat([instance_peaks.reshape(B, -1), instance_peak_vals.reshape(B, -1),
                      instance_scores.reshape(B, -1), n_valid.reshape(B, 1).to(torch.float32)], dim=1).contiguous()


def unpack_records(rec: torch.Tensor, max_instances: int, n_nodes: int):
    B, I, C = rec.shape[0], max_instances, n_nodes
    o = 0
    peaks = rec[:, o:o + I * C * 2].reshape(B, I, C, 2); o += I * C * 2
    vals = rec[:, o:o + I * C].reshape(B, I, C); o += I * C
    scores = rec[:, o:o + I]; o += I
    n_valid = rec[:, o].round().to(torch.int64)
    return peaks, vals, scores, n_valid


def all_gather_records(local: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """Every rank ends up with the records of the whole global batch, in frame order
    (rank-major = frame order because shards are contiguous and equally sized)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, local)
    else:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        out.copy_(torch.cat(parts, dim=0))
    return out
