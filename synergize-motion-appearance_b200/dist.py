"""Frame sharding across the GPUs of one box.

Driving frames are independent once the per-clip constants exist (basicsr/demo.py:117-132 carries no
state between iterations), so rank r renders a contiguous block of frames and a single all-gather of
the uint8 clip reassembles it: no tensor/pipeline parallelism, no collective inside the data path.
Works with NCCL (GPU tensors, NVLink/NVSwitch) and gloo (CPU tensors, used by the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(n/world) frames per rank (the last ranks may get fewer or none)."""
    per = (n_frames + world - 1) // world
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


def gather_clip(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """local: (n_local,H,W,3) uint8 frames of this rank's shard_range -> (n_frames,H,W,3) on every rank.
    One all_gather over equally padded blocks."""
    world = dist.get_world_size(group)
    per = (n_frames + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]].copy_(local)
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        parts: List[torch.Tensor] = list(out.view(world, per, *local.shape[1:]).unbind(0))
        dist.all_gather(parts, pad, group=group)
    return out[:n_frames]


def shard_sources(n_sources: int, rank: int, world: int) -> List[int]:
    """Cross-reenactment batches (16 identities x 64 frames): partition by source so that the cached source
    features stay local to one GPU."""
    lo, hi = shard_range(n_sources, rank, world)
    return list(range(lo, hi))
