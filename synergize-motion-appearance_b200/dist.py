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


def make_animation_sharded(source_image, driving_video, net_g, motion_estimator, relative=True, adapt_movement_scale=True,
                           batch: int = 64, w: float = 1.0, bgr: bool = False, group=None) -> torch.Tensor:
    """BASELINE configs[2]: one clip, its driving frames sharded over the ranks of `group` (one process per GPU), reassembled on every rank by
    ONE all-gather of the uint8 frames (basicsr/demo.py:117-132 has no cross-frame state, so the loop splits anywhere).

    Every rank receives the same `source_image` / `driving_video` (as make_animation: fp32 CHW tensors or uint8 HWC frames), renders frames
    shard_range(n, rank, world) in micro-batches of `batch`, and returns the whole clip as an (n,H,W,3) uint8 tensor on its device.  The
    per-clip constants (source features and key-points, initial driving key-points of driving_video[0], movement scale) are recomputed by
    every rank through the same batch-2 key-point pass, so the result does not depend on the number of ranks: for equal `batch` the gathered
    clip is bit-identical to the world-size-1 result."""
    from . import animate, ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = next(net_g.parameters()).device
    n = len(driving_video)
    lo, hi = shard_range(n, rank, world)
    with torch.no_grad(), torch.cuda.device(dev):
        src = animate.frames_to_device([source_image], dev)
        first = animate.frames_to_device([driving_video[0]], dev)
        anim = animate.ClipAnimator(net_g, motion_estimator, src, first, relative, adapt_movement_scale, w)
        H, W = src.shape[2], src.shape[3]
        local = torch.empty((hi - lo, H, W, 3), dtype=torch.uint8, device=dev)
        if hi > lo:
            mine = driving_video[lo:hi]
            upl = animate._Uploader(animate._io(dev), mine, batch)
            ticket = upl.upload(0)
            for i0 in range(0, hi - lo, batch):
                frames, release = upl.take(ticket)
                if i0 + batch < hi - lo:
                    ticket = upl.upload(i0 + batch)
                local[i0:i0 + frames.shape[0]] = anim.step(frames, bgr)
                release()
        return gather_clip(local, n, group) if world > 1 else local
