"""CPU, world_size 2 over gloo: frame sharding + the single all-gather reassemble the clip bit-for-bit."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    from importlib import import_module
    di = import_module('synergize-motion-appearance_b200.dist')
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    clip = torch.randint(0, 256, (n_frames, 8, 8, 3), generator=g, dtype=torch.uint8)   # same on every rank
    lo, hi = di.shard_range(n_frames, rank, world)
    out = di.gather_clip(clip[lo:hi].clone(), n_frames)
    q.put((rank, bool(torch.equal(out, clip))))
    dist.destroy_process_group()


def _run(n_frames, port):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(ok for _, ok in res), res


def test_gather_even():
    _run(8, 29611)


def test_gather_ragged():
    _run(5, 29612)
