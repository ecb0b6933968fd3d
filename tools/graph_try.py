"""Experiment: the steady-state 64-frame micro-batch step replayed from a CUDA graph against the eager enqueue (does the launch gap matter?)."""
import os, sys, json, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sma_oracle as O
import sma_b200 as S
from conftest import CFG, GOLD
inv = json.load(open(os.path.join(GOLD, 'state_keys.json')))
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(O.synthetic_state_dict(inv['net_g'], seed=0), strict=True); me.load_state_dict(O.synthetic_state_dict(inv['motion_estimator'], seed=1), strict=True)
g, me = g.eval().cuda(), me.eval().cuda()
src, drv = O.synthetic_frames(64, seed=77)
frames = torch.stack(drv).cuda()
anim = S.ClipAnimator(g, me, src.unsqueeze(0).cuda(), None, True, True, 1.0)
ref = anim.step(frames).clone()          # first step: clip key-points
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, k=10):
    ms = 0.0
    for _ in range(k):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / k


for _ in range(3):
    out = anim.step(frames)
assert torch.equal(out, ref)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        anim.step(frames)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    gout = anim.step(frames)
graph.replay(); torch.cuda.synchronize()
print('graph output identical:', torch.equal(gout, ref))
for rep in range(3):
    print('B64 eager %.3f ms   graph %.3f ms' % (timeit(lambda: anim.step(frames)), timeit(graph.replay)))
for Bs in (1, 4, 16):
    fr = frames[:Bs].contiguous()
    r0 = anim.step(fr).clone()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            anim.step(fr)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        go = anim.step(fr)
    gr.replay(); torch.cuda.synchronize()
    t0 = time.perf_counter(); anim.step(fr); t1 = time.perf_counter(); torch.cuda.synchronize()
    print('B%d identical %s  eager %.3f ms (host enqueue %.2f ms)  graph %.3f ms' % (Bs, torch.equal(go, r0), timeit(lambda: anim.step(fr)), 1e3 * (t1 - t0), timeit(gr.replay)))
