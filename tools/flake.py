"""Race hunt: the 64-frame clip through the public API from host uint8 frames and from fp32 tensors, and through the device-resident ClipAnimator,
repeated; every run must be bit-identical to the first.  python tools/flake.py [iterations]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sma_oracle as O
import sma_b200 as S
from conftest import CFG, GOLD

inv = json.load(open(os.path.join(GOLD, 'state_keys.json')))
wg, wm = O.synthetic_state_dict(inv['net_g'], seed=0), O.synthetic_state_dict(inv['motion_estimator'], seed=1)
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(wg, strict=True); me.load_state_dict(wm, strict=True)
g, me = g.eval().cuda(), me.eval().cuda()
NF = int(os.environ.get('FLAKE_FRAMES', '64'))
src, drv = O.synthetic_frames(NF, seed=77)
src8, drv8 = O.to_uint8(src), [O.to_uint8(f) for f in drv]
f32 = [(torch.from_numpy(f.astype(np.float32) / 255.).permute(2, 0, 1) - 0.5) / 0.5 for f in drv8]
s32 = (torch.from_numpy(src8.astype(np.float32) / 255.).permute(2, 0, 1) - 0.5) / 0.5
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
first = None
bad = 0
for it in range(n):
    for kind in (sys.argv[2].split(',') if len(sys.argv) > 2 else ('u8', 'f32', 'dev')):
        if kind == 'u8':
            p, _ = S.make_animation(src8, drv8, g, me, relative=True, adapt_movement_scale=True, batch=NF)
            p = np.stack(p)
        elif kind == 'f32':
            p, _ = S.make_animation(s32, f32, g, me, relative=True, adapt_movement_scale=True, batch=NF)
            p = np.stack(p)
        else:
            g.clear_source_cache(); me.dense_motion_network.clear_source_cache()
            anim = S.ClipAnimator(g, me, s32.unsqueeze(0).cuda(), None, True, True, 1.0)
            p = anim.step(torch.stack(f32).cuda()).cpu().numpy()
        if first is None:
            first = p.copy()
        d = (p != first)
        if d.any():
            bad += 1
            fr = np.nonzero(d.reshape(NF, -1).any(1))[0]
            print(f'iteration {it} {kind}: {int(d.sum())} differing values in frames {fr.tolist()[:16]} max level diff {int(np.abs(p.astype(int) - first.astype(int)).max())}', flush=True)
print('iterations', n, 'mismatching runs', bad, {k: v for k, v in os.environ.items() if k.startswith('SMA_')})
