"""Scratch: end-to-end error of net_g 'out' vs the reference fixture for the exact and the tensor-core conv paths (GPU box)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import sma_b200 as S
import sma_oracle as O
from conftest import CFG
inv = json.load(open(os.path.join(ROOT, 'tests/golden/state_keys.json')))
golden = torch.load(os.path.join(ROOT, 'tests/golden/reference_clip3.pt'))
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(O.synthetic_state_dict(inv['net_g'], 0)); me.load_state_dict(O.synthetic_state_dict(inv['motion_estimator'], 1))
g, me = g.eval().cuda(), me.eval().cuda()
src, drv = O.synthetic_frames(3, seed=1234)
dm = {'deformation': golden['deformation1'].cuda(), 'occlusion_map': golden['occlusion1'].cuda(),
      'driving_kp_heatmap': O.gaussian_heatmaps(golden['kp_norm1_value'], 64, 64).cuda()}
for mode in sys.argv[1:] or ['exact', 'tf32', 'f16', 'f16+kp+s1+s3m']:      # base[+fast stage ...]
    S.ops.USE_TF32X3 = not mode.startswith('exact')
    S.ops.USE_F16 = mode.startswith('f16')
    S.ops.FAST_STAGES = set(mode.split('+')[1:])     # e.g. tc+kp+s1+s3m
    g._src_cache = None
    out = g(src.unsqueeze(0).cuda(), dm, w=1, inference=True)
    d = (out['out'].cpu() - golden['out1']).abs()
    occ = max(float((a.cpu() - b).abs().max()) for a, b in zip(out['out_occ'], golden['out_occ1']))
    mo = max(float((a.cpu() - b).abs().max()) for a, b in zip(out['deformation_list'], golden['deformation_list1']))
    kp = me.estimate_kp(src.unsqueeze(0).cuda())
    kperr = float((kp['value'].cpu() - golden['kp_source_value']).abs().max())
    preds, _ = S.make_animation(src, drv, g, me, batch=3)
    u8 = max(int((torch.from_numpy(p).int() - r.int()).abs().max()) for p, r in zip(preds, golden['pred_uint8']))
    nmis = sum(float((torch.from_numpy(p) != r).float().mean()) for p, r in zip(preds, golden['pred_uint8'])) / 3
    print(f'{mode}: kp {kperr:.2e} | out max {float(d.max()):.3e} mean {float(d.mean()):.3e} | occ {occ:.2e} motion {mo:.2e} | uint8 maxdiff {u8} mismatch frac {nmis:.2e}', flush=True)
