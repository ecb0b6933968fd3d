"""Scratch: fp32 error map of one (source seed, frame) against the oracle, and whether the key-padding masks agree."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, torch.nn.functional as F
import sma_b200 as S
import sma_oracle as O
from conftest import CFG
inv = json.load(open(os.path.join(ROOT, 'tests/golden/state_keys.json')))
P_g, P_me = O.synthetic_state_dict(inv['net_g'], 0), O.synthetic_state_dict(inv['motion_estimator'], 1)
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(P_g); me.load_state_dict(P_me)
g, me = g.eval().cuda(), me.eval().cuda()
src, drv = O.synthetic_frames(3, seed=21)
with torch.no_grad():
    kp_s = O.kp_detector(P_me, src.unsqueeze(0)); kp_0 = O.kp_detector(P_me, drv[0].unsqueeze(0))
    kp_d = O.kp_detector(P_me, torch.stack(drv))
    kp_n = O.normalize_kp(kp_s, kp_d, kp_0, True, True, True)
    kp_sb = {k: v.expand(3, *v.shape[1:]) for k, v in kp_s.items()}
    dm = O.dense_motion(P_me, src.unsqueeze(0).expand(3, -1, -1, -1), kp_n, kp_sb)
    col = {}
    ref = O.generator_forward(P_g, O.encode_source(P_g, src.unsqueeze(0)), dm, 1.0, col)
for mode in ('f16', 'exact'):
    S.ops.USE_TF32X3 = mode != 'exact'
    g.clear_source_cache()
    anim = S.ClipAnimator(g, me, src.unsqueeze(0).cuda(), drv[0].unsqueeze(0).cuda(), True, True, 1.0)
    u8, out = anim.step(torch.stack(drv).cuda(), want_fp32=True)
    e = (out.permute(0, 3, 1, 2).cpu() - ref['out']).abs()
    print(mode, 'out err max per frame', e.amax(dim=(1, 2, 3)).tolist(), 'mean', e.mean(dim=(1, 2, 3)).tolist())
    # generator alone on the oracle's dense motion, with stage taps
    heat = S.ops.nchw_to_nhwc(dm['driving_kp_heatmap'].cuda().contiguous())
    gc = {}
    r2 = g.generate(anim.feats, dm['deformation'].cuda(), dm['occlusion_map'].cuda().view(3, 64, 64), heat, 1.0, collect=gc)
    e2 = (r2['out'].permute(0, 3, 1, 2).cpu() - ref['out']).abs()
    print(mode, 'gen-only err max', e2.amax(dim=(1, 2, 3)).tolist(), 'mean', e2.mean(dim=(1, 2, 3)).tolist())
    for k in sorted(gc):
        if k in col:
            d = (gc[k].permute(0, 3, 1, 2).cpu() - col[k]).abs()
            print(f'   {k:12s} max per frame {["%.1e" % v for v in d.amax(dim=(1, 2, 3)).tolist()]} mean {["%.1e" % v for v in d.mean(dim=(1, 2, 3)).tolist()]}')
    for i, (a, b) in enumerate(zip(r2['deformation_list'][1:], ref['deformation_list'][1:])):
        m_o = F.interpolate(b.permute(0, 3, 1, 2), size=(32, 32), mode='bilinear', align_corners=True)
        m_g = F.interpolate(a.cpu().permute(0, 3, 1, 2), size=(32, 32), mode='bilinear', align_corners=True)
        ig_o = ((m_o > 1) | (m_o < -1)).any(1); ig_g = ((m_g > 1) | (m_g < -1)).any(1)
        near = ((m_o.abs() - 1).abs() < 1e-4).any(1)
        print(f'   scale {i}: masked keys per frame {ig_o.flatten(1).sum(1).tolist()} mask mismatches {(ig_o != ig_g).flatten(1).sum(1).tolist()} |m| within 1e-4 of 1: {near.flatten(1).sum(1).tolist()}')
