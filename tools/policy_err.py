"""Scratch (GPU box): end-to-end fp32 error of the batched clip path against the oracle, per precision policy.
usage: policy_err.py [seed ...]   ; prints one row per policy: max-abs error of `out` on sampled frames, key-point / deformation errors."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import sma_b200 as S
import sma_oracle as O
from conftest import CFG
inv = json.load(open(os.path.join(ROOT, 'tests/golden/state_keys.json')))
SIZE = 512 if 'size512' in sys.argv else 256
P_g, P_me = O.synthetic_state_dict(O.variant_shapes(inv['net_g'], SIZE), 0), O.synthetic_state_dict(inv['motion_estimator'], 1)
CFG['network_g']['img_size'] = SIZE
g = S.build_network(CFG['network_g']); me = S.build_network(CFG['network_motion_estimator'])
g.load_state_dict(P_g); me.load_state_dict(P_me)
g, me = g.eval().cuda(), me.eval().cuda()
SIZE = 512 if 'size512' in sys.argv else 256
sys.argv = [a for a in sys.argv if a != 'size512']
seeds = [int(a) for a in sys.argv[1:] if a.isdigit()] or [77]
policies = [a for a in sys.argv[1:] if not a.isdigit()] or ['exact', 'f16', 'f16+kp', 'f16+s1', 'f16+s3m', 'f16+kp+s1+s3m']
idx = [0, 17, 38, 63]
torch.set_num_threads(os.cpu_count())
for seed in seeds:
    src, drv = O.synthetic_frames(64, seed=seed, size=SIZE)
    sel = [drv[i] for i in idx]
    t0 = time.time()
    with torch.no_grad():
        kp_s = O.kp_detector(P_me, src.unsqueeze(0)); kp_0 = O.kp_detector(P_me, drv[0].unsqueeze(0))
        kp_d = O.kp_detector(P_me, torch.stack(sel))
        kp_n = O.normalize_kp(kp_s, kp_d, kp_0, True, True, True)
        kp_sb = {k: v.expand(4, *v.shape[1:]) for k, v in kp_s.items()}
        dm = O.dense_motion(P_me, src.unsqueeze(0).expand(4, -1, -1, -1), kp_n, kp_sb)
        ref = O.generator_forward(P_g, O.encode_source(P_g, src.unsqueeze(0)), dm, 1.0)
    print(f'seed {seed}: oracle {time.time() - t0:.1f} s; out absmax {float(ref["out"].abs().max()):.2f}', flush=True)
    for mode in policies:
        S.ops.USE_TF32X3 = not mode.startswith('exact')
        S.ops.USE_F16 = mode.startswith('f16')
        S.ops.FAST_STAGES = set(t for t in mode.split('+')[1:] if not t.startswith('x2:'))
        S.ops.X2_STAGES = set(t[3:] for t in mode.split('+')[1:] if t.startswith('x2:'))
        g.clear_source_cache(); me.dense_motion_network.clear_source_cache()
        anim = S.ClipAnimator(g, me, src.unsqueeze(0).cuda(), drv[0].unsqueeze(0).cuda(), True, True, 1.0)
        frames = torch.stack(sel).cuda()
        anim.clip_keypoints()
        kp = me.estimate_kp(frames)
        kpn = S.normalize_kp(anim.kp_source, kp, anim.kp_initial, adapt_movement_scale=True, use_relative_movement=True, use_relative_jacobian=True, _scale=anim.scale)
        dmg = me.estimate_motion_w_kp(kp_source=anim.kp_source, kp_driving=kpn, source_image=anim.source)
        r = g.generate(anim.feats, dmg['deformation'], dmg['occlusion_map'].view(4, 64 * SIZE // 256, 64 * SIZE // 256), dmg['_driving_kp_heatmap_nhwc'], 1.0)
        out = r['out'].permute(0, 3, 1, 2).cpu()
        e = (out - ref['out']).abs().amax(dim=(1, 2, 3))
        ekp = float((kp['value'].cpu() - kp_d['value']).abs().max()); ekj = float((kp['jacobian'].cpu() - kp_d['jacobian']).abs().max())
        ekn = float((kpn['value'].cpu() - kp_n['value']).abs().max())
        edef = float((dmg['deformation'].cpu() - dm['deformation']).abs().max()); eocc = float((dmg['occlusion_map'].cpu() - dm['occlusion_map']).abs().max())
        em = [float((a.cpu() - b).abs().max()) for a, b in zip(r['deformation_list'][1:], ref['deformation_list'][1:])]
        # the generator alone, fed with the ORACLE's dense motion (isolates KP / S1 from S3m / S3a / S4)
        heat = S.ops.nchw_to_nhwc(dm['driving_kp_heatmap'].cuda().contiguous())
        r2 = g.generate(anim.feats, dm['deformation'].cuda(), dm['occlusion_map'].cuda().view(4, 64 * SIZE // 256, 64 * SIZE // 256), heat, 1.0)
        e2 = (r2['out'].permute(0, 3, 1, 2).cpu() - ref['out']).abs().amax(dim=(1, 2, 3))
        print(f'  {mode:34s} out {["%.2e" % float(v) for v in e]} | gen-only {["%.2e" % float(v) for v in e2]} | kp {ekp:.1e} jac {ekj:.1e} kpn {ekn:.1e} | deform {edef:.1e} occ {eocc:.1e} | m_com {["%.1e" % v for v in em]}', flush=True)
