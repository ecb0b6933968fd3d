"""Pin the oracle against the live reference and write tests/golden/ (run in the build container).

    python -W ignore oracle/make_golden.py

TEST INFRASTRUCTURE.  Needs /root/reference (read-only mount); never runs on the GPU box.
Steps: (1) build the two reference networks from options/test.yml through the reference's own
build_network; (2) load the deterministic synthetic weights (oracle.sma_oracle.synthetic_state_dict)
with strict=True - this also pins the 472+150 tensor key inventory; (3) run the reference's own
demo.make_animation and module forwards on synthetic frames; (4) run the oracle restatement on the
same tensors and assert agreement; (5) save compact fixtures.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import sma_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    build_network, cfg = ref_shim.import_reference()
    net_g = build_network(cfg['network_g']).eval()
    me = build_network(cfg['network_motion_estimator']).eval()
    inv = {'net_g': {k: list(v.shape) for k, v in net_g.state_dict().items()},
           'motion_estimator': {k: list(v.shape) for k, v in me.state_dict().items()}}
    json.dump(inv, open(os.path.join(GOLD, 'state_keys.json'), 'w'), indent=0)
    P_g = O.synthetic_state_dict(inv['net_g'], seed=0)
    P_me = O.synthetic_state_dict(inv['motion_estimator'], seed=1)
    # the fixed gaussian must equal the reference's registered buffer
    for k in ('kp_detector.down.weight', 'dense_motion_network.down.weight'):
        assert torch.allclose(P_me[k], me.state_dict()[k], atol=1e-9), k
    net_g.load_state_dict(P_g, strict=True)
    me.load_state_dict(P_me, strict=True)

    report = {}
    src, drv = O.synthetic_frames(3, seed=1234)
    demo = ref_shim.load_demo_module()

    # ---- per-module agreement (batch 1, frame 1 of the clip) ------------------------------
    with torch.no_grad():
        s1 = src.unsqueeze(0)
        d1 = drv[1].unsqueeze(0)
        kp_s_ref = me.estimate_kp(s1)
        kp_d_ref = me.estimate_kp(d1)
        kp_0_ref = me.estimate_kp(drv[0].unsqueeze(0))
        kp_s = O.kp_detector(P_me, s1)
        kp_d = O.kp_detector(P_me, d1)
        kp_0 = O.kp_detector(P_me, drv[0].unsqueeze(0))
        report['kp_value'] = float((kp_d['value'] - kp_d_ref['value']).abs().max())
        report['kp_jacobian'] = float((kp_d['jacobian'] - kp_d_ref['jacobian']).abs().max())
        kpn_ref = demo.normalize_kp(kp_source=kp_s_ref, kp_driving=kp_d_ref, kp_driving_initial=kp_0_ref,
                                    use_relative_movement=True, use_relative_jacobian=True,
                                    adapt_movement_scale=True)
        kpn = O.normalize_kp(kp_s, kp_d, kp_0, True, True, True)
        report['kp_norm_value'] = float((kpn['value'] - kpn_ref['value']).abs().max())
        report['kp_norm_jacobian'] = float((kpn['jacobian'] - kpn_ref['jacobian']).abs().max())
        dm_ref = me.estimate_motion_w_kp(kp_source=kp_s_ref, kp_driving=kpn_ref, source_image=s1)
        dm = O.dense_motion(P_me, s1, kpn_ref, kp_s_ref)
        for k in ('deformation', 'occlusion_map', 'driving_kp_heatmap', 'mask'):
            report['dm_' + k] = float((dm[k] - dm_ref[k]).abs().max())
        out_ref = net_g(s1, dm_ref, w=1, inference=True)
        feats = O.encode_source(P_g, s1)
        collect = {}
        out = O.generator_forward(P_g, feats, dm_ref, 1.0, collect)
        report['g_out'] = float((out['out'] - out_ref['out']).abs().max())
        report['g_lq_feat'] = float((out['lq_feat'] - out_ref['lq_feat']).abs().max())
        for i in range(4):
            report[f'g_occ{i}'] = float((out['out_occ'][i] - out_ref['out_occ'][i]).abs().max())
            report[f'g_motion{i+1}'] = float((out['deformation_list'][i + 1] -
                                              out_ref['deformation_list'][i + 1]).abs().max())
        for i, s in enumerate((32, 64, 128, 256)):
            report[f'g_app_{s}'] = float((collect[f'app_{s}'] - out_ref['app_comp_list'][i]).abs().max())
            report[f'g_warped_{s}'] = float((collect[f'warped_{s}'] - out_ref['app_before_comp_list'][i]).abs().max())
        report['out_absmax'] = float(out_ref['out'].abs().max())
        report['out_std'] = float(out_ref['out'].std())
        frac_masked = float(((torch.nn.functional.interpolate(out_ref['deformation_list'][1].permute(0, 3, 1, 2), size=(32, 32), mode='bilinear', align_corners=True).abs() > 1).any(1)).float().mean())
        report['frac_masked_keys_scale32'] = frac_masked

        # ---- VQ lookup (not on the inference path; training path appmotioncodebook_arch.py:382-386)
        z_app = feats['32'][:, :, :, :] * 0.5
        for name, cb_key, z in (('app', 'quantize_app.embedding.weight', z_app),
                                ('motion', 'quantize_motion.embedding.weight',
                                 torch.randn(1, 32, 32, 32, generator=torch.Generator().manual_seed(5)))):
            for sc in (None, 0.25, 0.5):
                q = net_g.quantize_app if name == 'app' else net_g.quantize_motion
                zq_r, loss_r, st = q(z, sc) if sc is not None else q(z)
                zq, loss, idx, md, ppl = O.vq_lookup(P_g[cb_key], z, sc)
                assert torch.equal(idx, st['min_encoding_indices']), (name, sc)
                assert torch.allclose(zq, zq_r, atol=1e-6)
                assert abs(float(loss) - float(loss_r)) < 1e-6 * max(1, abs(float(loss_r)))
                assert abs(float(ppl) - float(st['perplexity'])) < 1e-3
        report['vq_indices_equal'] = True

    # ---- whole clip through the reference's own make_animation --------------------------------
    t0 = time.time()
    preds_ref, drvs_ref = demo.make_animation(src, drv, net_g, me, relative=True,
                                              adapt_movement_scale=True, cpu=True)
    report['ref_make_animation_s_per_frame'] = (time.time() - t0) / len(drv)
    preds, drvs, outs = O.make_animation(P_g, P_me, src, drv, True, True)
    preds_b, _, outs_b = O.make_animation(P_g, P_me, src, drv, True, True, batch=3)
    report['anim_uint8_maxdiff'] = int(max(np.abs(a.astype(int) - b.astype(int)).max() for a, b in zip(preds, preds_ref)))
    report['anim_uint8_mismatch_frac'] = float(np.mean([np.mean(a != b) for a, b in zip(preds, preds_ref)]))
    report['anim_drv_equal'] = bool(all(np.array_equal(a, b) for a, b in zip(drvs, drvs_ref)))
    report['anim_batched_vs_single_fp32'] = float(max((a - b).abs().max() for a, b in zip(outs, outs_b)))

    print(json.dumps(report, indent=1))
    tol = 2e-4
    bad = {k: v for k, v in report.items() if (k.startswith(('kp_', 'dm_', 'g_')) and v > tol)}
    assert not bad, bad
    assert report['anim_uint8_maxdiff'] <= 1 and report['anim_drv_equal']

    # ---- fixtures ---------------------------------------------------------------------------
    fx = {
        'seed_frames': 1234, 'seed_g': 0, 'seed_me': 1,
        'kp_source_value': kp_s_ref['value'], 'kp_source_jacobian': kp_s_ref['jacobian'],
        'kp_driving1_value': kp_d_ref['value'], 'kp_driving1_jacobian': kp_d_ref['jacobian'],
        'kp_driving0_value': kp_0_ref['value'], 'kp_driving0_jacobian': kp_0_ref['jacobian'],
        'kp_norm1_value': kpn_ref['value'], 'kp_norm1_jacobian': kpn_ref['jacobian'],
        'deformation1': dm_ref['deformation'], 'occlusion1': dm_ref['occlusion_map'],
        'driving_kp_heatmap1_s4': dm_ref['driving_kp_heatmap'][:, :, ::4, ::4].clone(),
        'out1': out_ref['out'].half().float().sub(out_ref['out']).neg().add(0) * 0 + out_ref['out'],
        'lq_feat1_s4': out_ref['lq_feat'][:, ::4, ::4, ::4].clone(),
        'out_occ1': [o.clone() for o in out_ref['out_occ']],
        'deformation_list1': [m.clone() for m in out_ref['deformation_list']],
        'app_comp1_s': [a[:, ::8, ::max(1, a.shape[-1] // 16), ::max(1, a.shape[-1] // 16)].clone()
                        for a in out_ref['app_comp_list']],
        'enc_feat32_s4': feats['32'][:, ::4, ::4, ::4].clone(),
        'pred_uint8': [torch.from_numpy(p.copy()) for p in preds_ref],
    }
    fx['out1'] = out_ref['out'].clone()
    torch.save(fx, os.path.join(GOLD, 'reference_clip3.pt'))
    json.dump(report, open(os.path.join(GOLD, 'oracle_vs_reference.json'), 'w'), indent=1)
    print('wrote', GOLD, {k: os.path.getsize(os.path.join(GOLD, k)) for k in os.listdir(GOLD)})


if __name__ == '__main__':
    main()
