"""Pin the oracle's inference=False restatement and `encode_driving` against the live reference (build container only).

    python -W ignore oracle/make_golden_train.py

TEST INFRASTRUCTURE.  Runs the unmodified reference `AppMotionCompFormer.forward(source, dense_motion, w=1, inference=False, gt=driving)` and
`encode_driving(driving)` on the synthetic weights / frames of make_golden.py, asserts that oracle.generator_forward_train / encode_driving
reproduce every added output (forward values), and writes tests/golden/reference_train1.pt.
"""
import json
import os
import sys

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
    inv = json.load(open(os.path.join(GOLD, 'state_keys.json')))
    P_g, P_me = O.synthetic_state_dict(inv['net_g'], seed=0), O.synthetic_state_dict(inv['motion_estimator'], seed=1)
    net_g = build_network(cfg['network_g']).eval(); net_g.load_state_dict(P_g, strict=True)
    me = build_network(cfg['network_motion_estimator']).eval(); me.load_state_dict(P_me, strict=True)
    src, drv = O.synthetic_frames(3, seed=1234)
    s1, d1 = src.unsqueeze(0), drv[1].unsqueeze(0)
    rep = {}
    with torch.no_grad():
        dm_ref = me(d1, s1)                                   # Motion_Estimator_keypoint_aware.forward(driving, source): the training-time call
        ref = net_g(s1, dm_ref, w=1, inference=False, gt=d1)
        out = O.generator_forward_train(P_g, O.encode_source(P_g, s1), dm_ref, 1.0, gt=d1)
        rep['out'] = float((out['out'] - ref['out']).abs().max())
        rep['out_lr'] = float((out['out_lr'][0] - ref['out_lr'][0]).abs().max())
        for i in range(4):
            rep[f'motion_recon{i}'] = float((out['motion_recon_list'][i] - ref['motion_recon_list'][i]).abs().max())
            rep[f'loss_motion{i}'] = abs(float(out['codebook_loss_motion_list'][i]) - float(ref['codebook_loss_motion_list'][i])) / float(ref['codebook_loss_motion_list'][i])
            rep[f'loss_app{i}'] = abs(float(out['codebook_loss_app_list'][i]) - float(ref['codebook_loss_app_list'][i])) / float(ref['codebook_loss_app_list'][i])
            for j, nm in enumerate(('app_recon', 'app_feat_original', 'quant_app', 'app_feat', 'feat_com')):
                a, b = out['app_recon_list'][i][j], ref['app_recon_list'][i][j]
                assert a.shape == b.shape, (nm, i, a.shape, b.shape)
                rep[f'{nm}{i}'] = float((a - b).abs().max())
        ed_ref = net_g.encode_driving(d1)
        ed = O.encode_driving(P_g, d1)
        assert sorted(ed) == sorted(ed_ref)
        for k in ed:
            rep[f'encode_driving_{k}'] = float((ed[k] - ed_ref[k]).abs().max())
        rep['encode_driving_32_vs_latent'] = float((ed_ref['32'] - O.encode_source(P_g, d1)['32']).abs().max())      # (they differ: block 11 vs the last block)
    print(json.dumps(rep, indent=1))
    bad = {k: v for k, v in rep.items() if v > 2e-4 and k != 'encode_driving_32_vs_latent'}
    assert not bad, bad
    assert rep['encode_driving_32_vs_latent'] > 1e-2
    sub = lambda t: t[:, ::max(1, t.shape[1] // 16), ::max(1, t.shape[2] // 16), ::max(1, t.shape[3] // 16)].clone()
    fx = {
        'note': 'reference forward(source, me(driving1, source), w=1, inference=False, gt=driving1) on the make_golden.py weights / frames',
        'deformation': dm_ref['deformation'].clone(), 'occlusion_map': dm_ref['occlusion_map'].clone(),
        'driving_kp_heatmap': dm_ref['driving_kp_heatmap'].clone(),
        'out_lr_s4': ref['out_lr'][0][:, :, ::4, ::4].clone(),
        'out_s4': ref['out'][:, :, ::4, ::4].clone(),
        'motion_recon_list': [m.clone() for m in ref['motion_recon_list']],
        'codebook_loss_motion_list': [float(v) for v in ref['codebook_loss_motion_list']],
        'codebook_loss_app_list': [float(v) for v in ref['codebook_loss_app_list']],
        'app_recon_s': [[sub(t) for t in row] for row in ref['app_recon_list']],
        'encode_driving_s': {k: sub(v) for k, v in ed_ref.items()},
        'oracle_vs_reference': rep,
    }
    torch.save(fx, os.path.join(GOLD, 'reference_train1.pt'))
    print('wrote', os.path.getsize(os.path.join(GOLD, 'reference_train1.pt')), 'bytes')


if __name__ == '__main__':
    main()
