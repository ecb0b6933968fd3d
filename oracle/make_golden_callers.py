"""Pin the caller-surface pieces of SURVEY.md section 8f(1) against the live reference and write tests/golden/reference_callers.pt
(run in the build container; TEST INFRASTRUCTURE, needs /root/reference):

    python -W ignore oracle/make_golden_callers.py

* `net_g.generator(lq_feat)` - the plain decoder AppMotionCompModel.test calls (models/appmotioncomp_model.py:453-454);
* `motion_estimator(driving, source)` - Motion_Estimator_keypoint_aware.forward (archs/motion_estimator_arch.py:42-52).
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
    src, drv = O.synthetic_frames(2, seed=1234)
    with torch.no_grad():
        lq = torch.randn(1, 256, 32, 32, generator=torch.Generator().manual_seed(77)) * 0.5
        recon_ref = net_g.generator(lq)
        recon = O.decode_plain(P_g, lq)
        d_dec = float((recon - recon_ref).abs().max())
        dm_ref = me(drv[1].unsqueeze(0), src.unsqueeze(0))
        kp_d, kp_s = O.kp_detector(P_me, drv[1].unsqueeze(0)), O.kp_detector(P_me, src.unsqueeze(0))
        dm = O.dense_motion(P_me, src.unsqueeze(0), kp_d, kp_s)
        d_def = float((dm['deformation'] - dm_ref['deformation']).abs().max())
        d_occ = float((dm['occlusion_map'] - dm_ref['occlusion_map']).abs().max())
    print({'decode_plain': d_dec, 'forward_deformation': d_def, 'forward_occlusion': d_occ})
    assert max(d_dec, d_def, d_occ) < 2e-4
    torch.save({'lq_seed': 77, 'recon_s2': recon_ref[:, :, ::2, ::2].clone(), 'fwd_deformation': dm_ref['deformation'].clone(),
                'fwd_occlusion': dm_ref['occlusion_map'].clone(), 'fwd_kp_driving_value': dm_ref['kp_driving']['value'].clone(),
                'oracle_vs_reference': {'decode_plain': d_dec, 'forward_deformation': d_def, 'forward_occlusion': d_occ}},
               os.path.join(GOLD, 'reference_callers.pt'))
    print('wrote', os.path.getsize(os.path.join(GOLD, 'reference_callers.pt')), 'bytes')


if __name__ == '__main__':
    main()
