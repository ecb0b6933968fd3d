"""CPU: the oracle restatement against the fixtures produced by the live reference (oracle/make_golden.py)."""
import os

import numpy as np
import torch

from conftest import GOLD

import sma_oracle as O


def test_oracle_matches_reference_fixtures(golden, weights, clip):
    P_g, P_me = weights
    src, drv = clip
    with torch.no_grad():
        s1, d1 = src.unsqueeze(0), drv[1].unsqueeze(0)
        kp_s = O.kp_detector(P_me, s1)
        kp_d = O.kp_detector(P_me, d1)
        kp_0 = O.kp_detector(P_me, drv[0].unsqueeze(0))
        assert torch.allclose(kp_s['value'], golden['kp_source_value'], atol=1e-6)
        assert torch.allclose(kp_d['jacobian'], golden['kp_driving1_jacobian'], atol=1e-6)
        kpn = O.normalize_kp(kp_s, kp_d, kp_0, True, True, True)
        assert torch.allclose(kpn['value'], golden['kp_norm1_value'], atol=1e-6)
        assert torch.allclose(kpn['jacobian'], golden['kp_norm1_jacobian'], atol=1e-5)
        dm = O.dense_motion(P_me, s1, kpn, kp_s)
        assert torch.allclose(dm['deformation'], golden['deformation1'], atol=1e-5)
        assert torch.allclose(dm['occlusion_map'], golden['occlusion1'], atol=1e-5)
        feats = O.encode_source(P_g, s1)
        assert torch.allclose(feats['32'][:, ::4, ::4, ::4], golden['enc_feat32_s4'], atol=1e-5)
        out = O.generator_forward(P_g, feats, dm, 1.0)
        # 1e-3 is the north-star tolerance; the oracle itself sits at fp32 noise from the reference
        assert float((out['out'] - golden['out1']).abs().max()) < 1e-4
        for a, b in zip(out['out_occ'], golden['out_occ1']):
            assert torch.allclose(a, b, atol=1e-5)
        for a, b in zip(out['deformation_list'], golden['deformation_list1']):
            assert torch.allclose(a, b, atol=1e-5)
        u8 = O.to_uint8(out['out'][0])
        assert np.abs(u8.astype(int) - golden['pred_uint8'][1].numpy().astype(int)).max() <= 1


def test_vq_lookup_edge_cases():
    g = torch.Generator().manual_seed(3)
    for E, init in ((256, 'normal'), (32, 'normal'), (256, 'tiny'), (32, 'tiny')):
        cb = torch.randn(1024, E, generator=g) if init == 'normal' else (torch.rand(1024, E, generator=g) * 2 - 1) / 1024
        z = torch.randn(2, E, 32, 32, generator=g)
        for scale in (None, 0.25, 0.75):
            zq, loss, idx, md, ppl = O.vq_lookup(cb, z, scale)
            n = 1024 if scale is None else int(scale * 1024)
            assert idx.shape == (2048, 1) and idx.dtype == torch.int64 and int(idx.max()) < n
            # gathered rows are codebook rows
            assert torch.equal(zq.permute(0, 2, 3, 1).reshape(-1, E), cb[idx[:, 0]])
    # exact duplicates in the codebook -> lowest index wins
    cb = torch.randn(16, 32, generator=g)
    cb[9] = cb[4]
    z = cb[4].view(1, 32, 1, 1).expand(1, 32, 32, 32).contiguous()
    _, _, idx, _, _ = O.vq_lookup(cb, z)
    assert int(idx.min()) == 4 and int(idx.max()) == 4


def test_hull_area_matches_scipy():
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(0)
    for _ in range(20):
        p = rng.normal(size=(15, 2))
        assert abs(O.hull_area(p) - ConvexHull(p).volume) < 1e-9


def test_oracle_caller_surface_matches_reference_fixture(weights):
    """the plain decoder `net_g.generator(lq_feat)` and `motion_estimator(driving, source)` restated by the oracle agree with the
    outputs of the live reference stored by oracle/make_golden_callers.py."""
    import os
    import torch
    import sma_oracle as O
    from conftest import GOLD
    fx = torch.load(os.path.join(GOLD, 'reference_callers.pt'))
    assert max(fx['oracle_vs_reference'].values()) < 2e-4
    P_g, P_me = weights
    lq = torch.randn(1, 256, 32, 32, generator=torch.Generator().manual_seed(fx['lq_seed'])) * 0.5
    with torch.no_grad():
        recon = O.decode_plain(P_g, lq)
    assert float((recon[:, :, ::2, ::2] - fx['recon_s2']).abs().max()) < 2e-4


def test_oracle_training_forward_matches_reference_fixture(weights, clip):
    """The oracle's inference=False restatement and `encode_driving` against the outputs of the live reference (oracle/make_golden_train.py):
    subsampled tensors and the eight codebook losses."""
    fx = torch.load(os.path.join(GOLD, 'reference_train1.pt'))
    src, drv = clip
    P_g = weights[0]
    dm = {k: fx[k] for k in ('deformation', 'occlusion_map', 'driving_kp_heatmap')}
    with torch.no_grad():
        out = O.generator_forward_train(P_g, O.encode_source(P_g, src.unsqueeze(0)), dm, 1.0, gt=drv[1].unsqueeze(0))
        ed = O.encode_driving(P_g, drv[1].unsqueeze(0))
    sub = lambda t: t[:, ::max(1, t.shape[1] // 16), ::max(1, t.shape[2] // 16), ::max(1, t.shape[3] // 16)]
    assert float((out['out_lr'][0][:, :, ::4, ::4] - fx['out_lr_s4']).abs().max()) < 1e-5
    for i in range(4):
        assert float((out['motion_recon_list'][i] - fx['motion_recon_list'][i]).abs().max()) < 1e-5
        assert abs(float(out['codebook_loss_motion_list'][i]) - fx['codebook_loss_motion_list'][i]) < 1e-5 * fx['codebook_loss_motion_list'][i]
        assert abs(float(out['codebook_loss_app_list'][i]) - fx['codebook_loss_app_list'][i]) < 1e-5 * fx['codebook_loss_app_list'][i]
        for j in range(5):
            assert float((sub(out['app_recon_list'][i][j]) - fx['app_recon_s'][i][j]).abs().max()) < 1e-5, (i, j)
    for k, v in ed.items():
        assert float((sub(v) - fx['encode_driving_s'][k]).abs().max()) < 1e-5, k
