"""CPU: the oracle restatement against the fixtures produced by the live reference (oracle/make_golden.py)."""
import numpy as np
import torch

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
