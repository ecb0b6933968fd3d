import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLD = os.path.join(ROOT, 'tests', 'golden')

CFG = {
    'network_g': dict(type='AppMotionCompFormer', with_position_emb=True, img_size=256, nf=64, ch_mult=[1, 2, 2, 4], num_kp=15,
                      quantizer_type='nearest', beta=0.25, n_head=8, warp_s_d_kp_query=True, MRFA_motion_enc=True,
                      motion_codebook_split=True, multiscale_feature_fusion=True, codebook_size_motion=1024,
                      embed_dim_motion=32, dim_embd_motion=32, n_layers_motion=2, codebook_size_app=1024, embed_dim_app=256,
                      dim_embd_app=256, n_layers_app=2, split=1, app_codebook_split=True, connect_list=['64', '128', '256'],
                      connect_app_list=['32', '64', '128', '256'], fix_modules=[], ae_path=None),
    'network_motion_estimator': dict(type='Motion_Estimator_keypoint_aware', common_params=dict(num_kp=15, num_channels=3),
                                     dense_motion_params=dict(block_expansion=64, max_features=1024, num_blocks=5,
                                                              scale_factor=0.25, estimate_occlusion_map=True),
                                     kp_detector_params=dict(temperature=0.1, block_expansion=32, max_features=1024,
                                                             scale_factor=0.25, num_blocks=5, estimate_jacobian=True)),
}


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _poison_device_memory():
    """SMA_POISON=1: fill what the caching allocator will hand out with 0xFF bytes (NaN as fp32 / fp16) before any test runs, so that a kernel
    that reads memory nobody wrote (and multiplies it by zero) shows up as NaN instead of passing on a box whose stale memory happens to be finite."""
    free, _ = torch.cuda.mem_get_info()
    big = [torch.empty(int(free * 0.2), dtype=torch.uint8, device='cuda').fill_(0xFF) for _ in range(3)]
    small = [torch.empty(sz, dtype=torch.uint8, device='cuda').fill_(0xFF) for sz in (512, 4096, 65536, 524288) for _ in range(512)]
    torch.cuda.synchronize()
    del big, small


@pytest.fixture(scope='session', autouse=True)
def _poison():
    if os.environ.get('SMA_POISON', '0') == '1' and torch.cuda.is_available():
        _poison_device_memory()
    yield


@pytest.fixture(scope='session')
def inventory():
    return json.load(open(os.path.join(GOLD, 'state_keys.json')))


@pytest.fixture(scope='session')
def golden():
    return torch.load(os.path.join(GOLD, 'reference_clip3.pt'))


@pytest.fixture(scope='session')
def weights(inventory):
    import sma_oracle as O
    return O.synthetic_state_dict(inventory['net_g'], seed=0), O.synthetic_state_dict(inventory['motion_estimator'], seed=1)


@pytest.fixture(scope='session')
def clip():
    import sma_oracle as O
    return O.synthetic_frames(3, seed=1234)


@pytest.fixture(scope='session')
def nets(weights):
    """The B200 networks with the synthetic weights, on cuda:0 (gpu tests only)."""
    import sma_b200 as S
    g = S.build_network(CFG['network_g'])
    me = S.build_network(CFG['network_motion_estimator'])
    g.load_state_dict(weights[0], strict=True)
    me.load_state_dict(weights[1], strict=True)
    return g.eval().cuda(), me.eval().cuda()
