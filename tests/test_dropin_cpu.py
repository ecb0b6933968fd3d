"""CPU: the drop-in proven against the LIVE reference registry (container only: /root/reference is absent on the GPU box, where this
module skips), checkpoint ingest with the reference loader's semantics, and the host-side frame conversions."""
import copy
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import sma_b200 as S
import sma_oracle as O
from conftest import CFG, ROOT

REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_plugin_file_registers_in_the_reference_registry_and_strict_loads_reference_state():
    """integration/sma_b200_arch.py is what a maintainer drops into basicsr/archs/ (auto-imported, basicsr/archs/__init__.py:13-16).
    Here: import the unmodified reference, execute the plug-in file against the reference's OWN ARCH_REGISTRY, build both networks with
    the reference's OWN build_network from options/test.yml with only `type:` switched, and strict-load the state_dict() of the
    reference's own modules.  Construction + load only (no forward: no GPU here)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_shim
    build_network, cfg = ref_shim.import_reference()
    from basicsr.utils.registry import ARCH_REGISTRY as REF_REG
    assert 'AppMotionCompFormer' in REF_REG._obj_map                       # the reference's own class is registered ...
    if 'AppMotionCompFormerB200' not in REF_REG._obj_map:
        spec = importlib.util.spec_from_file_location('basicsr.archs.sma_b200_arch', os.path.join(ROOT, 'integration', 'sma_b200_arch.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    assert REF_REG.get('AppMotionCompFormerB200').__mro__[1] is S.AppMotionCompFormer      # ... next to ours, under a distinct name
    with pytest.raises(AssertionError):                                     # duplicate names assert (basicsr/utils/registry.py:38-41)
        REF_REG.register(REF_REG.get('AppMotionCompFormerB200'))
    torch.manual_seed(0)
    ref_g = build_network(copy.deepcopy(cfg['network_g'])).eval()
    ref_me = build_network(copy.deepcopy(cfg['network_motion_estimator'])).eval()
    opt_g = copy.deepcopy(cfg['network_g']); opt_g['type'] = 'AppMotionCompFormerB200'
    opt_me = copy.deepcopy(cfg['network_motion_estimator']); opt_me['type'] = 'Motion_Estimator_keypoint_awareB200'
    g = build_network(opt_g).eval()
    me = build_network(opt_me).eval()
    assert type(g).__name__ == 'AppMotionCompFormerB200' and isinstance(g, S.AppMotionCompFormer)
    assert isinstance(me, S.Motion_Estimator_keypoint_aware)
    for ours, ref in ((g, ref_g), (me, ref_me)):
        missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)
        assert not missing and not unexpected
        sd = ours.state_dict()
        for k, v in ref.state_dict().items():
            assert torch.equal(sd[k], v), k
    # the methods / attributes the reference callers touch
    for name in ('encode_driving', 'generator', 'forward'):
        assert hasattr(g, name)
    assert callable(g.generator)
    for name in ('estimate_kp', 'estimate_motion_w_kp', 'kp_detector', 'dense_motion_network'):
        assert hasattr(me, name)


def test_load_network_params_fallback_and_module_prefix(tmp_path, weights):
    """basicsr/demo.py:46-72: param_key 'params_ema' falls back to 'params' when absent; 'module.' prefixes are stripped; strict load."""
    me = S.build_network(CFG['network_motion_estimator'])
    path = str(tmp_path / 'me.pth')
    torch.save({'params': {'module.' + k: v for k, v in weights[1].items()}}, path)
    S.load_network(me, path, True, 'params_ema')
    sd = me.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in weights[1].items())
    torch.save({'params_ema': dict(weights[1]), 'params': {}}, path)      # params_ema present: it is the one loaded
    S.load_network(S.build_network(CFG['network_motion_estimator']), path, True, 'params_ema')
    bad = dict(weights[1]); bad.pop(next(iter(bad)))
    torch.save({'params': bad}, path)
    with pytest.raises(RuntimeError):
        S.load_network(S.build_network(CFG['network_motion_estimator']), path, True, 'params')
    S.load_network(S.build_network(CFG['network_motion_estimator']), path, False, 'params')           # strict=False tolerates it
    torch.save(dict(weights[1]), path)                                                            # param_key None: the file is the state dict
    S.load_network(S.build_network(CFG['network_motion_estimator']), path, True, None)


def test_pack_cache_key_changes_with_weights(tmp_path, weights):
    from importlib import import_module
    pc = import_module('synergize-motion-appearance_b200.packcache')
    me = S.build_network(CFG['network_motion_estimator'])
    me.load_state_dict(weights[1])
    S.enable_pack_cache(me, str(tmp_path))
    mods = pc.pack_modules(me)
    assert len(mods) == 2 and all(pc.cache_dir_of(m) == str(tmp_path) for m in mods)
    h0 = pc.state_hash(mods[0])
    assert h0 == pc.state_hash(mods[0])
    w2 = dict(weights[1]); k = 'kp_detector.kp.bias'; w2[k] = w2[k] + 1e-3
    me.load_state_dict(w2)
    assert pc.state_hash(mods[0]) != h0


def test_uint8_frames_round_trip_through_the_reference_preparation():
    """demo.py:177-185 (astype(float32)/255, normalize(0.5,0.5)) followed by tensor2img (utils/img_util.py:42-98) is the identity on uint8:
    make_animation therefore returns uint8 input frames themselves as `driving_imgs` instead of recomputing them."""
    u = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, axis=2)          # (1,256,3) HWC image with every level
    x = torch.from_numpy(u.astype(np.float32) / 255.).permute(2, 0, 1)
    x = (x - 0.5) / 0.5
    assert np.array_equal(O.to_uint8(x), u)
    assert np.array_equal(O.to_uint8(x, bgr=True), u[:, :, ::-1])
