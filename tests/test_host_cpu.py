"""CPU: host-side logic - registry semantics, parameter inventory, C-ABI exports, loud failure without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import sma_b200 as S
from conftest import CFG, ROOT


def test_registry_semantics():
    from importlib import import_module
    reg = import_module('synergize-motion-appearance_b200.registry')
    r = reg.Registry('t')

    @r.register()
    class A:
        pass
    assert r.get('A') is A and 'A' in r
    with pytest.raises(AssertionError):
        r.register(A)                      # duplicate names assert (basicsr/utils/registry.py:38-41)
    with pytest.raises(KeyError):
        r.get('missing')                   # (:62-66)
    for name in ('AppMotionCompFormer', 'Motion_Estimator_keypoint_aware', 'KPDetector', 'DenseMotionNetwork'):
        assert name in S.ARCH_REGISTRY
    opt = dict(CFG['network_motion_estimator'])
    net = S.build_network(opt)
    assert 'type' in opt                   # build_network deep-copies (basicsr/archs/__init__.py:20)
    assert type(net).__name__ == 'Motion_Estimator_keypoint_aware'
    assert hasattr(net, 'kp_detector') and hasattr(net, 'dense_motion_network')


def test_state_dict_inventory_and_strict_load(inventory, weights):
    g = S.build_network(CFG['network_g'])
    me = S.build_network(CFG['network_motion_estimator'])
    for net, ref, w in ((g, inventory['net_g'], weights[0]), (me, inventory['motion_estimator'], weights[1])):
        sd = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert sd == {k: list(v) for k, v in ref.items()}
        net.load_state_dict(w, strict=True)
        with pytest.raises(RuntimeError):
            bad = dict(w); bad.pop(next(iter(bad)))
            net.load_state_dict(bad, strict=True)
    # reference init conventions (fresh, unloaded networks)
    g = S.build_network(CFG['network_g'])
    me = S.build_network(CFG['network_motion_estimator'])
    assert float(g.state_dict()['position_emb_app'].abs().max()) == 0.0
    assert float(g.state_dict()['quantize_app.embedding.weight'].abs().max()) <= 1.0 / 1024
    assert float(me.state_dict()['kp_detector.jacobian.weight'].abs().max()) == 0.0
    assert torch.equal(me.state_dict()['kp_detector.jacobian.bias'][:4], torch.tensor([1., 0., 0., 1.]))
    assert torch.allclose(me.state_dict()['kp_detector.down.weight'], weights[1]['kp_detector.down.weight'], atol=1e-8)


def test_unsupported_configs_raise():
    for key, val in (('img_size', 384), ('nf', 32), ('split', 2), ('codebook_size_app', 1000)):
        bad = dict(CFG['network_g']); bad[key] = val
        with pytest.raises(NotImplementedError):
            S.build_network(bad)


def test_512_variant_inventory(inventory):
    """BASELINE configs[3] (Config 4): the 512x512 variant has the reference's key inventory and shapes except the position
    embeddings (one row per token of the 64x64 grid)."""
    import sma_oracle as O
    opt = dict(CFG['network_g']); opt['img_size'] = 512
    g = S.build_network(opt)
    sd = {k: list(v.shape) for k, v in g.state_dict().items()}
    assert sd == O.variant_shapes(inventory['net_g'], 512)
    assert sd['position_emb_app'] == [4096, 256] and sd['position_emb_motion'] == [4096, 32]
    assert (g.tg, g.fg, g.L, g.R) == (64, 128, 4096, 2)
    g.load_state_dict(O.synthetic_state_dict(O.variant_shapes(inventory['net_g'], 512), 0), strict=True)


def test_c_abi_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, 'synergize-motion-appearance_b200', 'csrc', 'libsma_b200.so')
    assert os.path.exists(lib_path), 'run __graft_entry__.build() first'
    header = open(os.path.join(ROOT, 'include', 'sma_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    names = sorted(set(re.findall(r'\b(sma_[a-z0-9_]+)\s*\(', header)))
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), n
    lib.sma_abi_version.restype = ctypes.c_int
    assert lib.sma_abi_version() == 11
    from importlib import import_module
    _lib = import_module('synergize-motion-appearance_b200._lib')
    assert sorted(_lib.SIGNATURES) == names       # the ctypes table binds exactly the declared ABI
    bound = _lib.load()
    assert bound.sma_sizeof_conv_desc() == ctypes.sizeof(_lib.ConvDesc)      # the struct mirror matches the compiled layout


def test_library_sass_carries_tcgen05_tmem_and_tma_instructions():
    """The built sm_100a library really is the Blackwell path: tcgen05 MMAs (UTCHMMA) with their commit barriers (UTCBAR), tensor-memory loads / stores
    (LDTM / STTM), tensor-map TMA loads (UTMALDG) and bulk copies (UBLKCP) in the SASS; no GPU needed (cuobjdump disassembles the cubin)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    lib_path = os.path.join(ROOT, 'synergize-motion-appearance_b200', 'csrc', 'libsma_b200.so')
    r = subprocess.run([cuobjdump, '-sass', lib_path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    assert 'sm_100a' in r.stdout
    count = {k: r.stdout.count(k + ' ') + r.stdout.count(k + '.') for k in ('UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP')}
    assert count['UTCHMMA'] > 1000 and count['UTCBAR'] > 100 and count['LDTM'] > 50 and count['STTM'] > 50 and count['UTMALDG'] > 10 and count['UBLKCP'] > 10, count
    for kernel in ('conv_tc2_kernel', 'attn_mh_kernel', 'attn256_kernel', 'warp_occlude_kernel', 'vq_lookup_tiled_kernel', 'tapsum3_tiled_kernel', 'layernorm32_kernel'):
        assert kernel in r.stdout, kernel


def test_cpu_tensors_fail_loudly(weights):
    me = S.build_network(CFG['network_motion_estimator'])
    me.load_state_dict(weights[1])
    with pytest.raises(RuntimeError):
        me.estimate_kp(torch.zeros(1, 3, 256, 256))


def test_hull_area_and_shard_ranges():
    from scipy.spatial import ConvexHull
    from importlib import import_module
    an = import_module('synergize-motion-appearance_b200.animate')
    di = import_module('synergize-motion-appearance_b200.dist')
    rng = np.random.default_rng(1)
    for _ in range(10):
        p = rng.normal(size=(15, 2))
        assert abs(an.hull_area(p) - ConvexHull(p).volume) < 1e-9
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [di.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert sum(h - l for l, h in spans) == n


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract keys; no GPU needed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--ref-frames', '2'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['unit'] == 'frames/s' and line['higher_is_better'] is True and line['value'] > 0
    assert line['metric'] == '256x256 frames/sec' and 'workload' in line['config']
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
