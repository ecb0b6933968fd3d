"""Checkpoint ingest and the persisted pre-packed weight blob.

`load_network` has the semantics of the reference loader (basicsr/demo.py:46-72, the same logic as
basicsr/models/base_model.py:236-262): `torch.load` on the CPU, `param_key` selection with the `params_ema` -> `params`
fallback, `module.` prefix stripping, `load_state_dict(strict)`.

Packing (BatchNorm folding, kernel layouts, fp16 / tf32 hi-lo tensor-core images, codebook K/V projections) happens once per
weight load; tensor-core images are packed lazily, the first time a launch needs them (ops.ConvW.image).  With a cache directory
configured (`enable_pack_cache(module, dir)`, `load_network(..., pack_cache=dir)` or the SMA_B200_PACK_CACHE environment variable)
the packed dictionary is persisted, keyed by a hash of the state dict + ABI version + packing policy, and the next load of the same
weights restores it without launching a single pack kernel.
"""
import hashlib
import os
from copy import deepcopy
from typing import Dict, Optional

import torch

from . import _lib, ops

ENV = 'SMA_B200_PACK_CACHE'


def load_network(net, load_path: str, strict: bool = True, param_key: Optional[str] = 'params', pack_cache: Optional[str] = None):
    """basicsr/demo.py:46-72.  Works for any nn.Module; for the B200 arch classes `pack_cache` additionally enables the pre-pack cache."""
    print(f'Loading {net.__class__.__name__} model from {load_path}.')
    load_net = torch.load(load_path, map_location=lambda storage, loc: storage)
    if param_key is not None:
        if param_key not in load_net and 'params' in load_net:
            param_key = 'params'
            print('Loading: params_ema does not exist, use params.')
        load_net = load_net[param_key]
    # remove unnecessary 'module.'
    for k, v in deepcopy(load_net).items():
        if k.startswith('module.'):
            load_net[k[7:]] = v
            load_net.pop(k)
    net.load_state_dict(load_net, strict=strict)
    if pack_cache is not None:
        enable_pack_cache(net, pack_cache)
    return net


def pack_modules(net):
    """The ParamModule instances (the units that own a packed-weight dictionary) inside `net`."""
    from .archs.params import ParamModule
    return [m for m in net.modules() if isinstance(m, ParamModule)]


def enable_pack_cache(net, directory: Optional[str]):
    for m in pack_modules(net):
        m._pack_cache_dir = directory


def cache_dir_of(module) -> Optional[str]:
    d = getattr(module, '_pack_cache_dir', None)
    return d if d is not None else os.environ.get(ENV)


def state_hash(module) -> str:
    """blake2b over (class, ABI version, packing policy, every state tensor's name / shape / bytes)."""
    h = hashlib.blake2b(digest_size=16)
    policy = (module.__class__.__name__, _lib.load().sma_abi_version(), ops.USE_F16, ops.USE_TS, sorted(ops.FAST_STAGES), sorted(ops.X2_STAGES))
    h.update(repr(policy).encode())
    for k, v in sorted(module.state_dict().items()):
        t = v.detach().cpu().contiguous()
        h.update(k.encode()); h.update(repr((tuple(t.shape), str(t.dtype))).encode())
        h.update(t.reshape(-1).view(torch.uint8).numpy() if t.numel() else b'')
    return h.hexdigest()


def _convw_to_blob(cw: ops.ConvW) -> dict:
    return {'__convw__': True, 'w': cw.w.detach().cpu(), 'bias': None if cw.bias is None else cw.bias.detach().cpu(),
            'dims': (cw.Cout, cw.Cin, cw.kh, cw.kw), 'pack_dims': cw.pack_dims,
            'images': {k: (None if t is None else t.detach().cpu()) for k, t in (cw.images or {}).items()},
            'plans': dict(cw.plans or {}),
            'slices': {k: {'images': {ik: (None if t is None else t.detach().cpu()) for ik, t in (c.images or {}).items()}, 'plans': dict(c.plans or {})}
                       for k, c in (cw._slices or {}).items()}}


def _convw_from_blob(b: dict, dev) -> ops.ConvW:
    w = b['w'].to(dev)
    Cout, Cin, kh, kw = b['dims']
    bias = None if b['bias'] is None else b['bias'].to(dev)
    cw = ops.ConvW(w, bias, Cout, Cin, kh, kw, pack_dims=b.get('pack_dims'))
    cw.images = {k: (None if t is None else t.to(dev)) for k, t in b['images'].items()}
    cw.plans = dict(b['plans'])
    for (start, n), sb in b['slices'].items():
        c = cw.cols(start, n)
        c.images = {k: (None if t is None else t.to(dev)) for k, t in sb['images'].items()}
        c.plans = dict(sb['plans'])
    return cw


def is_dirty(W: Dict[str, object]) -> bool:
    for v in W.values():
        if isinstance(v, ops.ConvW):
            if v.dirty or any(c.dirty for c in (v._slices or {}).values()):
                return True
    return False


def _clear_dirty(W):
    for v in W.values():
        if isinstance(v, ops.ConvW):
            v.dirty = False
            for c in (v._slices or {}).values():
                c.dirty = False


def path_for(module, directory: str) -> str:
    h = getattr(module, '_state_hash', None)            # cleared by ParamModule.invalidate_cache (load_state_dict / .to())
    if h is None:
        h = module._state_hash = state_hash(module)
    return os.path.join(directory, f'{module.__class__.__name__}-{h}.smapack')


def save(module, W: Dict[str, object], directory: str) -> str:
    os.makedirs(directory, exist_ok=True)
    blob = {}
    for k, v in W.items():
        if isinstance(v, ops.ConvW):
            # a ConvW that shares its base layout with an earlier entry (as_patch views) is stored once; slices are stored inside their parent
            blob[k] = _convw_to_blob(v)
        elif isinstance(v, torch.Tensor):
            blob[k] = v.detach().cpu()
        else:
            blob[k] = v
    path = path_for(module, directory)
    tmp = path + f'.tmp{os.getpid()}'
    torch.save(blob, tmp)
    os.replace(tmp, path)           # atomic: concurrent ranks write the same bytes
    _clear_dirty(W)
    return path


def load(module, directory: str, dev) -> Optional[Dict[str, object]]:
    path = path_for(module, directory)
    if not os.path.exists(path):
        return None
    blob = torch.load(path, map_location='cpu', weights_only=False)
    W: Dict[str, object] = {}
    for k, v in blob.items():
        if isinstance(v, dict) and v.get('__convw__'):
            W[k] = _convw_from_blob(v, dev)
        elif isinstance(v, torch.Tensor):
            W[k] = v.to(dev)
        else:
            W[k] = v
    return W
