"""Parameter container that reproduces the reference state_dict key names without mirroring its module
classes: dotted names are materialised as nested anonymous nn.Module containers, so
`load_state_dict(reference_checkpoint, strict=True)` works while the forward is a
kernel launch plan over the flat name -> tensor map."""
import math
from typing import Callable, Dict, Optional, Sequence

import torch
from torch import nn


class ParamModule(nn.Module):
    def __init__(self):
        super().__init__()
        self._packed: Optional[dict] = None
        self._pack_cache_dir: Optional[str] = None
        self._pack_saved = False
        self._state_hash: Optional[str] = None

    # ---- declaration -------------------------------------------------------------------------
    def _leaf_parent(self, dotted: str):
        parts = dotted.split('.')
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        return mod, parts[-1]

    def declare(self, name: str, shape: Sequence[int], init: Callable[[torch.Tensor], None], buffer: bool = False,
                dtype=torch.float32):
        t = torch.empty(tuple(shape), dtype=dtype)
        with torch.no_grad():
            init(t)
        parent, leaf = self._leaf_parent(name)
        if buffer:
            parent.register_buffer(leaf, t)
        else:
            parent.register_parameter(leaf, nn.Parameter(t))

    def declare_conv(self, name: str, cout: int, cin: int, kh: int, kw: Optional[int] = None):
        kw = kh if kw is None else kw
        bound = 1.0 / math.sqrt(cin * kh * kw)
        self.declare(name + '.weight', (cout, cin, kh, kw), lambda t: t.uniform_(-bound, bound))
        self.declare(name + '.bias', (cout,), lambda t: t.uniform_(-bound, bound))

    def declare_linear(self, name: str, cout: int, cin: int):
        bound = 1.0 / math.sqrt(cin)
        self.declare(name + '.weight', (cout, cin), lambda t: t.uniform_(-bound, bound))
        self.declare(name + '.bias', (cout,), lambda t: t.uniform_(-bound, bound))

    def declare_norm(self, name: str, c: int):
        self.declare(name + '.weight', (c,), lambda t: t.fill_(1.0))
        self.declare(name + '.bias', (c,), lambda t: t.zero_())

    def declare_bn(self, name: str, c: int):
        self.declare_norm(name, c)
        self.declare(name + '.running_mean', (c,), lambda t: t.zero_(), buffer=True)
        self.declare(name + '.running_var', (c,), lambda t: t.fill_(1.0), buffer=True)
        self.declare(name + '.num_batches_tracked', (), lambda t: t.zero_(), buffer=True, dtype=torch.long)

    # ---- access ------------------------------------------------------------------------------
    def tensors(self) -> Dict[str, torch.Tensor]:
        d = {k: v.detach() for k, v in self.named_parameters()}
        d.update({k: v for k, v in self.named_buffers()})
        return d

    # ---- packed-weight cache invalidation ----------------------------------------------------
    def invalidate_cache(self):
        self._packed = None
        self._state_hash = None

    def _apply(self, fn, *a, **k):
        self.invalidate_cache()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate_cache()
        return super().load_state_dict(*a, **k)

    # ---- persisted pre-packed blob (packcache.py) --------------------------------
    def _restore_packed(self):
        """The packed dictionary of these exact weights from the pack cache, or None (cache disabled / miss)."""
        from .. import packcache
        d = packcache.cache_dir_of(self)
        if d is None:
            return None
        return packcache.load(self, d, next(self.parameters()).device)

    def save_pack_cache(self, force: bool = False):
        """Persist the packed dictionary (base layouts + every tensor-core image packed so far) if a cache directory is configured and
        something was packed since the last save.  Returns the file path or None."""
        from .. import packcache
        d = packcache.cache_dir_of(self)
        if d is None or self._packed is None:
            return None
        if not force and not packcache.is_dirty(self._packed) and getattr(self, '_pack_saved', False):
            return None
        path = packcache.save(self, self._packed, d)
        self._pack_saved = True
        return path
