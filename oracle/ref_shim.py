"""Import shim for the UNMODIFIED reference (container only; /root/reference is absent on the GPU box).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to (a) pin the oracle restatement
(oracle/sma_oracle.py) against the live reference modules and (b) generate tests/golden/*.
Nothing in the product path, tests -m gpu, smoke() or bench.py imports this file.

The reference star-imports every util at package import (basicsr/__init__.py:3-10) and so needs
imageio/skimage/decord/flow_vis/lpips/insightface/mediapipe; they are absent here and unused by
the hot path, so they are stubbed with MagicMock.
"""
import importlib
import sys
from unittest.mock import MagicMock

REF_ROOT = '/root/reference'


def import_reference(ref_root: str = REF_ROOT):
    sys.dont_write_bytecode = True
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    for _ in range(80):
        try:
            for k in [k for k in sys.modules if k == 'basicsr' or k.startswith('basicsr.')]:
                del sys.modules[k]
            import basicsr.archs  # noqa: F401
            break
        except ModuleNotFoundError as e:
            m = MagicMock()
            m.__name__ = e.name
            m.__path__ = []
            m.__spec__ = None
            sys.modules[e.name] = m
    else:
        raise RuntimeError('could not import the reference')
    import yaml
    from basicsr.archs import build_network
    from basicsr.utils.options import ordered_yaml
    cfg = yaml.load(open(ref_root + '/options/test.yml'), Loader=ordered_yaml()[0])
    return build_network, cfg


def load_demo_module(ref_root: str = REF_ROOT):
    for name in ('ffmpeg', 'cv2'):
        try:
            importlib.import_module(name)
        except ModuleNotFoundError:
            m = MagicMock(); m.__name__ = name; m.__path__ = []; m.__spec__ = None
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location('ref_demo', ref_root + '/basicsr/demo.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
