"""The arch plug-in a reference maintainer drops into `basicsr/archs/`.

Any file named `basicsr/archs/*_arch.py` is auto-imported by the reference (basicsr/archs/__init__.py:13-16) and its
`@ARCH_REGISTRY.register()` classes become buildable through `build_network(opt)` (:19-25).  The reference registry asserts
on duplicate class names (basicsr/utils/registry.py:38-41), so the B200 classes are registered under distinct names and the
yml switches only the `type:` strings (options/test.yml:8,47):

    network_g:                 {type: AppMotionCompFormerB200, ...every other key unchanged...}
    network_motion_estimator:  {type: Motion_Estimator_keypoint_awareB200, ...}

Exercised against the live reference registry by tests/test_dropin_cpu.py.
"""
import sma_b200 as S
from basicsr.utils.registry import ARCH_REGISTRY


@ARCH_REGISTRY.register()
class AppMotionCompFormerB200(S.AppMotionCompFormer):                 # same kwargs, same state_dict keys
    pass


@ARCH_REGISTRY.register()
class Motion_Estimator_keypoint_awareB200(S.Motion_Estimator_keypoint_aware):
    pass
