"""B200-native per-driving-frame talking-head generator (drop-in for the hot path of
ShaelynZ/synergize-motion-appearance).  The directory name contains '-', so import it through
`importlib.import_module('synergize-motion-appearance_b200')` or the `sma_b200` alias at the repo root."""
from . import _lib, ops, dist, packcache  # noqa: F401
from .registry import ARCH_REGISTRY, build_network  # noqa: F401
from .archs import motion_estimator_arch, appmotioncodebook_arch  # noqa: F401
from .archs.motion_estimator_arch import Motion_Estimator_keypoint_aware, KPDetector, DenseMotionNetwork  # noqa: F401
from .archs.appmotioncodebook_arch import AppMotionCompFormer  # noqa: F401
from .animate import make_animation, make_animation_model, make_animation_multi, normalize_kp, ClipAnimator  # noqa: F401
from .dist import make_animation_sharded  # noqa: F401
from .packcache import load_network, enable_pack_cache  # noqa: F401

__all__ = ['ARCH_REGISTRY', 'build_network', 'Motion_Estimator_keypoint_aware', 'KPDetector', 'DenseMotionNetwork',
           'AppMotionCompFormer', 'make_animation', 'make_animation_model', 'make_animation_multi', 'make_animation_sharded', 'normalize_kp',
           'ClipAnimator', 'load_network', 'enable_pack_cache', 'ops', 'dist', 'packcache']
