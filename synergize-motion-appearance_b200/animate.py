"""make_animation on the B200 path.

Drop-in for `demo.make_animation` (basicsr/demo.py:103-134) and its method twin
`AppMotionCompModel.make_animation` (basicsr/models/appmotioncomp_model.py:607-639): same arguments,
same returns (lists of HWC uint8 frames).  What changes is the execution plan:
  * driving frames are processed `batch` at a time instead of one by one;
  * per-clip constants are hoisted: kp_source, kp_driving_initial, the convex-hull movement scale and
    the source encoder features (the reference recomputes the hull with scipy and the encoder for every
    frame, demo.py:26-29 and appmotioncodebook_arch.py:549-554);
  * normalize_kp runs on the device; the uint8 conversion (tensor2img) runs on the device and only the
    196,608-byte uint8 frame crosses PCIe, asynchronously, instead of the 786,432-byte fp32 one.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


def hull_area(points: np.ndarray) -> float:
    """Area of the 2-D convex hull (== scipy.spatial.ConvexHull(points).volume used by demo.py:26-28)."""
    pts = sorted(map(tuple, np.asarray(points, dtype=np.float64).tolist()))
    if len(pts) < 3:
        return 0.0

    def turn(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    def half(seq):
        h = []
        for q in seq:
            while len(h) >= 2 and turn(h[-2], h[-1], q) <= 0:
                h.pop()
            h.append(q)
        return h[:-1]
    hull = half(pts) + half(reversed(pts))
    return 0.5 * abs(sum(hull[i][0] * hull[(i + 1) % len(hull)][1] - hull[(i + 1) % len(hull)][0] * hull[i][1]
                         for i in range(len(hull))))


def movement_scale(kp_source: dict, kp_driving_initial: dict) -> float:
    a = hull_area(kp_source['value'][0].detach().cpu().numpy())
    b = hull_area(kp_driving_initial['value'][0].detach().cpu().numpy())
    return float(np.sqrt(a) / np.sqrt(b))


def normalize_kp(kp_source, kp_driving, kp_driving_initial, adapt_movement_scale=False, use_relative_movement=False,
                 use_relative_jacobian=False, adjust_shape_movement=False, _scale: Optional[float] = None):
    """demo.py:24-44 on the device.  `_scale` lets callers hoist the per-clip hull-area ratio."""
    if use_relative_movement and not use_relative_jacobian:
        raise NotImplementedError('relative movement with absolute jacobians is not used by the reference callers')
    if not use_relative_movement:
        return dict(kp_driving)
    if adapt_movement_scale:
        scale = movement_scale(kp_source, kp_driving_initial) if _scale is None else _scale
    else:
        scale = 1.0
    v, j = ops.normalize_kp(kp_source['value'].contiguous(), kp_source['jacobian'].contiguous(),
                            kp_driving['value'].contiguous(), kp_driving['jacobian'].contiguous(),
                            kp_driving_initial['value'].contiguous(), kp_driving_initial['jacobian'].contiguous(), scale, True)
    out = dict(kp_driving)
    out['value'], out['jacobian'] = v, j
    return out


class ClipAnimator:
    """Per-clip state (source features, source keypoints, initial driving keypoints, movement scale) plus the
    batched per-frame step.  `make_animation` is a thin loop over `step`."""

    def __init__(self, net_g, motion_estimator, source: torch.Tensor, driving_initial: torch.Tensor, relative=True,
                 adapt_movement_scale=True, w: float = 1.0):
        self.net_g, self.me = net_g, motion_estimator
        self.relative, self.w = relative, float(w)
        self.source = source.contiguous().float()
        with torch.no_grad():
            # one key-point detector pass for the source and the first driving frame (the tiny hourglass bottleneck layers are
            # weight-streaming bound at batch 1: batching the two per-clip calls halves that cost)
            kp2 = self.me.estimate_kp(torch.cat([self.source, driving_initial.contiguous().float()], dim=0))
            self.kp_source = {k: v[0:1].contiguous() for k, v in kp2.items()}
            self.kp_initial = {k: v[1:2].contiguous() for k, v in kp2.items()}
            self.scale = movement_scale(self.kp_source, self.kp_initial) if (adapt_movement_scale and relative) else 1.0
            self.feats = self.net_g.encode_source(self.source)
            self.me.dense_motion_network.source_down(self.source)
        self.adapt = adapt_movement_scale

    @torch.no_grad()
    def step(self, frames: torch.Tensor, bgr: bool = False, want_fp32: bool = False):
        """frames (B,3,H,W) device fp32 in [-1,1] -> (B,H,W,3) uint8 on the device (+ NHWC fp32 if asked)."""
        kp_d = self.me.estimate_kp(frames)
        kp_n = normalize_kp(self.kp_source, kp_d, self.kp_initial, adapt_movement_scale=self.adapt,
                            use_relative_movement=self.relative, use_relative_jacobian=self.relative, _scale=self.scale)
        dm = self.me.estimate_motion_w_kp(kp_source=self.kp_source, kp_driving=kp_n, source_image=self.source)
        r = self.net_g.generate(self.feats, dm['deformation'], dm['occlusion_map'].view(frames.shape[0], 64, 64),
                                dm['_driving_kp_heatmap_nhwc'], self.w)
        u8 = ops.to_uint8(r['out'], bgr)
        return (u8, r['out']) if want_fp32 else u8


_PINNED = {}


def _pinned(tag: str, shape, dtype) -> torch.Tensor:
    """Page-locked staging buffers are kept across calls (cudaHostAlloc costs milliseconds; a clip needs three of them)."""
    key = (tag, tuple(shape), dtype)
    t = _PINNED.get(key)
    if t is None:
        if len(_PINNED) > 16:
            _PINNED.clear()
        t = _PINNED[key] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
    return t


def make_animation(source_image, driving_video: Sequence[torch.Tensor], net_g, motion_estimator, relative=True,
                   adapt_movement_scale=True, cpu=False, batch: int = 64, w: float = 1.0, bgr: bool = False):
    """Same signature/returns as demo.make_animation (demo.py:103-134); `cpu=True` raises (no CPU path)."""
    if cpu:
        raise RuntimeError('the B200 path has no CPU fallback; use the reference for cpu=True')
    dev = next(net_g.parameters()).device
    with torch.no_grad():
        src = source_image.unsqueeze(0).to(dev, non_blocking=True)
        first = driving_video[0].unsqueeze(0).to(dev, non_blocking=True)
        anim = ClipAnimator(net_g, motion_estimator, src, first, relative, adapt_movement_scale, w)
        n = len(driving_video)
        pred_host = _pinned('pred', (n, src.shape[2], src.shape[3], 3), torch.uint8)
        drv_host = _pinned('drv', (n, src.shape[2], src.shape[3], 3), torch.uint8)
        for i0 in range(0, n, batch):
            chunk = driving_video[i0:i0 + batch]
            if chunk[0].is_cuda:
                host = torch.stack(list(chunk)).float()
            else:                       # gather the frames straight into page-locked memory (one pass over the data)
                host = _pinned('in%d' % ((i0 // batch) & 1), (len(chunk),) + tuple(chunk[0].shape), torch.float32)
                torch.stack([f.float() for f in chunk], out=host)
            frames = host.to(dev, non_blocking=True)
            u8 = anim.step(frames, bgr)
            pred_host[i0:i0 + len(chunk)].copy_(u8, non_blocking=True)
            drv_host[i0:i0 + len(chunk)].copy_(ops.to_uint8(ops.nchw_to_nhwc(frames), bgr), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    p, d = pred_host.numpy().copy(), drv_host.numpy().copy()        # the staging buffers are reused by the next call
    return [p[i] for i in range(n)], [d[i] for i in range(n)]


def make_animation_model(model_opt: dict, net_g, motion_estimator, source_img, driving: List[torch.Tensor], cpu=False,
                         batch: int = 64):
    """Twin of AppMotionCompModel.make_animation (appmotioncomp_model.py:607-639): batch-1 NCHW tensors in a list,
    options read from opt['val'] ('relative', 'adapt_scale', 'w'), BGR output."""
    val = model_opt.get('val', {})
    return make_animation(source_img[0], [f[0] for f in driving], net_g, motion_estimator,
                          relative=val.get('relative', False), adapt_movement_scale=val.get('adapt_scale', False),
                          cpu=cpu, batch=batch, w=val.get('w', 1), bgr=True)
