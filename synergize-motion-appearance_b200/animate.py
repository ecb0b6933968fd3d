"""make_animation on the B200 path.

Drop-in for `demo.make_animation` (basicsr/demo.py:103-134) and its method twin
`AppMotionCompModel.make_animation` (basicsr/models/appmotioncomp_model.py:607-639): same arguments,
same returns (lists of HWC uint8 frames).  What changes is the execution plan:
  * driving frames are processed `batch` at a time instead of one by one;
  * per-clip constants are hoisted: kp_source, kp_driving_initial, the convex-hull movement scale and
    the source encoder features (the reference recomputes the hull with scipy and the encoder for every
    frame, demo.py:26-29 and appmotioncodebook_arch.py:549-554).  The source and the first driving frame ride in
    the key-point pass of the first micro-batch and the hull-area scale is computed on the device, so a clip
    is enqueued without a single host synchronisation;
  * frame I/O either side of the path runs on the device: frames may be given as the uint8 HWC
    images a video reader produces (what demo.py:166-185 converts on the host, one frame at a time); they are staged
    through two page-locked buffers on a copy stream (H2D of micro-batch i+1 overlaps the compute of micro-batch i;
    a quarter of the PCIe bytes of fp32 frames) and converted by `sma_u8hwc_to_f32nchw`; the uint8 conversion of the
    result (tensor2img) runs on the device and only the 196,608-byte uint8 frame crosses PCIe, on a third stream.
"""
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

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
    """Host version of the per-clip scale (one .cpu() round trip); the animation path uses ops.hull_scale instead."""
    if kp_source['value'].shape[0] != 1 or kp_driving_initial['value'].shape[0] != 1:
        raise ValueError('movement_scale: kp_source and kp_driving_initial are per-clip constants (batch 1)')
    a = hull_area(kp_source['value'][0].detach().cpu().numpy())
    b = hull_area(kp_driving_initial['value'][0].detach().cpu().numpy())
    return float(np.sqrt(a) / np.sqrt(b))


def normalize_kp(kp_source, kp_driving, kp_driving_initial, adapt_movement_scale=False, use_relative_movement=False,
                 use_relative_jacobian=False, adjust_shape_movement=False, _scale=None):
    """demo.py:24-44 on the device.  `_scale` lets callers hoist the per-clip hull-area ratio (python float or device scalar)."""
    if use_relative_movement and not use_relative_jacobian:
        raise NotImplementedError('relative movement with absolute jacobians is not used by the reference callers')
    if not use_relative_movement:
        return dict(kp_driving)
    if kp_source['value'].shape[0] != 1 or kp_driving_initial['value'].shape[0] != 1:
        raise ValueError('normalize_kp: kp_source and kp_driving_initial are per-clip constants (batch 1)')
    if adapt_movement_scale:
        scale = ops.hull_scale(kp_source['value'].contiguous(), kp_driving_initial['value'].contiguous()) if _scale is None else _scale
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
    batched per-frame step.  `make_animation` is a thin loop over `step`.

    `driving_initial=None`: the first frame of the first `step` is the initial driving frame (what demo.make_animation does:
    kp_driving_initial = kp_detector(driving_video[0])).  Either way the source (and an explicit initial frame) ride in the key-point
    pass of the first `step` (one batch of B+1 / B+2 frames instead of a separate batch-2 pass through the weight-streaming-bound
    hourglass bottleneck)."""

    def __init__(self, net_g, motion_estimator, source: torch.Tensor, driving_initial: Optional[torch.Tensor] = None, relative=True,
                 adapt_movement_scale=True, w: float = 1.0):
        if source.dim() != 4 or source.shape[0] != 1:
            raise ValueError('ClipAnimator: source must be (1,3,H,W); use make_animation_multi for several identities')
        self.net_g, self.me = net_g, motion_estimator
        self.relative, self.w = relative, float(w)
        self.adapt = adapt_movement_scale
        self.source = source.contiguous().float()
        self.kp_source = self.kp_initial = self.scale = None
        self._initial = None if driving_initial is None else driving_initial.contiguous().float()
        with torch.no_grad():
            self.feats = self.net_g.encode_source(self.source)
            self.me.dense_motion_network.source_down(self.source)

    def _set_clip_kp(self, kp_source, kp_initial):
        self.kp_source, self.kp_initial = kp_source, kp_initial
        self.scale = (ops.hull_scale(kp_source['value'], kp_initial['value']) if (self.adapt and self.relative) else 1.0)

    @torch.no_grad()
    def clip_keypoints(self):
        """(kp_source, kp_initial, movement scale), computing them now if no `step` has run yet (needs an explicit initial frame)."""
        if self.kp_source is None:
            if self._initial is None:
                raise RuntimeError('clip_keypoints() before the first step needs driving_initial')
            kp = self.me.estimate_kp(torch.cat([self.source, self._initial], dim=0))
            self._set_clip_kp({k: v[0:1].contiguous() for k, v in kp.items()}, {k: v[1:2].contiguous() for k, v in kp.items()})
            self._initial = None
        return self.kp_source, self.kp_initial, self.scale

    @torch.no_grad()
    def detect(self, frames: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Key-points of a micro-batch (source-independent: cross-reenactment batches compute them once for all identities)."""
        return self.me.estimate_kp(frames)

    @torch.no_grad()
    def step(self, frames: torch.Tensor, bgr: bool = False, want_fp32: bool = False, kp_driving: Optional[dict] = None):
        """frames (B,3,H,W) device fp32 in [-1,1] -> (B,H,W,3) uint8 on the device (+ NHWC fp32 if asked).
        `kp_driving`: precomputed `detect(frames)`."""
        B = frames.shape[0]
        if self.kp_source is None:
            extra = [self.source] + ([self._initial] if self._initial is not None else [])
            ne = len(extra)
            if kp_driving is None:
                kp = self.me.estimate_kp(torch.cat(extra + [frames], dim=0))
                kp_driving = {k: v[ne:] for k, v in kp.items()}
                self._set_clip_kp({k: v[0:1].contiguous() for k, v in kp.items()}, {k: v[1:2].contiguous() for k, v in kp.items()})
            else:
                kp = self.me.estimate_kp(torch.cat(extra, dim=0))
                self._set_clip_kp({k: v[0:1].contiguous() for k, v in kp.items()},
                                  {k: (v[1:2] if ne == 2 else kp_driving[k][0:1]).contiguous() for k, v in kp.items()})
            self._initial = None
        kp_d = self.me.estimate_kp(frames) if kp_driving is None else kp_driving
        kp_n = normalize_kp(self.kp_source, kp_d, self.kp_initial, adapt_movement_scale=self.adapt,
                            use_relative_movement=self.relative, use_relative_jacobian=self.relative, _scale=self.scale)
        dm = self.me.estimate_motion_w_kp(kp_source=self.kp_source, kp_driving=kp_n, source_image=self.source)
        r = self.net_g.generate(self.feats, dm['deformation'], dm['occlusion_map'].view(B, *dm['deformation'].shape[1:3]),
                                dm['_driving_kp_heatmap_nhwc'], self.w)
        u8 = ops.to_uint8(r['out'], bgr)
        return (u8, r['out']) if want_fp32 else u8


def _autosave_pack_cache(*nets):
    """With a pack-cache directory configured, persist images that were packed lazily during this clip (no-op otherwise)."""
    from . import packcache
    for net in nets:
        for m in packcache.pack_modules(net):
            if packcache.cache_dir_of(m) is not None:
                m.save_pack_cache()


# ---------------------------------------------------------------------------------------------------------------------
# frame I/O: page-locked staging, copy streams
# ---------------------------------------------------------------------------------------------------------------------
class _DeviceIO:
    """Per-device staging state: an upload and a download stream, two page-locked input buffers with their device twins and the events that
    order their reuse, and a pool of page-locked result buffers handed out as numpy views without a copy."""

    def __init__(self, dev: torch.device):
        self.dev = dev
        self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.slots: Dict[tuple, list] = {}
        self.results: List[list] = []            # [page-locked tensor, weakref to the numpy array handed out (or None), busy]

    def slot(self, i: int, shape, dtype):
        """(host pinned, device twin, h2d_done event, consumed event) of input slot i & 1."""
        key = (tuple(shape), dtype)
        s = self.slots.get(key)
        if s is None:
            if len(self.slots) > 4:
                self.slots.clear()
            s = self.slots[key] = [[torch.empty(tuple(shape), dtype=dtype).pin_memory(), torch.empty(tuple(shape), dtype=dtype, device=self.dev),
                                    torch.cuda.Event(), torch.cuda.Event(), False] for _ in range(2)]
        return s[i & 1]

    def result(self, shape, dtype) -> torch.Tensor:
        """A page-locked buffer nobody holds a numpy view of any more (cudaHostAlloc costs milliseconds: buffers are recycled once the arrays
        returned by an earlier call have been garbage-collected)."""
        n = int(np.prod(shape))
        for e in self.results:
            t, ref, busy = e
            if not busy and t.numel() >= n and t.dtype == dtype and (ref is None or ref() is None):
                e[1], e[2] = None, True
                return t[:n].view(*shape)
        if len(self.results) >= 8:                 # drop what is neither lent out nor being filled
            self.results = [e for e in self.results if e[2] or (e[1] is not None and e[1]() is not None)]
        t = torch.empty((n,), dtype=dtype).pin_memory()
        self.results.append([t, None, True])
        return t.view(*shape)

    def hand_out(self, view: torch.Tensor) -> np.ndarray:
        arr = view.numpy()            # every row view keeps `arr` alive (numpy collapses view chains onto it)
        for e in self.results:
            if e[0].data_ptr() == view.data_ptr():
                e[1], e[2] = weakref.ref(arr), False
        return arr


_IO: Dict[int, _DeviceIO] = {}


def _io(dev: torch.device) -> _DeviceIO:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    io = _IO.get(idx)
    if io is None:
        io = _IO[idx] = _DeviceIO(torch.device('cuda', idx))
    return io


def _is_u8_image(f) -> bool:
    return (isinstance(f, np.ndarray) and f.dtype == np.uint8 and f.ndim == 3) or \
           (isinstance(f, torch.Tensor) and f.dtype == torch.uint8 and f.dim() == 3)


def frames_to_device(frames: Sequence, dev: torch.device, swap_rb: bool = False) -> torch.Tensor:
    """A few frames (uint8 HWC images or fp32 CHW tensors in [-1,1]) -> (n,3,H,W) fp32 on `dev`, without the staging machinery."""
    if _is_u8_image(frames[0]):
        u8 = torch.stack([torch.as_tensor(np.ascontiguousarray(f) if isinstance(f, np.ndarray) else f) for f in frames]).to(dev, non_blocking=True)
        return ops.u8hwc_to_f32nchw(u8.contiguous(), swap_rb)
    return torch.stack([f.float() for f in frames]).to(dev, non_blocking=True)


class _Uploader:
    """Stages micro-batches of driving frames: host gather into page-locked slot i&1, H2D on the upload stream, device-side conversion on the
    compute stream.  Reuse of a slot waits on the event recorded after its previous H2D (host side) and on the event recorded after the
    compute stream last read its device twin (upload-stream side)."""

    PIECE = 16

    def __init__(self, io: _DeviceIO, frames: Sequence, batch: int):
        self.io, self.frames, self.batch = io, frames, batch
        f0 = frames[0]
        self.u8 = _is_u8_image(f0)
        self.on_device = isinstance(f0, torch.Tensor) and f0.is_cuda
        self.shape = tuple(f0.shape)
        self.dtype = torch.uint8 if self.u8 else torch.float32
        self.i = 0

    def bytes_per_frame(self) -> int:
        return int(np.prod(self.shape)) * (1 if self.u8 else 4)

    def upload(self, i0: int):
        """Enqueue the H2D of frames[i0:i0+batch]; returns a ticket for `take`."""
        chunk = self.frames[i0:i0 + self.batch]
        n = len(chunk)
        if self.on_device:
            return ('dev', torch.stack(list(chunk)), n)
        host, devbuf, h2d_done, consumed, used = slot = self.io.slot(self.i, (self.batch,) + self.shape, self.dtype)
        self.i += 1
        if used:
            h2d_done.synchronize()                  # the previous copy out of this page-locked buffer has finished: safe to overwrite
        cur = torch.cuda.current_stream(self.io.dev)
        with torch.cuda.stream(self.io.up):
            if used:
                self.io.up.wait_event(consumed)     # the compute stream is done reading the device twin
            # the host gather (a memcpy per frame out of the reader's separate arrays, ~20 us each) runs in pieces so that the H2D of one piece
            # overlaps the gather of the next
            hn = host.numpy() if (self.u8 and isinstance(chunk[0], np.ndarray)) else None
            for a in range(0, n, self.PIECE):
                b = min(n, a + self.PIECE)
                if hn is not None:
                    for i in range(a, b):
                        hn[i] = chunk[i]
                else:
                    torch.stack([f if f.dtype == self.dtype else f.to(self.dtype) for f in chunk[a:b]], out=host[a:b])
                devbuf[a:b].copy_(host[a:b], non_blocking=True)
            h2d_done.record(self.io.up)
        slot[4] = True
        return ('slot', slot, n, cur)

    def take(self, ticket, swap_rb: bool = False):
        """-> ((n,3,H,W) fp32 frames on the compute stream, release callback to call once the step that reads them is enqueued)."""
        if ticket[0] == 'dev':
            t = ticket[1]
            return (ops.u8hwc_to_f32nchw(t.contiguous(), swap_rb) if self.u8 else t.float()), (lambda: None)
        _, slot, n, cur = ticket
        host, devbuf, h2d_done, consumed, _ = slot
        cur.wait_event(h2d_done)
        if self.u8:
            frames = ops.u8hwc_to_f32nchw(devbuf[:n], swap_rb)
            consumed.record(cur)                    # the fp32 copy is private: the twin may be overwritten right away
            return frames, (lambda: None)
        return devbuf[:n], (lambda: consumed.record(cur))


def make_animation(source_image, driving_video: Sequence, net_g, motion_estimator, relative=True,
                   adapt_movement_scale=True, cpu=False, batch: int = 64, w: float = 1.0, bgr: bool = False, source_bgr2rgb: bool = False):
    """Same signature/returns as demo.make_animation (demo.py:103-134); `cpu=True` raises (no CPU path).

    source_image: (3,H,W) fp32 in [-1,1] (the reference's input) or an (H,W,3) uint8 image (`source_bgr2rgb=True` for a cv2.imread result, as
    demo.py:180).  driving_video: list of (3,H,W) fp32 tensors or of (H,W,3) uint8 RGB frames (numpy / torch), e.g. straight from a video reader."""
    if cpu:
        raise RuntimeError('the B200 path has no CPU fallback; use the reference for cpu=True')
    dev = next(net_g.parameters()).device
    n = len(driving_video)
    with torch.no_grad(), torch.cuda.device(dev):
        io = _io(dev)
        cur = torch.cuda.current_stream(dev)
        src = frames_to_device([source_image], dev, swap_rb=source_bgr2rgb)
        anim = ClipAnimator(net_g, motion_estimator, src, None, relative, adapt_movement_scale, w)      # enqueues the source encoder
        upl = _Uploader(io, driving_video, batch)
        H, W = src.shape[2], src.shape[3]
        pred_host = io.result((n, H, W, 3), torch.uint8)
        echo_inputs = upl.u8 and not upl.on_device       # tensor2img(normalise(u8 frame)) == the frame itself: nothing to compute or copy
        drv_host = None if echo_inputs else io.result((n, H, W, 3), torch.uint8)
        ticket = upl.upload(0)
        for i0 in range(0, n, batch):
            frames, release = upl.take(ticket)
            if i0 + batch < n:
                ticket = upl.upload(i0 + batch)          # host gather + H2D of the next micro-batch overlap this one's compute
            nb = frames.shape[0]
            u8 = anim.step(frames, bgr)
            d8 = None if echo_inputs else ops.to_uint8(ops.nchw_to_nhwc(frames), bgr)
            release()
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(io.down):             # D2H off the compute stream
                io.down.wait_event(done)
                pred_host[i0:i0 + nb].copy_(u8, non_blocking=True)
                u8.record_stream(io.down)
                if d8 is not None:
                    drv_host[i0:i0 + nb].copy_(d8, non_blocking=True)
                    d8.record_stream(io.down)
        io.down.synchronize()
    _autosave_pack_cache(net_g, motion_estimator)
    p = io.hand_out(pred_host)
    if echo_inputs:
        drv = [np.asarray(f)[:, :, ::-1] if bgr else np.asarray(f) for f in driving_video]
    else:
        d = io.hand_out(drv_host)
        drv = [d[i] for i in range(n)]
    return [p[i] for i in range(n)], drv


def make_animation_multi(source_images: Sequence, driving_video: Sequence, net_g, motion_estimator, relative=True,
                         adapt_movement_scale=True, batch: int = 64, w: float = 1.0, bgr: bool = False, source_bgr2rgb: bool = False):
    """Cross-reenactment batch (BASELINE configs[4]: S source identities x T shared driving frames): `demo.make_animation` for every source
    over the same driving clip (demo.py:103-134 called once per identity).  The driving frames are uploaded once, their key-points - which do
    not depend on the source - are detected once per micro-batch and shared by all identities; per identity only the source encoder and its
    key-points are recomputed.  Returns (predictions[s][t] HWC uint8, driving frames[t] HWC uint8)."""
    dev = next(net_g.parameters()).device
    S_, n = len(source_images), len(driving_video)
    with torch.no_grad(), torch.cuda.device(dev):
        io = _io(dev)
        cur = torch.cuda.current_stream(dev)
        srcs = [frames_to_device([s], dev, swap_rb=source_bgr2rgb) for s in source_images]
        anims = [ClipAnimator(net_g, motion_estimator, s, None, relative, adapt_movement_scale, w) for s in srcs]
        # every ClipAnimator caches its own encoder features; the nets' single-entry source caches are refilled per identity below
        upl = _Uploader(io, driving_video, batch)
        H, W = srcs[0].shape[2], srcs[0].shape[3]
        pred_host = io.result((S_, n, H, W, 3), torch.uint8)
        echo_inputs = upl.u8 and not upl.on_device
        drv_host = None if echo_inputs else io.result((n, H, W, 3), torch.uint8)
        ticket = upl.upload(0)
        for i0 in range(0, n, batch):
            frames, release = upl.take(ticket)
            if i0 + batch < n:
                ticket = upl.upload(i0 + batch)
            nb = frames.shape[0]
            kp_d = anims[0].detect(frames)
            if i0 == 0:          # source key-points of all identities in one pass; the initial driving key-points are those of frame 0
                kp_s = motion_estimator.estimate_kp(torch.cat(srcs, dim=0))
                kp_0 = {k: v[0:1].contiguous() for k, v in kp_d.items()}
                for s_i, a in enumerate(anims):
                    a._set_clip_kp({k: v[s_i:s_i + 1].contiguous() for k, v in kp_s.items()}, kp_0)
            outs = [a.step(frames, bgr, kp_driving=kp_d) for a in anims]
            d8 = None if echo_inputs else ops.to_uint8(ops.nchw_to_nhwc(frames), bgr)
            release()
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(io.down):
                io.down.wait_event(done)
                for s_i, u8 in enumerate(outs):
                    pred_host[s_i, i0:i0 + nb].copy_(u8, non_blocking=True)
                    u8.record_stream(io.down)
                if d8 is not None:
                    drv_host[i0:i0 + nb].copy_(d8, non_blocking=True)
                    d8.record_stream(io.down)
        io.down.synchronize()
    _autosave_pack_cache(net_g, motion_estimator)
    p = io.hand_out(pred_host)
    if echo_inputs:
        drv = [np.asarray(f)[:, :, ::-1] if bgr else np.asarray(f) for f in driving_video]
    else:
        d = io.hand_out(drv_host)
        drv = [d[i] for i in range(n)]
    return [[p[s_i, i] for i in range(n)] for s_i in range(S_)], drv


def make_animation_model(model_opt: dict, net_g, motion_estimator, source_img, driving: List[torch.Tensor], cpu=False,
                         batch: int = 64):
    """Twin of AppMotionCompModel.make_animation (appmotioncomp_model.py:607-639): batch-1 NCHW tensors in a list,
    options read from opt['val'] ('relative', 'adapt_scale', 'w'), BGR output."""
    if source_img.dim() != 4 or source_img.shape[0] != 1:
        raise ValueError('make_animation_model: source_img must be (1,3,H,W) as in the reference')
    val = model_opt.get('val', {})
    return make_animation(source_img[0], [f[0] for f in driving], net_g, motion_estimator,
                          relative=val.get('relative', False), adapt_movement_scale=val.get('adapt_scale', False),
                          cpu=cpu, batch=batch, w=val.get('w', 1), bgr=True)
