"""Motion_Estimator_keypoint_aware on B200 kernels.

Drop-in for basicsr/archs/motion_estimator_arch.py:14-52 (which wraps keypoint_detector_arch.py:13-86
and dense_motion_arch.py:12-161): same constructor kwargs (options/test.yml:47-66), same state_dict
keys, same methods (`estimate_kp`, `estimate_motion_w_kp`, `forward`) and the attributes callers touch
(`kp_detector`, `dense_motion_network`).  Inference only.  Inputs/outputs keep the reference layout
(NCHW images, (B,K,2) / (B,K,2,2) keypoints, NHWC deformation, NCHW maps); internally everything is
NHWC and the forward is a launch plan over csrc/ kernels: BatchNorm(eval) is folded into the packed
conv weights, nearest-x2 upsampling into the conv operand load, skip concatenations into channel
slices of preallocated buffers, the two 7x7 heads of each network into one conv.
"""
from typing import Dict, List, Optional

import torch

from .. import ops
from ..registry import ARCH_REGISTRY
from .params import ParamModule


def _gaussian13(t: torch.Tensor):
    """AntiAliasInterpolation2d buffer for scale 0.25: sigma 1.5, 13 taps (motion_estimator_util.py:603-631)."""
    ax = torch.arange(13, dtype=torch.float32)
    g = torch.exp(-(ax - 6.0) ** 2 / (2 * 1.5 ** 2))
    k = g.view(13, 1) * g.view(1, 13)
    t.copy_((k / k.sum()).view(1, 1, 13, 13).expand_as(t))


class _HourglassPlan:
    """Channel bookkeeping of Hourglass(block_expansion, in_features, num_blocks, max_features)
    (motion_estimator_util.py:440-492)."""

    def __init__(self, e: int, cin: int, nb: int, mx: int):
        self.nb = nb
        self.enc = [(cin if i == 0 else min(mx, e * 2 ** i), min(mx, e * 2 ** (i + 1))) for i in range(nb)]
        self.dec = [((1 if i == nb - 1 else 2) * min(mx, e * 2 ** (i + 1)), min(mx, e * 2 ** i)) for i in reversed(range(nb))]
        self.out_filters = e + cin
        self.cin = cin

    def declare(self, m: ParamModule, prefix: str):
        for i, (ci, co) in enumerate(self.enc):
            m.declare_conv(f'{prefix}.encoder.down_blocks.{i}.conv', co, ci, 3)
            m.declare_bn(f'{prefix}.encoder.down_blocks.{i}.norm', co)
        for i, (ci, co) in enumerate(self.dec):
            m.declare_conv(f'{prefix}.decoder.up_blocks.{i}.conv', co, ci, 3)
            m.declare_bn(f'{prefix}.decoder.up_blocks.{i}.norm', co)

    def pack(self, T: Dict[str, torch.Tensor], prefix: str) -> dict:
        out = {}
        for kind, n in (('encoder.down_blocks', len(self.enc)), ('decoder.up_blocks', len(self.dec))):
            for i in range(n):
                p = f'{prefix}.{kind}.{i}'
                bn = {k: T[f'{p}.norm.{k}'] for k in ('weight', 'bias', 'running_mean', 'running_var')}
                out[p] = ops.pack_conv(T[p + '.conv.weight'], T[p + '.conv.bias'], bn)
        return out

    def run(self, W: dict, prefix: str, cat0: torch.Tensor, fast: bool = False) -> torch.Tensor:
        """cat0: (B,S,S,e+cin) buffer whose last `cin` channels already hold the hourglass input.
        Returns cat0 with the first `e` channels filled by the last up-block (== the reference's final
        torch.cat([out, skip]))."""
        B, S = cat0.shape[0], cat0.shape[1]
        dev = cat0.device
        # concat buffers per level: [decoder out | encoder skip]
        cats: List[torch.Tensor] = [cat0]
        for lvl in range(1, self.nb):
            co_dec = self.dec[self.nb - 1 - lvl][1]
            c_skip = self.enc[lvl - 1][1]
            s = S >> lvl
            cats.append(torch.empty((B, s, s, co_dec + c_skip), device=dev, dtype=torch.float32))
        x = cat0[..., cat0.shape[-1] - self.cin:]
        for i, (ci, co) in enumerate(self.enc):
            s = S >> i
            y = ops.conv2d(x, W[f'{prefix}.encoder.down_blocks.{i}'], pad=1, act='relu', fast=fast)
            if i + 1 < self.nb:
                dst = cats[i + 1][..., cats[i + 1].shape[-1] - co:]
            else:
                dst = torch.empty((B, s // 2, s // 2, co), device=dev, dtype=torch.float32)
            x = ops.avgpool2(y, out=dst)
        for j, (ci, co) in enumerate(self.dec):
            lvl = self.nb - 1 - j
            dst = cats[lvl][..., :co]
            ops.conv2d(x, W[f'{prefix}.decoder.up_blocks.{j}'], pad=1, act='relu', upsample2=True, out=dst, fast=fast)
            x = cats[lvl]
        return x


@ARCH_REGISTRY.register()
class KPDetector(ParamModule):
    """basicsr/archs/keypoint_detector_arch.py:13-86."""

    def __init__(self, block_expansion, num_kp, num_channels, max_features, num_blocks, temperature,
                 estimate_jacobian=False, scale_factor=1, single_jacobian_map=False, pad=0, model_path=None):
        super().__init__()
        if scale_factor != 0.25 or single_jacobian_map or pad != 0 or not estimate_jacobian:
            raise NotImplementedError('B200 KPDetector implements the options/test.yml configuration '
                                      '(scale_factor 0.25, per-keypoint jacobians, pad 0)')
        self.num_kp, self.temperature, self.num_channels = num_kp, temperature, num_channels
        self.plan = _HourglassPlan(block_expansion, num_channels, num_blocks, max_features)
        self.plan.declare(self, 'predictor')
        self.declare_conv('kp', num_kp, self.plan.out_filters, 7)
        self.declare_conv('jacobian', 4 * num_kp, self.plan.out_filters, 7)
        with torch.no_grad():   # identity jacobian init, keypoint_detector_arch.py:33-34
            self._modules['jacobian'].weight.zero_()
            self._modules['jacobian'].bias.copy_(torch.tensor([1., 0., 0., 1.] * num_kp))
        self.declare('down.weight', (num_channels, 1, 13, 13), _gaussian13, buffer=True)
        if model_path is not None:
            ck = torch.load(model_path, map_location='cpu')
            self.load_state_dict({k.replace('module.', ''): v for k, v in ck['kp_detector'].items()})

    def _weights(self):
        if self._packed is None:
            T = self.tensors()
            self._cpad = (self.plan.out_filters + 31) // 32 * 32
            W = self._restore_packed()
            if W is not None:
                self._packed, self._pack_saved = W, True
                return W
            W = self.plan.pack(T, 'predictor')
            # the 35-channel hourglass output lives in a 64-channel zero-padded buffer so that the 7x7 heads run on the tensor cores
            W['heads'] = ops.pack_conv_cat([T['kp.weight'], T['jacobian.weight']], [T['kp.bias'], T['jacobian.bias']], pad_cin=self._cpad)
            W['k13'] = T['down.weight'][0, 0].contiguous()
            self._packed = W
        return self._packed

    @torch.no_grad()
    def forward(self, x, isSource=False):
        W = self._weights()
        x = x.contiguous().float()
        B, _, H, Wd = x.shape
        full = torch.zeros((B, H // 4, Wd // 4, self._cpad), device=x.device, dtype=torch.float32)
        cat0 = full[..., :self.plan.out_filters]
        ops.antialias_down4(x, W['k13'], out=cat0[..., self.plan.out_filters - self.num_channels:])
        self.plan.run(W, 'predictor', cat0, fast=ops.fast('kp'))
        pred = ops.conv2d(full, W['heads'], pad=0, fast=ops.fast('kp'))                       # (B,58,58,5K): kp logits | jacobian maps
        value, jac = ops.kp_head(pred, self.num_kp, float(self.temperature))
        return {'value': value, 'jacobian': jac}


@ARCH_REGISTRY.register()
class DenseMotionNetwork(ParamModule):
    """basicsr/archs/dense_motion_arch.py:12-161 (single occlusion map)."""

    def __init__(self, block_expansion, num_blocks, max_features, num_kp, num_channels, estimate_occlusion_map=False,
                 scale_factor=1, kp_variance=0.01, multi_mask=False, occlusion_num=5, model_path=None):
        super().__init__()
        if scale_factor != 0.25 or multi_mask or not estimate_occlusion_map:
            raise NotImplementedError('B200 DenseMotionNetwork implements the options/test.yml configuration')
        self.num_kp, self.kp_variance, self.num_channels = num_kp, kp_variance, num_channels
        self.plan = _HourglassPlan(block_expansion, (num_kp + 1) * (num_channels + 1), num_blocks, max_features)
        self.plan.declare(self, 'hourglass')
        self.declare_conv('mask', num_kp + 1, self.plan.out_filters, 7)
        self.declare_conv('occlusion', 1, self.plan.out_filters, 7)
        self.declare('down.weight', (num_channels, 1, 13, 13), _gaussian13, buffer=True)
        self._src_cache = None
        if model_path is not None:
            ck = torch.load(model_path, map_location='cpu')
            pre = 'module.dense_motion_network.'
            self.load_state_dict({k.replace(pre, ''): v for k, v in ck['generator'].items() if k.startswith(pre)})

    def _weights(self):
        if self._packed is None:
            T = self.tensors()
            W = self._restore_packed()
            if W is not None:
                self._packed, self._src_cache, self._pack_saved = W, None, True
                return W
            W = self.plan.pack(T, 'hourglass')
            W['heads'] = ops.pack_conv_cat([T['mask.weight'], T['occlusion.weight']], [T['mask.bias'], T['occlusion.bias']])
            W['k13'] = T['down.weight'][0, 0].contiguous()
            self._packed = W
            self._src_cache = None
        return self._packed

    def source_down(self, source_image: torch.Tensor) -> torch.Tensor:
        """Anti-aliased 64x64 source (1,64,64,3) NHWC; cached per source tensor (the reference recomputes
        it for every frame, dense_motion_arch.py:119-120)."""
        W = self._weights()
        c = self._src_cache           # keyed on the tensor object (kept alive by the entry: its address cannot be recycled) + version
        if c is None or c[0] is not source_image or c[1] != source_image._version:
            c = self._src_cache = (source_image, source_image._version, ops.antialias_down4(source_image[:1].contiguous().float(), W['k13']))
        return c[2]

    def clear_source_cache(self):
        self._src_cache = None

    @torch.no_grad()
    def forward(self, source_image, kp_driving, kp_source):
        W = self._weights()
        B, K = kp_driving['value'].shape[:2]
        if source_image.shape[0] != 1 and source_image.shape[0] != B:
            raise ValueError('source batch must be 1 or match the keypoint batch')
        if source_image.shape[0] > 1:
            # per-frame sources (cross-identity batches): run them one source at a time
            outs = [self.forward(source_image[i:i + 1], {k: v[i:i + 1] for k, v in kp_driving.items()},
                                 {k: v[i:i + 1] for k, v in kp_source.items()}) for i in range(B)]
            return {k: torch.cat([o[k] for o in outs], dim=0) for k in outs[0]}
        src64 = self.source_down(source_image)
        h, w = src64.shape[1], src64.shape[2]
        sv, sj = kp_source['value'][:1].contiguous(), kp_source['jacobian'][:1].contiguous()
        dv, dj = kp_driving['value'].contiguous(), kp_driving['jacobian'].contiguous()
        cin = self.plan.cin
        cat0 = torch.empty((B, h, w, self.plan.out_filters), device=src64.device, dtype=torch.float32)
        heat = ops.dense_motion_prep(src64, sv, sj, dv, dj, cat0[..., self.plan.out_filters - cin:], self.kp_variance)
        feat = self.plan.run(W, 'hourglass', cat0, fast=ops.fast('s1'))
        logits = ops.conv2d(feat, W['heads'], pad=3, fast=ops.fast('s1'))                     # (B,64,64,K+2): mask logits | occlusion logit
        deform, occ, _ = ops.dense_motion_head(logits, sv, sj, dv, dj)
        return {'deformation': deform, 'occlusion_map': occ.view(B, 1, h, w),
                'driving_kp_heatmap': heat.permute(0, 3, 1, 2), '_driving_kp_heatmap_nhwc': heat,
                'source': src64.permute(0, 3, 1, 2)}


@ARCH_REGISTRY.register()
class Motion_Estimator_keypoint_aware(torch.nn.Module):
    """basicsr/archs/motion_estimator_arch.py:14-52."""

    def __init__(self, common_params, dense_motion_params, kp_detector_params):
        super().__init__()
        if kp_detector_params is None:
            raise NotImplementedError('Shoule have kp_detector.')
        if dense_motion_params is None:
            raise NotImplementedError('Shoule have dense_motion_network.')
        self.kp_detector = KPDetector(**common_params, **kp_detector_params)
        self.dense_motion_network = DenseMotionNetwork(**common_params, **dense_motion_params)

    def estimate_kp(self, image):
        return self.kp_detector(image)

    def estimate_motion_w_kp(self, kp_source, kp_driving, source_image):
        dense_motion = self.dense_motion_network(source_image=source_image, kp_driving=kp_driving, kp_source=kp_source)
        dense_motion.update({'kp_driving': kp_driving, 'kp_source': kp_source})
        return dense_motion

    def forward(self, driving_image, source_image, only_return_kp_driving=False, relative=False):
        kp_driving = self.kp_detector(driving_image)
        if only_return_kp_driving:
            return kp_driving
        kp_source = self.kp_detector(source_image, isSource=True)
        dense_motion = self.dense_motion_network(source_image=source_image, kp_driving=kp_driving, kp_source=kp_source)
        dense_motion.update({'kp_driving': kp_driving, 'kp_source': kp_source})
        return dense_motion
