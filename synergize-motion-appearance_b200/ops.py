"""Tensor-level wrappers over the C ABI (include/sma_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every computation is a
hand-written kernel in csrc/.  Tensors are fp32 CUDA, channels-last *views* `(B,H,W,C)` whose
strides satisfy stride(3)==1 and stride(1)==W*stride(2) (so channel slices of concat buffers and
batch-expanded per-source tensors are passed without copies).  The caller-allocates convention
follows the reference's native-op slot (basicsr/ops/dcn/deform_conv.py:51-64); CPU tensors raise
(no fallback, as deform_conv.py:55-56).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import os

import torch

from . import _lib
from ._lib import ConvDesc, check

ACT = {'none': 0, 'relu': 1, 'leaky': 2, 'gelu': 3, 'sigmoid': 4, 'swish': 5}

# Per-launch timing hook used by bench.py's roofline leg: when PROFILE is a list, the wrappers of the dominant
# kernels append (kind, algorithmic_flops, algorithmic_bytes, start_event, end_event) with CUDA events recorded on
# the launching (current) stream.  None (the default) costs nothing.
PROFILE = None


class _Prof:
    __slots__ = ('kind', 'flops', 'nbytes', 'e0', 'label')

    def __init__(self, kind, flops, nbytes, label=''):
        self.kind, self.flops, self.nbytes, self.label = kind, flops, nbytes, label

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, self.flops, self.nbytes, self.e0, e1, self.label))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _nhwc(t: torch.Tensor) -> Tuple[int, int, int, int, int, int]:
    """-> (B,H,W,C,bstride,ld) of a channels-last view; raises for anything the kernels cannot address."""
    if not t.is_cuda:
        raise _lib.SmaError('sma_b200 ops need CUDA tensors (no CPU fallback)')
    if t.dtype != torch.float32 or t.dim() != 4:
        raise _lib.SmaError(f'expected fp32 (B,H,W,C) view, got {t.dtype} {tuple(t.shape)}')
    B, H, W, Cc = t.shape
    sb, sh, sw, sc = t.stride()
    if Cc > 1 and sc != 1:
        raise _lib.SmaError('channel stride must be 1')
    if H > 1 and sh != W * sw:
        raise _lib.SmaError(f'row stride {sh} != W*ld {W * sw}')
    if B == 1:
        sb = 0 if sb == 0 else sb
    return B, H, W, Cc, sb, sw


PACK_EVENTS = 0          # number of weight-image packs since import (tests: a pre-packed cache load must not add any)


@dataclass
class ConvW:
    """Packed conv / linear weight: w[(ky*kw+kx)*Cin+c][ldw], bias[Cout], plus the tensor-core weight images, which are packed LAZILY - the
    first time a launch plan (sma_conv2d_fwd with plan_only) says a kernel needs them - so that only the image a layer really uses is ever
    built (and persisted by packcache.py)."""
    w: torch.Tensor
    bias: Optional[torch.Tensor]
    Cout: int
    Cin: int
    kh: int
    kw: int
    images: Optional[dict] = None            # ('tc', nt) tf32 image packed with N tile nt | ('tc16',) fp16 image | ('ts',) tensor-memory-operand image
    plans: Optional[dict] = None             # launch signature -> (kernel, nt) as reported by the library's plan mode
    _slices: Optional[dict] = None           # cache of cols() results
    dirty: bool = False                      # an image was packed since the last packcache save
    pack_dims: Optional[tuple] = None        # (Cin, kh, kw) the tensor-core image packers see when they differ from the conv geometry (pack_conv_plus_1x1)

    def image(self, kind: str, nt: int = 0) -> Optional[torch.Tensor]:
        global PACK_EVENTS
        if self.images is None:
            self.images = {}
        key = (kind, nt) if kind == 'tc' else (kind,)
        t = self.images.get(key, False)
        if t is False:
            with torch.cuda.device(self.w.device):
                t = {'tc': lambda: _pack_tc(self, nt), 'tc16': lambda: _pack_tc16(self), 'ts': lambda: _pack_ts(self)}[kind]()
            self.images[key] = t
            self.dirty = True
            PACK_EVENTS += 1
        return t

    # eager accessors (tests / the tensor-memory-operand experiment)
    @property
    def w_tc(self):
        return self.image('tc', 0)

    @property
    def w_tc16(self):
        return self.image('tc16')

    @property
    def w_ts(self):
        return self.image('ts')

    def cols(self, start: int, n: int) -> 'ConvW':
        """Output-column slice (free for the CUDA-core layout: row-major [K][ldw]; tensor-core images are packed on first use)."""
        assert start % 4 == 0
        if self._slices is None:
            self._slices = {}
        cw = self._slices.get((start, n))
        if cw is None:
            cw = ConvW(self.w[:, start:], None if self.bias is None else self.bias[start:start + n], n, self.Cin, self.kh, self.kw)
            self._slices[(start, n)] = cw
        return cw

    def as_patch(self, p: int) -> 'ConvW':
        """View a Linear(C*p*p -> N) whose features are ordered (p1 p2 c) as a pxp stride-p conv
        (appmotioncodebook_arch.py:222,229,236)."""
        assert self.kh == 1 and self.kw == 1 and self.Cin % (p * p) == 0
        return ConvW(self.w, self.bias, self.Cout, self.Cin // (p * p), p, p)


def _pack_tc(cw: 'ConvW', nt: int = 0) -> Optional[torch.Tensor]:
    lib = _lib.load()
    n = lib.sma_conv_weight_tc_floats(cw.Cout, cw.Cin, cw.kh, cw.kw, nt)
    if n <= 0:
        return None
    out = torch.empty((n,), device=cw.w.device, dtype=torch.float32)
    check(lib.sma_pack_conv_weight_tc(cw.w.data_ptr(), cw.w.stride(0), cw.Cout, cw.Cin, cw.kh, cw.kw, nt, out.data_ptr(), _stream()),
          'sma_pack_conv_weight_tc')
    return out


def _pack_tc16(cw: 'ConvW') -> Optional[torch.Tensor]:
    lib = _lib.load()
    Cin, kh, kw = cw.pack_dims or (cw.Cin, cw.kh, cw.kw)
    n = lib.sma_conv_weight_tc16_floats(cw.Cout, Cin, kh, kw)
    if n <= 0:
        return None
    out = torch.empty((n,), device=cw.w.device, dtype=torch.float32)
    check(lib.sma_pack_conv_weight_tc16(cw.w.data_ptr(), cw.w.stride(0), cw.Cout, Cin, kh, kw, out.data_ptr(), _stream()),
          'sma_pack_conv_weight_tc16')
    return out


def _pack_ts(cw: 'ConvW') -> Optional[torch.Tensor]:
    lib = _lib.load()
    n = lib.sma_conv_weight_ts_floats(cw.Cout, cw.Cin, cw.kh, cw.kw)
    if n <= 0:
        return None
    out = torch.empty((n,), device=cw.w.device, dtype=torch.float32)
    check(lib.sma_pack_conv_weight_ts(cw.w.data_ptr(), cw.w.stride(0), cw.Cout, cw.Cin, cw.kh, cw.kw, out.data_ptr(), _stream()),
          'sma_pack_conv_weight_ts')
    return out


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], bn: Optional[dict] = None, pad_cin: int = 0) -> ConvW:
    """OIHW conv weight or (N,K) linear weight -> ConvW, optionally folding an eval-mode BatchNorm.
    `pad_cin`: zero-pad the input channels to this count (the caller feeds a zero-padded activation buffer) so that
    narrow inputs (2, 35 channels) become eligible for the tensor-core kernels (Cin % 32 == 0)."""
    lib = _lib.load()
    w = weight.detach().float().contiguous()
    if w.dim() == 2:
        w = w.view(w.shape[0], w.shape[1], 1, 1)
    if pad_cin and pad_cin > w.shape[1]:
        wp_ = torch.zeros((w.shape[0], pad_cin, w.shape[2], w.shape[3]), device=w.device, dtype=torch.float32)
        wp_[:, :w.shape[1]] = w
        w = wp_
    Cout, Cin, kh, kw = w.shape
    ldw = (Cout + 3) // 4 * 4
    wp = torch.empty((kh * kw * Cin, ldw), device=w.device, dtype=torch.float32)
    need_bias = bias is not None or bn is not None
    bo = torch.empty((ldw,), device=w.device, dtype=torch.float32).zero_() if need_bias else None
    b = None if bias is None else bias.detach().float().contiguous()
    g = be = mu = var = None
    eps = 0.0
    if bn is not None:
        g, be, mu, var = (bn[k].detach().float().contiguous() for k in ('weight', 'bias', 'running_mean', 'running_var'))
        eps = float(bn.get('eps', 1e-5))
    check(lib.sma_pack_conv_weight(_ptr(w), _ptr(b), Cout, Cin, kh, kw, _ptr(g), _ptr(be), _ptr(mu), _ptr(var), eps,
                                   _ptr(wp), ldw, _ptr(bo), _stream()), 'sma_pack_conv_weight')
    return ConvW(wp, None if bo is None else bo[:Cout], Cout, Cin, kh, kw)


def pack_conv_plus_1x1(cw3: ConvW, cw1: ConvW) -> ConvW:
    """A k x k conv over x (cw3) and a 1x1 conv over x2 (cw1) summed in one accumulator (conv2d(..., x2=, x2_1x1=True)): a ResBlock's conv2 + its skip
    conv.  The weight rows are laid out in the order the kernel consumes them - x: 64-channel chunk outer, tap inner; then x2's 64-channel chunks - and
    packed as ONE 1x1 conv of that depth, so that both parts share the per-output-channel power-of-two scaling of the fp16 image."""
    assert cw1.kh == 1 and cw1.kw == 1 and cw3.Cout == cw1.Cout and cw3.Cin % 64 == 0 and cw1.Cin % 64 == 0
    taps, c = cw3.kh * cw3.kw, cw3.Cin
    idx = torch.cat([t * c + cc * 64 + torch.arange(64) for cc in range(c // 64) for t in range(taps)]).to(cw3.w.device)
    w = torch.cat([cw3.w.index_select(0, idx), cw1.w], dim=0).contiguous()
    bias = None if (cw3.bias is None and cw1.bias is None) else ((0 if cw3.bias is None else cw3.bias) + (0 if cw1.bias is None else cw1.bias)).contiguous()
    return ConvW(w, bias, cw3.Cout, cw3.Cin + cw1.Cin, cw3.kh, cw3.kw, pack_dims=(w.shape[0], 1, 1))


def pack_conv_cat(weights, biases, pad_cin: int = 0) -> ConvW:
    """Several convs over the same input fused along the output dim (q|k|v, scale.0|shift.0, kp|jacobian ...)."""
    w = torch.cat([x.detach().float() for x in weights], dim=0)
    b = torch.cat([x.detach().float() for x in biases], dim=0)
    return pack_conv(w, b, pad_cin=pad_cin)


def pack_conv_blockdiag(weights, biases) -> ConvW:
    """Sibling convs over *different* channel ranges of one concat buffer fused into one launch: conv i reads input
    channels [sum(Cin_<i), sum(Cin_<=i)) and writes output columns [sum(Cout_<i), ...)."""
    cin = sum(int(w.shape[1]) for w in weights); cout = sum(int(w.shape[0]) for w in weights)
    kh, kw = weights[0].shape[2], weights[0].shape[3]
    big = torch.zeros((cout, cin, kh, kw), device=weights[0].device, dtype=torch.float32)
    o = c = 0
    for w in weights:
        big[o:o + w.shape[0], c:c + w.shape[1]] = w.detach().float()
        o += w.shape[0]; c += w.shape[1]
    return pack_conv(big, torch.cat([x.detach().float() for x in biases], dim=0))


PREC = {'exact': 0, 'tf32x3': 1, 'tf32': 2, 'f16x3': 3, 'f16': 4, 'f16x2': 5}     # SMA_PREC_*
USE_TF32X3 = True        # let sma_conv2d_fwd pick a tcgen05 kernel where the shape allows (False: exact CUDA-core kernels everywhere)
ALLOW_TF32_1PASS = True  # honour `fast=True` requests (single pass)
USE_TS = False           # weights as the tensor-memory A operand (csrc/conv_ts.cu) where they fit in tensor memory; False: shared-memory-operand kernels
USE_MH_F16 = True        # 8-head E=256 attention: fp16-split kernel with pre-split tile images (csrc/attn_mh.cu); False: tf32 kernel (attn_tc.cu)
SPLIT_FUSE = os.environ.get('SMA_NO_SPLITFUSE', '0') != '1'      # q | k | v projections write the attention operand images in their epilogue (conv2d(attn_split=))
MHA_D4_MMA = os.environ.get('SMA_NO_D4_MMA', '0') != '1'      # head-dim-4 attention of single-pass stages (S3m) as register-level mma with P in fp16 (csrc/attn.cu)
VQ_TILED = True          # large-N VQ lookups as the register-tiled exact-fp32 GEMM (bit-identical); False: warp-per-row kernel
FUSE_GN = os.environ.get('SMA_NO_FUSE_GN', '0') != '1'           # (env: A/B on one box) GroupNorm partial sums of a conv's output from its own epilogue (conv2d(gn=...)); False: always the standalone statistics pass
USE_F16 = True           # split operands into fp16 halves (kind::f16, 2x the tensor rate of kind::tf32) where Cin % 64 == 0
# Per-stage precision policy: stages listed here run their convolutions as single-pass TF32 (3x fewer tensor-core
# instructions); everything else is fp32-faithful 3xTF32.  See DESIGN.md section 4 for the measured error budget.
FAST_STAGES = {'s1', 's3m'}     # measured (tools/policy_err.py, gpurun_out/r2_policy_err.log): single-pass key-point detection moves the key-points by 8e-5 and the
                                # image by up to 1.9e-3 on general frames (the round-1 fixture clip hid it); S1 costs 8e-5, S3m 3e-5 of the 1e-3 budget


X2_STAGES: set = set()          # stages whose convolutions drop the activations' lo halves (weights hi + lo x activation hi: SMA_PREC_F16X2)


def fast(stage: str):
    """-> the `fast` argument of conv2d for a stage: True = single pass, 'x2' = two products, False = fp32-faithful three products."""
    return True if stage in FAST_STAGES else ('x2' if stage in X2_STAGES else False)


TC_VARIANT = int(os.environ.get('SMA_TC_VARIANT', '0'))           # (env override: A/B experiments on one box) 0: library picks the tensor-core kernel variant; 1: force the gather kernel (tests)
LAST_CONV_KERNEL = -1    # which kernel the last conv2d ran on: 0 CUDA-core, 1 tcgen05 gather, 2 tcgen05 persistent halo (tf32), 3 (fp16), 4 fp16 with weights in tensor memory


def conv2d(x: torch.Tensor, cw: ConvW, *, stride: int = 1, pad: int = 0, pad_tl: Optional[Tuple[int, int]] = None,
           out: Optional[torch.Tensor] = None, out_hw: Optional[Tuple[int, int]] = None, act: str = 'none',
           pre: Optional[Tuple[torch.Tensor, torch.Tensor, str]] = None, res: Optional[torch.Tensor] = None,
           upsample2: bool = False, d2s: int = 0, out_nchw: bool = False, exact: bool = False, fast: bool = False,
           sft: Optional[Tuple[torch.Tensor, float]] = None, gn: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
           x2: Optional[torch.Tensor] = None, x2_1x1: bool = False, attn_split: Optional[Tuple[torch.Tensor, float]] = None):
    """`exact`: force the CUDA-core fp32 kernel; `fast`: allow single-pass TF32 on the tensor cores (layers whose
    contribution to the output error budget was measured to be negligible); default: 3xTF32 (fp32-faithful).
    `sft=(scale, w)`: Fuse_sft_block tail, y = res + w*(res*scale + conv(x)) (needs `res`): fused into the epilogue of the persistent
    tensor-core kernel; launches that cannot run there (exact mode, odd shapes) do conv -> sma_sft_combine instead.
    `gn=(gamma, beta)`: also return the GroupNorm(32, eps 1e-6) scale / shift of the OUTPUT, `(y, (scale, shift))`: the partial sums come out of the
    convolution's epilogue where the persistent tensor-core kernel runs (no second pass over y), else from groupnorm_stats(y).
    `attn_split=(workspace, softmax_scale)`: a flat 1x1 layer producing q (256 columns) or q | k | v (768) of an E = 256 attention writes the attention
    kernels' fp16 hi / lo operand images into `workspace` instead of fp32 rows (returns None; then mha(..., presplit=)); SmaError(unsupported) otherwise.
    `x2`: second input tensor of the same geometry; `cw` holds the weights of both (input channels of x first): conv(x, w1) + conv(x2, w2) in one
    accumulator (staged-input fp16 kernel only; raises SmaError(unsupported) otherwise - callers keep a two-convolution form)."""
    lib = _lib.load()
    B, Hi, Wi, Cin, ibs, ild = _nhwc(x)
    Cin1 = Cin
    if x2 is not None:
        B2, H2, W2, C2, ibs2, ild2 = _nhwc(x2)
        assert (B2, H2, W2) == (B, Hi, Wi), (tuple(x2.shape), tuple(x.shape))
        Cin = Cin1 + C2
    if Cin != cw.Cin:
        raise _lib.SmaError(f'conv2d: input has {Cin} channels, weight expects {cw.Cin}')
    pt, pl = (pad, pad) if pad_tl is None else pad_tl
    Hv, Wv = (Hi * 2, Wi * 2) if upsample2 else (Hi, Wi)
    if out_hw is None:
        Ho = (Hv + 2 * pt - cw.kh) // stride + 1
        Wo = (Wv + 2 * pl - cw.kw) // stride + 1
    else:
        Ho, Wo = out_hw
    d = ConvDesc()
    d.x, d.B, d.Hi, d.Wi, d.Cin, d.in_bstride, d.in_ld = x.data_ptr(), B, Hi, Wi, Cin, ibs, ild
    d.w, d.ldw, d.bias = cw.w.data_ptr(), cw.w.stride(0), _ptr(cw.bias)
    d.Cout, d.kh, d.kw, d.stride, d.pad_t, d.pad_l = cw.Cout, cw.kh, cw.kw, stride, pt, pl
    d.upsample2 = 1 if upsample2 else 0
    if x2 is not None:
        d.x2, d.in2_bstride, d.in2_ld, d.Cin1, d.x2_k1 = x2.data_ptr(), ibs2, ild2, Cin1, 1 if x2_1x1 else 0
    if pre is not None:
        d.pre_scale, d.pre_shift, d.pre_act = pre[0].data_ptr(), pre[1].data_ptr(), ACT[pre[2]]
    if out_nchw:
        if out is None:
            out = torch.empty((B, cw.Cout, Ho, Wo), device=x.device, dtype=torch.float32)
        d.y, d.out_bstride, d.out_ld = out.data_ptr(), cw.Cout * Ho * Wo, 0
    elif d2s > 1:
        Cq = cw.Cout // (d2s * d2s)
        if out is None:
            out = torch.empty((B, Ho * d2s, Wo * d2s, Cq), device=x.device, dtype=torch.float32)
        oB, oH, oW, oC, obs, old = _nhwc(out)
        assert (oB, oH, oW, oC) == (B, Ho * d2s, Wo * d2s, Cq), (out.shape, (B, Ho * d2s, Wo * d2s, Cq))
        d.y, d.out_bstride, d.out_ld = out.data_ptr(), obs, old
    elif attn_split is not None:
        out = attn_split[0]                               # (y is not written: the images go to the workspace; any aligned address satisfies the ABI's checks)
        d.y, d.out_bstride, d.out_ld = out.data_ptr(), Ho * Wo * cw.Cout, cw.Cout
        d.split_ws, d.split_qscale = attn_split[0].data_ptr(), float(attn_split[1]) * 1.4426950408889634
    else:
        if out is None:
            out = torch.empty((B, Ho, Wo, cw.Cout), device=x.device, dtype=torch.float32)
        oB, oH, oW, oC, obs, old = _nhwc(out)
        assert (oB, oH, oW, oC) == (B, Ho, Wo, cw.Cout), (tuple(out.shape), (B, Ho, Wo, cw.Cout))
        d.y, d.out_bstride, d.out_ld = out.data_ptr(), obs, old
    d.Ho, d.Wo, d.act = Ho, Wo, ACT[act]
    if res is not None:
        rB, rH, rW, rC, rbs, rld = _nhwc(res)
        assert (rB, rH, rW, rC) == (B, Ho, Wo, cw.Cout), (tuple(res.shape), (B, Ho, Wo, cw.Cout))
        d.res, d.res_bstride, d.res_ld = res.data_ptr(), rbs, rld
    d.d2s, d.out_nchw = d2s, 1 if out_nchw else 0
    if sft is not None:
        assert res is not None
        aB, aH, aW, aC, abs_, ald = _nhwc(sft[0])
        assert (aB, aH, aW, aC) == (B, Ho, Wo, cw.Cout)
        d.aux, d.aux_bstride, d.aux_ld, d.sft_w = sft[0].data_ptr(), abs_, ald, float(sft[1])
    d.tc_variant = TC_VARIANT
    d.kernel_used = -1
    planned = None
    gn_partial = None
    if exact or not USE_TF32X3 or Cin % 32:
        d.precision = PREC['exact']
    else:
        one = fast is True and ALLOW_TF32_1PASS
        d.precision = (PREC['f16'] if one else (PREC['f16x2'] if fast == 'x2' else PREC['f16x3'])) if USE_F16 else (PREC['tf32'] if one else PREC['tf32x3'])
        if USE_TS:       # the tensor-memory-operand experiment: every image up front, the library picks
            d.w_tc, d.w_tc16, d.w_ts = _ptr(cw.image('tc', 0)), _ptr(cw.image('tc16')), _ptr(cw.image('ts'))
        else:
            # ask the library which kernel this launch will run on (nothing is launched), then pack / fetch exactly the image it reads
            gn_C = cw.Cout // (d2s * d2s) if d2s > 1 else cw.Cout                  # channels of the normalised tensor (a depth-to-space conv folds d2s^2 column blocks)
            gn_groups = gn[2] if (gn is not None and len(gn) > 2) else 32
            d.gn_want = 1 if (gn is not None and FUSE_GN and not out_nchw and gn_C % (2 * gn_groups) == 0) else 0      # (pairs must not straddle the groups)
            key = (B, Hi, Wi, stride, pt, pl, upsample2, Ho, Wo, out_nchw, d2s, d.precision, TC_VARIANT, x.data_ptr() & 15, ild & 3, ibs & 3,
                   pre is not None, sft is not None, 0 if res is None else (d.res_ld & 7, d.res_bstride & 7), d.out_ld & 7, d.out_bstride & 7,
                   d.gn_want, out.data_ptr() & 31, 0 if res is None else res.data_ptr() & 31,
                   None if x2 is None else (Cin1, ild2 & 3, ibs2 & 3, x2.data_ptr() & 15, x2_1x1), attn_split is not None)
            if cw.plans is None:
                cw.plans = {}
            planned = cw.plans.get(key)
            if planned is None:
                d.plan_only = 1
                check(lib.sma_conv2d_fwd(C.byref(d), _stream()), 'sma_conv2d_fwd(plan)')
                planned = cw.plans[key] = (d.kernel_used, d.w_tc_nt, d.gn_chunks)
                d.plan_only, d.kernel_used, d.w_tc_nt = 0, -1, 0
            if planned[0] in (1, 2):
                d.w_tc, d.w_tc_nt = _ptr(cw.image('tc', planned[1])), planned[1]
            elif planned[0] == 3:
                d.w_tc16 = _ptr(cw.image('tc16'))
            gn_partial = None
            if d.gn_want and planned[2] > 0:
                gn_partial = torch.empty((B * planned[2] * cw.Cout,), device=x.device, dtype=torch.float32)
                d.gn_partial = gn_partial.data_ptr()
            else:
                d.gn_want = 0
    if attn_split is not None and (planned is None or planned[0] != 3):
        raise _lib.SmaError('conv2d(attn_split=...): needs the persistent fp16 kernel (unsupported shape / layout): write fp32 rows and let the attention split them')
    if x2 is not None and (planned is None or planned[0] != 3):
        raise _lib.SmaError('conv2d(x2=...): the two-tensor input needs the staged-input fp16 kernel (unsupported shape / layout): run the two convolutions')
    if sft is not None and (planned is None or planned[0] not in (2, 3)):
        # unfused form: the CUDA-core / gather kernels have no SFT epilogue
        shift = conv2d(x, cw, stride=stride, pad=pad, pad_tl=pad_tl, out_hw=out_hw, act=act, pre=pre, upsample2=upsample2, exact=exact, fast=fast)
        dec = res if res.is_contiguous() else affine_act(res, None, None)
        sc = sft[0] if sft[0].is_contiguous() else affine_act(sft[0], None, None)
        return sft_combine(dec, sc, shift, float(sft[1]), out=out if (out is not None and out.is_contiguous()) else None)
    K = cw.kh * cw.kw * Cin if not (x2 is not None and x2_1x1) else cw.kh * cw.kw * Cin1 + (Cin - Cin1)      # (the second tensor enters through one tap)
    with _Prof('conv', 2.0 * B * Ho * Wo * K * cw.Cout,
               4.0 * (B * Hi * Wi * Cin + B * Ho * Wo * cw.Cout * (2 if res is not None else 1) + K * cw.Cout),
               f'conv B{B} {Hi}x{Wi} Cin{Cin} Cout{cw.Cout} k{cw.kh} s{stride}{" up" if upsample2 else ""}{" pre" if pre is not None else ""}') as pr:
        check(lib.sma_conv2d_fwd(C.byref(d), _stream()), f'sma_conv2d_fwd Cin={Cin} Cout={cw.Cout} k={cw.kh}x{cw.kw}')
        pr.label += (' simt', ' tc-gather', ' tc-halo', ' tc-halo-f16', ' ts-f16')[d.kernel_used] + (' 1pass' if d.precision in (2, 4) else (' x2' if d.precision == 5 and d.kernel_used == 3 else ''))
    if DET_CHECK and attn_split is None and d.res != d.y and d.x != d.y:
        # debugging aid (SMA_DET_CHECK=1): launch every convolution twice and compare the outputs bit for bit
        first = out.clone()
        gp1 = None if gn_partial is None else gn_partial.clone()
        check(lib.sma_conv2d_fwd(C.byref(d), _stream()), 'sma_conv2d_fwd(det)')
        same = torch.equal(first, out) and (gp1 is None or torch.equal(gp1, gn_partial))
        if not same:
            nd = int((first != out).sum())
            print(f'DET_CHECK: conv B{B} {Hi}x{Wi} Cin{Cin} Cout{cw.Cout} k{cw.kh} s{stride} up{upsample2} pre{pre is not None} res{res is not None} kernel{d.kernel_used} '
                  f'prec{d.precision} gn{d.gn_want} x2{x2 is not None}: {nd} differing outputs, first at {(first != out).nonzero()[0].tolist() if nd else None}', flush=True)
    if planned is not None and d.kernel_used != planned[0]:
        raise _lib.SmaError(f'conv2d ran on kernel {d.kernel_used} but was planned on {planned[0]} (Cin={Cin} Cout={cw.Cout} k={cw.kh}): binding bug')
    global LAST_CONV_KERNEL
    LAST_CONV_KERNEL = d.kernel_used
    if attn_split is not None:
        return None
    if gn is not None:
        # gn = (gamma, beta[, groups[, scale_out, shift_out]]): the optional outputs are (B, C) column slices of wider buffers (the two halves of a
        # concatenated tensor are normalised with separate statistics: GroupNorm's groups do not straddle them)
        groups = gn[2] if len(gn) > 2 else 32
        C_gn = cw.Cout // (d2s * d2s) if d2s > 1 else cw.Cout
        scale = gn[3] if len(gn) > 3 else torch.empty((B, C_gn), device=x.device, dtype=torch.float32)
        shift = gn[4] if len(gn) > 4 else torch.empty((B, C_gn), device=x.device, dtype=torch.float32)
        if d.gn_want and d.gn_chunks > 0:
            check(lib.sma_groupnorm_finalize_pairs(gn_partial.data_ptr(), B, d.gn_chunks, C_gn, groups, Ho * Wo, d2s * d2s if d2s > 1 else 1, 1e-6,
                                                   gn[0].data_ptr(), gn[1].data_ptr(), scale.data_ptr(), shift.data_ptr(), scale.stride(0), _stream()),
                  'sma_groupnorm_finalize_pairs')
        else:
            sc_, sh_ = groupnorm_stats(out, gn[0], gn[1], groups, 1e-6)
            scale.copy_(sc_); shift.copy_(sh_)
        return out, (scale, shift)
    return out


def linear(x: torch.Tensor, cw: ConvW, **kw) -> torch.Tensor:
    """Linear over the last dim of a (B,L,E) token tensor == 1x1 conv over (B,1,L,E)."""
    y = conv2d(x.unsqueeze(1), cw, **{k: (v.unsqueeze(1) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
    return None if y is None else y.squeeze(1)


def groupnorm_stats(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32, eps: float = 1e-6):
    """-> (scale, shift) of shape (B,C) such that GN(x)[b,:,c] = x*scale[b,c]+shift[b,c]."""
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(x)
    HW = H * W
    nchunk = (HW + 255) // 256
    partial = torch.empty((B * nchunk * Cc * 2,), device=x.device, dtype=torch.float32)
    scale = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    shift = torch.empty((B, Cc), device=x.device, dtype=torch.float32)
    check(lib.sma_groupnorm_stats(x.data_ptr(), B, HW, Cc, bs, ld, groups, eps, gamma.data_ptr(), beta.data_ptr(),
                                  partial.data_ptr(), scale.data_ptr(), shift.data_ptr(), _stream()), 'sma_groupnorm_stats')
    return scale, shift


def affine_act(x: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], act: str = 'none',
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(x)
    if out is None:
        out = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    check(lib.sma_affine_act(x.data_ptr(), B, H * W, Cc, bs, ld, _ptr(scale), _ptr(shift), ACT[act], out.data_ptr(), obs, old,
                             _stream()), 'sma_affine_act')
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, pos: Optional[torch.Tensor] = None,
              want_y: bool = True, eps: float = 1e-5):
    """x (B,L,E) contiguous -> (LN(x), LN(x)+pos) ; either may be skipped (None)."""
    lib = _lib.load()
    assert x.is_contiguous() and x.is_cuda
    E = x.shape[-1]
    rows = x.numel() // E
    y = torch.empty_like(x) if want_y else None
    yq = torch.empty_like(x) if pos is not None else None
    check(lib.sma_layernorm(x.data_ptr(), rows, E, gamma.data_ptr(), beta.data_ptr(), eps, _ptr(pos),
                            0 if pos is None else pos.shape[0], _ptr(y), _ptr(yq), _stream()), 'sma_layernorm')
    return y, yq


def warp_occlude(feat: torch.Tensor, flow: torch.Tensor, occ: Optional[torch.Tensor], out: Optional[torch.Tensor] = None):
    """feat (B or 1-expanded,H,W,C) NHWC contiguous per frame; flow (B,hf,wf,2); occ (B,hf,wf) or None."""
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(feat)
    assert ld == Cc, 'warp source must be dense NHWC'
    assert flow.is_contiguous() and flow.shape[0] == B and flow.shape[3] == 2
    hf, wf = flow.shape[1], flow.shape[2]
    if occ is not None:
        assert occ.is_contiguous() and occ.numel() == B * hf * wf
    if out is None:
        out = torch.empty((B, H, W, Cc), device=feat.device, dtype=torch.float32)
    assert out.is_contiguous()
    with _Prof('warp', 11.0 * B * H * W * Cc, 8.0 * B * H * W * Cc):       # read every feature once + write it once
        check(lib.sma_warp_occlude_fwd(feat.data_ptr(), bs, B, H, W, Cc, flow.data_ptr(), _ptr(occ), hf, wf, out.data_ptr(), _stream()),
              'sma_warp_occlude_fwd')
    return out


def resize_ac(x: torch.Tensor, size: Tuple[int, int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    B, Hi, Wi, Cc, ibs, ild = _nhwc(x)
    Ho, Wo = size
    if out is None:
        out = torch.empty((B, Ho, Wo, Cc), device=x.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    check(lib.sma_resize_bilinear_ac(x.data_ptr(), B, Hi, Wi, Cc, ibs, ild, out.data_ptr(), Ho, Wo, obs, old, _stream()),
          'sma_resize_bilinear_ac')
    return out


def warp_occlude_gather(feat: torch.Tensor, flow: torch.Tensor, occ: Optional[torch.Tensor], size: Tuple[int, int]) -> torch.Tensor:
    """warp_occlude evaluated only at the four bilinear neighbours of every sample of a later resize to `size`: (B,2Ho,2Wo,C), the layout blend_bil4 consumes."""
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(feat)
    assert ld == Cc and flow.is_contiguous() and flow.shape[0] == B
    hf, wf = flow.shape[1], flow.shape[2]
    out = torch.empty((B, 2 * size[0], 2 * size[1], Cc), device=feat.device, dtype=torch.float32)
    with _Prof('warp', 11.0 * B * 4 * size[0] * size[1] * Cc, 8.0 * B * 4 * size[0] * size[1] * Cc):
        check(lib.sma_warp_occlude_gather_fwd(feat.data_ptr(), bs, B, H, W, Cc, flow.data_ptr(), _ptr(occ), hf, wf, size[0], size[1], out.data_ptr(), _stream()),
              'sma_warp_occlude_gather_fwd')
    return out


def gather_bil4(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """(B,Hi,Wi,C) -> (B,2Ho,2Wo,C): the four bilinear (align_corners=True) neighbours of every sample of a resize to `size`."""
    lib = _lib.load()
    B, Hi, Wi, Cc, ibs, ild = _nhwc(x)
    Ho, Wo = size
    g = torch.empty((B, 2 * Ho, 2 * Wo, Cc), device=x.device, dtype=torch.float32)
    check(lib.sma_gather_bilinear4(x.data_ptr(), B, Hi, Wi, Cc, ibs, ild, g.data_ptr(), Ho, Wo, _stream()), 'sma_gather_bilinear4')
    return g


def blend_bil4(g: torch.Tensor, src_hw: Tuple[int, int], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,2Ho,2Wo,C) gathered neighbours (after a pointwise layer) -> (B,Ho,Wo,C): the bilinear blend of a resize from `src_hw`."""
    lib = _lib.load()
    assert g.is_contiguous()
    B, H2, W2, Cc = g.shape
    Ho, Wo = H2 // 2, W2 // 2
    if out is None:
        out = torch.empty((B, Ho, Wo, Cc), device=g.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    check(lib.sma_blend_bilinear4(g.data_ptr(), B, src_hw[0], src_hw[1], Cc, out.data_ptr(), Ho, Wo, obs, old, _stream()), 'sma_blend_bilinear4')
    return out


def mha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, key_mask: Optional[torch.Tensor] = None,
        scale: Optional[float] = None, out: Optional[torch.Tensor] = None, exact: bool = False, fast: bool = False) -> torch.Tensor:
    """q (B,L,E-view) ; k,v (B,S,E-view) or (S,E-view) shared by all frames.  Views may be column slices."""
    return _mha(q, k, v, heads, key_mask, scale, out, exact, fast)


def attn_workspace(B: int, kvB: int, L: int, S: int, device) -> torch.Tensor:
    """Workspace of the E = 256 attention kernels (fp16 hi / lo operand images of q, k, v): conv2d(attn_split=) writes into it, mha_presplit reads it."""
    return torch.empty((_lib.load().sma_mha_e256_workspace_bytes(B, kvB, L, S),), device=device, dtype=torch.uint8)


def attn_kv_images(k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """fp16 hi / lo operand images of k, v (S,256) shared by every frame (the cached codebook projections): split once, then mha_presplit(kv_images=)."""
    lib = _lib.load()
    S = k.shape[0]
    assert k.dim() == 2 and v.dim() == 2 and k.shape[1] == 256 and k.stride(1) == 1 and v.stride(1) == 1
    img = torch.empty((8 * S * 256,), device=k.device, dtype=torch.uint8)
    check(lib.sma_attn_split_kv(k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), S, img.data_ptr(), _stream()), 'sma_attn_split_kv')
    return img


def mha_presplit(ws: torch.Tensor, B: int, L: int, S: int, k: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None,
                 key_mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, kv_images: Optional[torch.Tensor] = None) -> torch.Tensor:
    """8-head E = 256 attention whose q images (k, v given: shared (S,E) codebook projections, split here) or q, k and v images (k = v = None:
    self-attention, S = L) were written into `ws` by the projection's epilogue."""
    lib = _lib.load()
    E = 256
    if out is None:
        out = torch.empty((B, L, E), device=ws.device, dtype=torch.float32)
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and key_mask.numel() == B * S
    shared = k is not None or kv_images is not None
    if kv_images is not None:
        assert kv_images.numel() == 8 * S * 256
        with _Prof('mha', 4.0 * B * L * S * E, 4.0 * (2 * B * L * E + 2 * S * E), f'mha-f16 B{B} L{L} S{S} h8 D32'):
            check(lib.sma_mha_e256_fwd(None, E, kv_images.data_ptr(), E, None, E, L * E, 0, B, L, S, 32 ** -0.5, _ptr(key_mask), ws.data_ptr(), out.data_ptr(),
                                       out.stride(1), 3, _stream()), 'sma_mha_e256_fwd(presplit 3)')
        return out
    if shared:
        assert k.dim() == 2 and v.dim() == 2 and k.stride(-1) == 1 and v.stride(-1) == 1
    with _Prof('mha', 4.0 * B * L * S * E, 4.0 * (2 * B * L * E + 2 * (1 if shared else B) * S * E), f'mha-f16 B{B} L{L} S{S} h8 D32'):
        check(lib.sma_mha_e256_fwd(None, E, _ptr(k), k.stride(0) if shared else E, _ptr(v), v.stride(0) if shared else E, L * E, 0 if shared else S * E,
                                   B, L, S, 32 ** -0.5, _ptr(key_mask), ws.data_ptr(), out.data_ptr(), out.stride(1), 1 if shared else 2, _stream()),
              'sma_mha_e256_fwd(presplit)')
    return out


def _mha(q, k, v, heads, key_mask=None, scale=None, out=None, exact=False, fast=False):
    lib = _lib.load()
    B, L, E = q.shape
    D = E // heads
    if k.dim() == 2:
        S, kvbs = k.shape[0], 0
        ldk, ldv = k.stride(0), v.stride(0)
    else:
        S, kvbs = k.shape[1], k.stride(0)
        ldk, ldv = k.stride(1), v.stride(1)
        assert v.stride(0) == kvbs
    assert q.stride(2) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1 and q.stride(0) == L * q.stride(1)
    if out is None:
        out = torch.empty((B, L, E), device=q.device, dtype=torch.float32)
    if scale is None:
        scale = float(D) ** -0.5
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and key_mask.numel() == B * S
    if D == 256 and heads == 1 and key_mask is None and kvbs and not exact and USE_TF32X3 and L % 128 == 0 and S % 64 == 0:
        return attn256(q, k, v, scale, out)
    if D == 32 and heads == 8 and not exact and USE_TF32X3 and USE_MH_F16 and L % 128 == 0 and S % 64 == 0:
        ws = torch.empty((lib.sma_mha_e256_workspace_bytes(B, B if kvbs else 1, L, S),), device=q.device, dtype=torch.uint8)
        with _Prof('mha', 4.0 * B * L * S * E, 4.0 * (2 * B * L * E + 2 * (B if kvbs else 1) * S * E), f'mha-f16 B{B} L{L} S{S} h{heads} D{D}'):
            check(lib.sma_mha_e256_fwd(q.data_ptr(), q.stride(1), k.data_ptr(), ldk, v.data_ptr(), ldv, q.stride(0), kvbs, B, L, S, scale,
                                       _ptr(key_mask), ws.data_ptr(), out.data_ptr(), out.stride(1), 0, _stream()), 'sma_mha_e256_fwd')
        return out
    with _Prof('mha', 4.0 * B * L * S * E, 4.0 * (2 * B * L * E + 2 * (B if kvbs else 1) * S * E), f'mha B{B} L{L} S{S} h{heads} D{D}'):
        check(lib.sma_mha_fwd(q.data_ptr(), q.stride(1), k.data_ptr(), ldk, v.data_ptr(), ldv, kvbs, B, L, S, heads, D, scale,
                              _ptr(key_mask), out.data_ptr(), out.stride(1), 1 if (exact or not USE_TF32X3) else (2 if (fast is True and MHA_D4_MMA) else 0), _stream()), f'sma_mha_fwd D={D}')
    return out


def attn256(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Single-head, head-dim-256 attention on the tensor cores (AttnBlock).  q (B,L,256-view), k, v (B,S,256-view): column slices of one
    qkv buffer are fine."""
    lib = _lib.load()
    B, L, D = q.shape
    S = k.shape[1]
    assert D == 256 and q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1 and k.stride(0) == v.stride(0)
    if out is None:
        out = torch.empty((B, L, D), device=q.device, dtype=torch.float32)
    ws = torch.empty((lib.sma_attn256_workspace_bytes(B, L, S),), device=q.device, dtype=torch.uint8)
    with _Prof('mha', 4.0 * B * L * S * D, 4.0 * (2 * B * L * D + 2 * B * S * D), f'attn256 B{B} L{L} S{S}'):
        check(lib.sma_attn256_fwd(q.data_ptr(), q.stride(1), k.data_ptr(), k.stride(1), v.data_ptr(), v.stride(1), q.stride(0), k.stride(0),
                                  B, L, S, scale, ws.data_ptr(), out.data_ptr(), out.stride(1), 0, _stream()), 'sma_attn256_fwd')
    return out


def attn256_workspace(B: int, L: int, device) -> torch.Tensor:
    return torch.empty((_lib.load().sma_attn256_workspace_bytes(B, L, L),), device=device, dtype=torch.uint8)


def attn256_presplit(ws: torch.Tensor, B: int, L: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """AttnBlock attention whose q | k | v operand images were written into `ws` by the qkv conv's epilogue (conv2d(attn_split=(ws, 256 ** -0.5)))."""
    lib = _lib.load()
    if out is None:
        out = torch.empty((B, L, 256), device=ws.device, dtype=torch.float32)
    with _Prof('mha', 4.0 * B * L * L * 256, 4.0 * 4 * B * L * 256, f'attn256 B{B} L{L} S{L}'):
        check(lib.sma_attn256_fwd(None, 256, None, 256, None, 256, L * 256, L * 256, B, L, L, 1.0, ws.data_ptr(), out.data_ptr(), out.stride(1), 1, _stream()),
              'sma_attn256_fwd(presplit)')
    return out


def vq_lookup(z: torch.Tensor, codebook: torch.Tensor, n_codes: Optional[int] = None):
    """z (N,E) contiguous -> (idx int64 (N,), zq (N,E), min_dist (N,))"""
    lib = _lib.load()
    assert z.is_cuda and z.is_contiguous() and codebook.is_contiguous()
    N, E = z.shape
    n = codebook.shape[0] if n_codes is None else n_codes
    idx = torch.empty((N,), device=z.device, dtype=torch.int64)
    zq = torch.empty_like(z)
    md = torch.empty((N,), device=z.device, dtype=torch.float32)
    ws = torch.empty((n,), device=z.device, dtype=torch.float32) if VQ_TILED else None
    check(lib.sma_vq_lookup_fwd(z.data_ptr(), N, E, codebook.data_ptr(), n, idx.data_ptr(), zq.data_ptr(), md.data_ptr(), _ptr(ws), _stream()),
          'sma_vq_lookup_fwd')
    return idx, zq, md


def vq_quantize(z: torch.Tensor, codebook: torch.Tensor, n_codes: Optional[int] = None, beta: float = 0.25):
    """VectorQuantizer.forward's forward values (archs/vqgan_arch.py:33-93) on channels-last rows: z (..., E) contiguous ->
    (z + (zq - z) with z's shape, loss (0-dim device tensor), idx int64 (N,))."""
    lib = _lib.load()
    E = z.shape[-1]
    zf = z.reshape(-1, E)
    idx, zq, _ = vq_lookup(zf, codebook, n_codes)
    st = torch.empty_like(zf)
    ws = torch.empty((lib.sma_vq_workspace_floats() + 1,), device=z.device, dtype=torch.float32)
    check(lib.sma_vq_commit_fwd(zf.data_ptr(), zq.data_ptr(), zf.numel(), float(beta), st.data_ptr(), ws.data_ptr(), ws[-1:].data_ptr(), _stream()),
          'sma_vq_commit_fwd')
    return st.view(z.shape), ws[-1], idx


def antialias_down4(x_nchw: torch.Tensor, kernel13: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    assert x_nchw.is_cuda and x_nchw.is_contiguous() and x_nchw.dtype == torch.float32
    B, Cc, H, W = x_nchw.shape
    if out is None:
        out = torch.empty((B, H // 4, W // 4, Cc), device=x_nchw.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    assert obs == (H // 4) * (W // 4) * old or B == 1
    check(lib.sma_antialias_down4(x_nchw.data_ptr(), B, Cc, H, W, kernel13.data_ptr(), out.data_ptr(), old, _stream()),
          'sma_antialias_down4')
    return out


def avgpool2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_contiguous()
    B, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, H // 2, W // 2, Cc), device=x.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    assert obs == (H // 2) * (W // 2) * old or B == 1
    check(lib.sma_avgpool2(x.data_ptr(), B, H, W, Cc, out.data_ptr(), old, _stream()), 'sma_avgpool2')
    return out


def kp_head(pred: torch.Tensor, K: int, temperature: float):
    lib = _lib.load()
    B, h, w, Cc, bs, ld = _nhwc(pred)
    value = torch.empty((B, K, 2), device=pred.device, dtype=torch.float32)
    jac = torch.empty((B, K, 2, 2), device=pred.device, dtype=torch.float32)
    check(lib.sma_kp_head_fwd(pred.data_ptr(), B, h, w, ld, K, temperature, value.data_ptr(), jac.data_ptr(), _stream()), 'sma_kp_head_fwd')
    return value, jac


def normalize_kp(src_v, src_j, drv_v, drv_j, drv0_v, drv0_j, scale, relative: bool):
    """`scale`: python float, or a 1-element device tensor (written by hull_scale: no host round trip).  Source / initial key-points are
    per-clip constants: batch 1."""
    lib = _lib.load()
    B, K, _ = drv_v.shape
    if src_v.shape[0] != 1 or drv0_v.shape[0] != 1 or src_j.shape[0] != 1 or drv0_j.shape[0] != 1:
        raise _lib.SmaError('normalize_kp: kp_source and kp_driving_initial must have batch 1 (per-clip constants)')
    ov, oj = torch.empty_like(drv_v), torch.empty_like(drv_j)
    dev_scale = scale if isinstance(scale, torch.Tensor) else None
    check(lib.sma_normalize_kp(src_v.data_ptr(), src_j.data_ptr(), drv_v.data_ptr(), drv_j.data_ptr(), drv0_v.data_ptr(),
                               drv0_j.data_ptr(), B, K, 1.0 if dev_scale is not None else float(scale), _ptr(dev_scale), 1 if relative else 0,
                               ov.data_ptr(), oj.data_ptr(), _stream()), 'sma_normalize_kp')
    return ov, oj


def hull_scale(src_v: torch.Tensor, drv0_v: torch.Tensor) -> torch.Tensor:
    """sqrt(hull area of the source key-points) / sqrt(hull area of the initial driving key-points) as a device scalar (demo.py:26-29)."""
    lib = _lib.load()
    assert src_v.is_cuda and src_v.is_contiguous() and drv0_v.is_contiguous() and src_v.dtype == torch.float32
    K = src_v.shape[-2]
    out = torch.empty((1,), device=src_v.device, dtype=torch.float32)
    check(lib.sma_hull_scale(src_v.data_ptr(), drv0_v.data_ptr(), K, out.data_ptr(), _stream()), 'sma_hull_scale')
    return out


def u8hwc_to_f32nchw(frames: torch.Tensor, swap_rb: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,H,W,C) uint8 device frames -> (B,C,H,W) fp32 in [-1,1]: the reference's astype(float32)/255 + img2tensor + normalize(0.5,0.5)
    (demo.py:177-185) on the device, bit-exact."""
    lib = _lib.load()
    if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4 or not frames.is_contiguous():
        raise _lib.SmaError('u8hwc_to_f32nchw needs a contiguous (B,H,W,C) uint8 CUDA tensor')
    B, H, W, Cc = frames.shape
    if out is None:
        out = torch.empty((B, Cc, H, W), device=frames.device, dtype=torch.float32)
    assert out.is_contiguous() and tuple(out.shape) == (B, Cc, H, W)
    check(lib.sma_u8hwc_to_f32nchw(frames.data_ptr(), B, H, W, Cc, 1 if swap_rb else 0, out.data_ptr(), _stream()), 'sma_u8hwc_to_f32nchw')
    return out


def dense_motion_prep(src64: torch.Tensor, kp_src_v, kp_src_j, kp_drv_v, kp_drv_j, hg_in: torch.Tensor, var: float = 0.01):
    lib = _lib.load()
    h, w = src64.shape[-3], src64.shape[-2]
    B, K, _ = kp_drv_v.shape
    _, _, _, _, hbs, hld = _nhwc(hg_in)
    heat = torch.empty((B, h, w, K), device=hg_in.device, dtype=torch.float32)
    check(lib.sma_dense_motion_prep(src64.data_ptr(), h, w, kp_src_v.data_ptr(), kp_src_j.data_ptr(), kp_drv_v.data_ptr(),
                                    kp_drv_j.data_ptr(), B, K, var, hg_in.data_ptr(), hld, heat.data_ptr(), _stream()),
          'sma_dense_motion_prep')
    return heat


def dense_motion_head(logits: torch.Tensor, kp_src_v, kp_src_j, kp_drv_v, kp_drv_j, want_mask: bool = False):
    lib = _lib.load()
    B, h, w, Cc, bs, ld = _nhwc(logits)
    K = kp_drv_v.shape[1]
    deform = torch.empty((B, h, w, 2), device=logits.device, dtype=torch.float32)
    occ = torch.empty((B, h, w), device=logits.device, dtype=torch.float32)
    mask = torch.empty((B, h, w, K + 1), device=logits.device, dtype=torch.float32) if want_mask else None
    check(lib.sma_dense_motion_head(logits.data_ptr(), ld, h, w, kp_src_v.data_ptr(), kp_src_j.data_ptr(), kp_drv_v.data_ptr(),
                                    kp_drv_j.data_ptr(), B, K, deform.data_ptr(), occ.data_ptr(), _ptr(mask), _stream()),
          'sma_dense_motion_head')
    return deform, occ, mask


def flow_to_px(m: torch.Tensor, out: torch.Tensor):
    lib = _lib.load()
    B, h, w, _ = m.shape
    assert m.is_contiguous()
    _, _, _, _, obs, old = _nhwc(out)
    check(lib.sma_flow_to_px(m.data_ptr(), B, h, w, out.data_ptr(), old, _stream()), 'sma_flow_to_px')
    return out


def im2col_small(x: torch.Tensor, C: int, k: int, pad: int, Kp: int) -> torch.Tensor:
    """x (B,H,W,>=C) channels-last view (the first C channels are used) -> (B,H,W,Kp), column (ky*k+kx)*C + c; zero padded."""
    lib = _lib.load()
    B, H, W, _, bs, ld = _nhwc(x)
    assert bs == H * W * ld
    out = torch.empty((B, H, W, Kp), device=x.device, dtype=torch.float32)
    check(lib.sma_im2col_small(x.data_ptr(), B, H, W, ld, C, k, pad, out.data_ptr(), Kp, _stream()), 'sma_im2col_small')
    return out


def pack_conv_unfolded(weight: torch.Tensor, bias: Optional[torch.Tensor], Kp: int) -> 'ConvW':
    """OIHW weight of a few-channel conv -> the (O, Kp) linear weight over im2col_small's columns ((ky*k+kx)*C + c)."""
    O, Cc, kh, kw = weight.shape
    w2 = torch.zeros((O, Kp), device=weight.device, dtype=torch.float32)
    w2[:, :kh * kw * Cc] = weight.detach().float().permute(0, 2, 3, 1).reshape(O, kh * kw * Cc)
    return pack_conv(w2, bias)


DET_CHECK = os.environ.get('SMA_DET_CHECK', '0') == '1'      # debugging aid: every convolution is launched twice and the two outputs compared bit for bit
TAPSUM = os.environ.get('SMA_NO_TAPSUM', '0') != '1'      # 3x3 convs with <= 4 outputs as a pointwise layer over (tap, c) columns + a gather-sum (conv_tapsum)


def pack_conv_tapcols(weight: torch.Tensor) -> 'ConvW':
    """OIHW weight of a k x k conv with few outputs -> the (k*k*O padded to 8, Cin) pointwise weight whose column (ky*k+kx)*O + o is tap (ky,kx) of output o
    (the bias is added by conv_tapsum)."""
    O, Cc, kh, kw = weight.shape
    n = kh * kw * O
    w2 = torch.zeros(((n + 7) // 8 * 8, Cc), device=weight.device, dtype=torch.float32)
    w2[:n] = weight.detach().float().permute(2, 3, 0, 1).reshape(n, Cc)
    return pack_conv(w2, None)


def conv_tapsum(P: torch.Tensor, bias: Optional[torch.Tensor], C_out: int, k: int, pad: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """P (B,H,W,>=k*k*C) from the pointwise layer of pack_conv_tapcols -> the k x k conv's output (B,H,W,C) (`out` may be a column slice)."""
    lib = _lib.load()
    B, H, W, Cp, bs, ld = _nhwc(P)
    assert bs == H * W * ld
    if out is None:
        out = torch.empty((B, H, W, C_out), device=P.device, dtype=torch.float32)
    oB, oH, oW, oC, obs, old = _nhwc(out)
    assert (oB, oH, oW, oC) == (B, H, W, C_out) and obs == H * W * old
    check(lib.sma_conv_tapsum(P.data_ptr(), ld, B, H, W, C_out, k, pad, _ptr(bias), out.data_ptr(), old, _stream()), 'sma_conv_tapsum')
    return out


def flow_update(m_prev: torch.Tensor, occ_prev: torch.Tensor, res: torch.Tensor):
    lib = _lib.load()
    B, h, w, _ = m_prev.shape
    assert m_prev.is_contiguous() and occ_prev.is_contiguous() and res.is_contiguous()
    m_new, occ_new = torch.empty_like(m_prev), torch.empty_like(occ_prev)
    check(lib.sma_flow_update(m_prev.data_ptr(), occ_prev.data_ptr(), res.data_ptr(), res.shape[-1], B, h, w, m_new.data_ptr(),
                              occ_new.data_ptr(), _stream()), 'sma_flow_update')
    return m_new, occ_new


def motion_ignore_mask(m: torch.Tensor, size=(32, 32)) -> torch.Tensor:
    lib = _lib.load()
    B, h, w, _ = m.shape
    mask = torch.empty((B, size[0] * size[1]), device=m.device, dtype=torch.uint8)
    check(lib.sma_motion_ignore_mask(m.data_ptr(), B, h, w, size[0], size[1], mask.data_ptr(), _stream()), 'sma_motion_ignore_mask')
    return mask


def sft_combine(dec: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, w: float, out: Optional[torch.Tensor] = None):
    lib = _lib.load()
    assert dec.is_contiguous() and scale.is_contiguous() and shift.is_contiguous()
    if out is None:
        out = torch.empty_like(dec)
    check(lib.sma_sft_combine(dec.data_ptr(), scale.data_ptr(), shift.data_ptr(), w, dec.numel(), out.data_ptr(), _stream()), 'sma_sft_combine')
    return out


def to_uint8(x: torch.Tensor, bgr: bool = False) -> torch.Tensor:
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(x)
    out = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.uint8)
    check(lib.sma_to_uint8(x.data_ptr(), B, H, W, Cc, ld, 1 if bgr else 0, out.data_ptr(), _stream()), 'sma_to_uint8')
    return out


def nchw_to_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), device=x.device, dtype=torch.float32)
    _, _, _, _, obs, old = _nhwc(out)
    check(lib.sma_nchw_to_nhwc(x.data_ptr(), B, Cc, H, W, out.data_ptr(), old, _stream()), 'sma_nchw_to_nhwc')
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    B, H, W, Cc, bs, ld = _nhwc(x)
    out = torch.empty((B, Cc, H, W), device=x.device, dtype=torch.float32)
    check(lib.sma_nhwc_to_nchw(x.data_ptr(), B, Cc, H, W, ld, out.data_ptr(), _stream()), 'sma_nhwc_to_nchw')
    return out


def launch_count() -> int:
    return _lib.load().sma_kernel_launch_count()


# ---------------------------------------------------------------------------------------------------------------------
# The library acts on the CURRENT device (include/sma_b200.h); run every op on the device that owns its first tensor argument, so that a
# process driving several GPUs never launches on the wrong one.
# ---------------------------------------------------------------------------------------------------------------------
import functools as _functools


def _on_device(fn):
    @_functools.wraps(fn)
    def wrap(x, *a, **k):
        t = x.w if isinstance(x, ConvW) else (x[0] if isinstance(x, (list, tuple)) and x else x)
        if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(x, *a, **k)
        return fn(x, *a, **k)
    return wrap


for _name in ('pack_conv', 'pack_conv_cat', 'pack_conv_blockdiag', 'conv2d', 'linear', 'groupnorm_stats', 'affine_act', 'layernorm', 'warp_occlude',
              'resize_ac', 'mha', 'attn256', 'vq_lookup', 'antialias_down4', 'avgpool2', 'kp_head', 'normalize_kp', 'hull_scale', 'u8hwc_to_f32nchw',
              'dense_motion_prep', 'dense_motion_head', 'flow_to_px', 'flow_update', 'motion_ignore_mask', 'sft_combine', 'to_uint8', 'nchw_to_nhwc',
              'nhwc_to_nchw'):
    globals()[_name] = _on_device(globals()[_name])
