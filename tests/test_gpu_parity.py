"""GPU parity: every kernel behind the C ABI against the CPU oracle (oracle/sma_oracle.py) and against the fixtures
produced by the live reference (tests/golden/, written by oracle/make_golden.py).  All calls go through
`sma_b200.ops` -> ctypes -> include/sma_b200.h.  Tolerances: 1e-3 max-abs on the final image (north star),
tighter per stage; bit-exact for indices, masks and uint8 conversion.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import sma_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def S():
    import sma_b200 as S
    assert torch.cuda.is_available()
    assert S._lib.load().sma_device_check(0) == 0, 'not an sm_100 device'
    return S


def nhwc(x):   # NCHW cpu -> NHWC cuda contiguous
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):   # NHWC cuda -> NCHW cpu
    return x.permute(0, 3, 1, 2).contiguous().cpu()


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


# ---------------------------------------------------------------------------------------------------
# convolution (both the exact-fp32 CUDA-core kernel and, where the shape allows, the tcgen05 3xTF32 one)
# ---------------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, Cin, H, W, Cout, k, stride, pad, extras
    (2, 3, 32, 32, 64, 3, 1, 1, {}),
    (1, 64, 24, 20, 64, 3, 1, 1, {}),
    (2, 128, 16, 16, 256, 3, 1, 1, {'act': 'relu'}),
    (1, 256, 32, 32, 256, 3, 1, 1, {'res': True}),
    (2, 256, 32, 32, 768, 1, 1, 0, {}),
    (1, 35, 64, 64, 75, 7, 1, 0, {}),
    (1, 128, 64, 64, 17, 7, 1, 3, {}),
    (2, 2, 64, 64, 128, 7, 1, 3, {'act': 'relu'}),
    (1, 32, 64, 64, 32, 3, 2, 0, {'down': True}),
    (2, 128, 16, 16, 128, 3, 1, 1, {'up': True, 'act': 'relu'}),
    (1, 64, 32, 32, 3, 3, 1, 1, {'pre': 'none', 'nchw': True}),
    (2, 64, 16, 16, 128, 3, 1, 1, {'pre': 'swish', 'res': True}),
    (1, 160, 64, 64, 126, 3, 1, 1, {'act': 'relu'}),
    (1, 256, 32, 32, 512, 3, 1, 1, {'act': 'gelu'}),
    (1, 128, 32, 32, 256, 3, 1, 1, {'act': 'leaky'}),
    (3, 15, 32, 32, 32, 1, 1, 0, {'act': 'relu'}),
    (2, 64, 40, 24, 192, 1, 1, 0, {'act': 'relu'}),
    (1, 192, 64, 64, 128, 3, 1, 1, {'act': 'relu', 'res': True}),
    (2, 128, 32, 32, 64, 1, 1, 0, {'res': True}),
    (1, 256, 64, 64, 3, 3, 1, 1, {}),
    (2, 64, 45, 27, 64, 3, 1, 1, {'pre': 'swish', 'res': True}),
    (1, 64, 64, 64, 48, 3, 1, 1, {'act': 'leaky'}),
    (2, 256, 32, 32, 512, 1, 1, 0, {'act': 'gelu'}),
]


def _act_ref(x, act):
    return {'none': lambda v: v, 'relu': F.relu, 'leaky': lambda v: F.leaky_relu(v, 0.2), 'gelu': F.gelu,
            'sigmoid': torch.sigmoid, 'swish': lambda v: v * torch.sigmoid(v)}[act](x)


@pytest.mark.parametrize('mode', ['exact', 'default', 'ts', 'ts-stream', 'ts-stream128', 'f16x3', 'f16x3-3mma', 'tf32x3', 'gather'])
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d_matches_torch(S, case, mode):
    """exact: fp32 FFMA kernel; default: the library's choice; ts: fp16-split tcgen05 kernel with the weights as the tensor-memory operand
    wherever Cin % 64 == 0 and stride 1 (resident where they fit, streamed over 256-pixel tiles otherwise); ts-stream: the same, always streamed; ts-stream128: streamed over
    128-pixel tiles with two accumulators; f16x3-3mma: the halo kernel with three separate MMAs per k-step instead of the fused
    [hi | lo] weight tile; f16x3:
    fp16-split halo kernel with both operands in shared memory; tf32x3: tf32-split tcgen05 kernels; gather: force the non-persistent
    tcgen05 kernel."""
    B, Cin, H, W, Cout, k, stride, pad, ex = case
    x = rnd(B, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, k, k, seed=2, scale=(Cin * k * k) ** -0.5)
    b = rnd(Cout, seed=3, scale=0.1)
    xr = x.double()
    pre = None
    if 'pre' in ex:
        sc, sh = 1 + 0.1 * rnd(B, Cin, seed=4), 0.1 * rnd(B, Cin, seed=5)
        xr = _act_ref(xr * sc.double().view(B, Cin, 1, 1) + sh.double().view(B, Cin, 1, 1), ex['pre'])
        pre = (sc.cuda(), sh.cuda(), ex['pre'])
    kw = {}
    if ex.get('up'):
        xr = F.interpolate(xr, scale_factor=2, mode='nearest')
        kw['upsample2'] = True
    if ex.get('down'):
        xr = F.pad(xr, (0, 1, 0, 1))
        kw.update(pad_tl=(0, 0), out_hw=(H // 2, W // 2))
    ref = F.conv2d(xr, w.double(), b.double(), stride=stride, padding=pad)
    ref = _act_ref(ref, ex.get('act', 'none'))
    res = None
    if ex.get('res'):
        r = rnd(*ref.shape, seed=6)
        ref = ref + r.double()
        res = nhwc(r)
    cw = S.ops.pack_conv(w.cuda(), b.cuda())
    saved = (S.ops.TC_VARIANT, S.ops.USE_F16, S.ops.USE_TS)
    S.ops.TC_VARIANT = {'gather': 1, 'ts': 32, 'ts-stream': 48, 'ts-stream128': 112, 'f16x3-3mma': 128}.get(mode, 0)
    S.ops.USE_F16 = mode in ('f16x3', 'f16x3-3mma', 'ts', 'ts-stream', 'ts-stream128', 'default')
    S.ops.USE_TS = mode in ('ts', 'ts-stream', 'ts-stream128', 'default')
    try:
        y = S.ops.conv2d(nhwc(x), cw, stride=stride, pad=pad, act=ex.get('act', 'none'), pre=pre, res=res,
                         out_nchw=bool(ex.get('nchw')), exact=mode == 'exact', **kw)
    finally:
        S.ops.TC_VARIANT, S.ops.USE_F16, S.ops.USE_TS = saved
    got = y.cpu() if ex.get('nchw') else nchw(y)
    err = float((got.double() - ref).abs().max())
    assert got.shape == ref.shape
    tc_eligible = mode != 'exact' and Cin % 32 == 0 and not ex.get('nchw')
    assert (S.ops.LAST_CONV_KERNEL >= 1) == tc_eligible, S.ops.LAST_CONV_KERNEL      # the tensor-core kernels really ran
    if mode in ('f16x3', 'f16x3-3mma', 'ts', 'ts-stream', 'ts-stream128') and tc_eligible and Cin % 64 == 0 and stride == 1:
        # ... and the fp16-split ones where eligible (7x7 halos leave the tensor-memory-operand kernel fewer than two stages: it declines)
        assert S.ops.LAST_CONV_KERNEL in ((3,) if mode.startswith('f16x3') else ((3, 4) if k == 7 else (4,))), S.ops.LAST_CONV_KERNEL
    # exact kernel: fp32 FFMA; split kernels: the TMEM accumulator adds with truncation, error grows ~1e-8 * K
    tol = (2e-5 + (1e-8 * Cin * k * k if tc_eligible else 0.0)) * max(1.0, float(ref.abs().max()))
    assert err < tol, (err, tol)


@pytest.mark.parametrize('scale_w,scale_x,tol', [(1e-3, 1.0, 2e-5), (30.0, 1.0, 2e-5), (0.05, 1e-3, 1e-4), (0.05, 300.0, 2e-5)])
def test_conv2d_f16_split_dynamic_range(S, scale_w, scale_x, tol):
    """fp16-split kernel away from unit scale: per-channel power-of-two weight scaling keeps tiny / large weights exact to ~2^-22;
    small activations lose their lo half below 2^-25 absolute (still ~1e-7 of the output scale), large ones stay finite."""
    B, Cin, H, W, Cout, k = 1, 128, 24, 24, 64, 3
    x = rnd(B, Cin, H, W, seed=1) * scale_x
    w = rnd(Cout, Cin, k, k, seed=2, scale=(Cin * k * k) ** -0.5) * scale_w
    w[3] *= 1e-4; w[5] *= 50.0                                       # per-channel spread
    ref = F.conv2d(x.double(), w.double(), None, padding=1)
    y = S.ops.conv2d(nhwc(x), S.ops.pack_conv(w.cuda(), None), pad=1)
    assert S.ops.LAST_CONV_KERNEL in (3, 4)                         # one of the fp16-split kernels
    err = (nchw(y).double() - ref).abs().amax(dim=(0, 2, 3))
    mag = ref.abs().amax(dim=(0, 2, 3)).clamp_min(1e-30)
    assert float((err / mag).max()) < tol, (err / mag).max()        # measured: 2.5e-5 for the 1e-3-scale activations, < 1e-5 otherwise


def test_conv2d_concat_slices_patchify_and_bn_fold(S):
    """channel-slice views as input/output (torch.cat elimination), stride-p patch embedding, depth-to-space."""
    B, C, s, p = 2, 128, 64, 2
    x = rnd(B, C, s, s, seed=1)
    Wl, bl = rnd(256, C * p * p, seed=2, scale=0.05), rnd(256, seed=3, scale=0.1)
    # reference patchify: (p1 p2 c) feature order, appmotioncodebook_arch.py:222
    xp = x.view(B, C, 32, p, 32, p).permute(0, 2, 4, 3, 5, 1).reshape(B, 1024, p * p * C)
    ref = F.linear(xp.double(), Wl.double(), bl.double())
    cw = S.ops.pack_conv(Wl.cuda(), bl.cuda()).as_patch(p)
    buf = torch.zeros(B, s, s, C + 32, device='cuda')            # input is a channel slice of a wider buffer
    buf[..., 16:16 + C] = nhwc(x)
    tok = S.ops.conv2d(buf[..., 16:16 + C], cw, stride=p).view(B, 1024, 256)
    assert float((tok.cpu().double() - ref).abs().max()) < 1e-4
    # inverse: Linear(256 -> C*p*p) + un-patchify
    Wi, bi = rnd(C * p * p, 256, seed=4, scale=0.05), rnd(C * p * p, seed=5, scale=0.1)
    y = F.linear(ref.float(), Wi, bi)
    ref2 = y.view(B, 32, 32, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, s, s)
    out = torch.zeros(B, s, s, 2 * C, device='cuda')
    S.ops.conv2d(ref.float().cuda().view(B, 32, 32, 256), S.ops.pack_conv(Wi.cuda(), bi.cuda()), d2s=p, out=out[..., :C])
    assert float((nchw(out[..., :C]) - ref2).abs().max()) < 1e-4
    assert float(out[..., C:].abs().max()) == 0.0
    # BatchNorm(eval) fold
    w, b = rnd(64, 32, 3, 3, seed=6, scale=0.1), rnd(64, seed=7, scale=0.1)
    bn = {'weight': 1 + 0.1 * rnd(64, seed=8), 'bias': 0.1 * rnd(64, seed=9), 'running_mean': 0.1 * rnd(64, seed=10),
          'running_var': torch.rand(64, generator=torch.Generator().manual_seed(11)) + 0.5}
    xx = rnd(2, 32, 16, 16, seed=12)
    refb = F.relu(F.batch_norm(F.conv2d(xx, w, b, padding=1), bn['running_mean'], bn['running_var'], bn['weight'], bn['bias'],
                               False, 0.1, 1e-5))
    cwb = S.ops.pack_conv(w.cuda(), b.cuda(), {k: v.cuda() for k, v in bn.items()})
    got = nchw(S.ops.conv2d(nhwc(xx), cwb, pad=1, act='relu'))
    assert float((got - refb).abs().max()) < 2e-5


def test_conv2d_bad_arguments_return_status(S):
    lib = S._lib.load()
    d = S._lib.ConvDesc()
    assert lib.sma_conv2d_fwd(ctypes.byref(d), None) == -1          # null pointers -> SMA_ERR_BAD_ARG
    assert lib.sma_conv2d_fwd(None, None) == -1
    assert lib.sma_vq_lookup_fwd(None, 0, 0, None, 0, None, None, None, None) == -1
    assert lib.sma_warp_occlude_fwd(None, 0, 1, 4, 4, 4, None, None, 4, 4, None, None) == -1
    z = torch.zeros(8, 48, device='cuda'); cb = torch.zeros(16, 48, device='cuda')
    with pytest.raises(RuntimeError):                               # unsupported embedding width -> status -2 -> raise
        S.ops.vq_lookup(z, cb)
    with pytest.raises(RuntimeError):                               # channel mismatch
        S.ops.conv2d(torch.zeros(1, 8, 8, 16, device='cuda'), S.ops.pack_conv(torch.zeros(8, 32, 3, 3, device='cuda'), None))
    assert lib.sma_status_string(-2) == b'unsupported shape'


# ---------------------------------------------------------------------------------------------------
# norms
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(2, 64, 64, 64), (1, 256, 32, 32), (3, 32, 32, 32), (1, 128, 128, 128)])
def test_groupnorm_prologue(S, shape):
    B, C, H, W = shape
    x = rnd(*shape, seed=1) * 2 + 0.5
    g, b = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    ref = F.group_norm(x.double(), 32, g.double(), b.double(), eps=1e-6)
    sc, sh = S.ops.groupnorm_stats(nhwc(x), g.cuda(), b.cuda(), 32, 1e-6)
    got = nchw(S.ops.affine_act(nhwc(x), sc, sh, 'none'))
    assert float((got.double() - ref).abs().max()) < 2e-5
    got = nchw(S.ops.affine_act(nhwc(x), sc, sh, 'swish'))
    assert float((got.double() - ref * torch.sigmoid(ref)).abs().max()) < 2e-5


@pytest.mark.parametrize('E', [32, 256])
def test_layernorm(S, E):
    x = rnd(2, 1024, E, seed=1) * 3
    g, b, pos = 1 + 0.1 * rnd(E, seed=2), 0.1 * rnd(E, seed=3), rnd(1024, E, seed=4, scale=0.02)
    ref = F.layer_norm(x.double(), (E,), g.double(), b.double())
    y, yq = S.ops.layernorm(x.cuda(), g.cuda(), b.cuda(), pos.cuda())
    assert float((y.cpu().double() - ref).abs().max()) < 1e-5
    assert float((yq.cpu().double() - (ref + pos.double())).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------
# stage 2: warp + occlude, resize
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('C,s', [(256, 32), (128, 64), (128, 128), (64, 256)])
def test_warp_occlude_matches_oracle(S, C, s):
    B = 2
    feat = rnd(1, C, s, s, seed=1)
    base = O.coord_grid(64, 64).unsqueeze(0)
    flow = base + 0.3 * rnd(B, 64, 64, 2, seed=2)                  # some samples fall outside [-1,1] -> zero padding
    occ = torch.rand(B, 1, 64, 64, generator=torch.Generator().manual_seed(3))
    ref0 = O.warp_ac(feat, flow)
    ref1 = O.occlude(ref0, occ)
    f = nhwc(feat).expand(B, -1, -1, -1)                           # shared source: batch stride 0
    got0 = nchw(S.ops.warp_occlude(f, flow.cuda(), None))
    got1 = nchw(S.ops.warp_occlude(f, flow.cuda(), occ.view(B, 64, 64).cuda()))
    # the sampling position carries ~1 ulp of the resized flow times (s-1)/2 pixels; feature gradients are O(1) per pixel
    tol = 2e-6 * s
    assert float((got0 - ref0).abs().max()) < tol
    assert float((got1 - ref1).abs().max()) < tol


def test_warp_identity_and_linearity_full_size(S):
    """Size-independent properties at the BASELINE batch (64 frames): identity flow reproduces the source bit-for-bit,
    out-of-range flow gives zeros, and the warp is linear in the features."""
    B, C, s = 64, 128, 64
    feat = rnd(1, s, s, C, seed=1).cuda()
    ident = O.coord_grid(64, 64).unsqueeze(0).expand(B, -1, -1, -1).contiguous().cuda()
    out = S.ops.warp_occlude(feat.expand(B, -1, -1, -1), ident, None)
    assert float((out - feat).abs().max()) < 5e-5      # grid coordinates 2*i/63-1 are not exactly representable
    far = torch.full((B, 64, 64, 2), 3.0, device='cuda')
    assert float(S.ops.warp_occlude(feat.expand(B, -1, -1, -1), far, None).abs().max()) == 0.0
    flow = (ident + 0.2 * rnd(B, 64, 64, 2, seed=2).cuda()).contiguous()
    f2 = rnd(1, s, s, C, seed=3).cuda()
    a = S.ops.warp_occlude(feat.expand(B, -1, -1, -1), flow, None)
    b = S.ops.warp_occlude(f2.expand(B, -1, -1, -1), flow, None)
    ab = S.ops.warp_occlude((feat + 2 * f2).expand(B, -1, -1, -1), flow, None)
    assert float((ab - (a + 2 * b)).abs().max()) < 1e-4


@pytest.mark.parametrize('C,si,so', [(15, 64, 32), (32, 32, 64), (192, 128, 64), (64, 256, 32), (2, 64, 256)])
def test_resize_bilinear_align_corners(S, C, si, so):
    x = rnd(2, C, si, si, seed=1)
    ref = F.interpolate(x, size=(so, so), mode='bilinear', align_corners=True)
    got = nchw(S.ops.resize_ac(nhwc(x), (so, so)))
    assert float((got - ref).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------
# stage 3: attention core, VQ lookup
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('E,heads,S_kv,shared,masked', [(256, 8, 1024, False, True), (256, 8, 256, True, False), (256, 8, 768, True, False),
                                                        (32, 8, 1024, False, False), (32, 8, 512, True, False),
                                                        (256, 1, 1024, False, False)])
@pytest.mark.parametrize('exact', [True, False])
def test_mha_core(S, E, heads, S_kv, shared, masked, exact):
    B, L = 2, 1024
    D = E // heads
    q = rnd(B, L, E, seed=1)
    k = rnd(S_kv, E, seed=2) if shared else rnd(B, S_kv, E, seed=2)
    v = rnd(S_kv, E, seed=3) if shared else rnd(B, S_kv, E, seed=3)
    mask = None
    if masked:
        mask = torch.rand(B, S_kv, generator=torch.Generator().manual_seed(4)) < 0.1
    kk = k.unsqueeze(0).expand(B, -1, -1) if shared else k
    vv = v.unsqueeze(0).expand(B, -1, -1) if shared else v
    qh = q.double().view(B, L, heads, D).transpose(1, 2) * (D ** -0.5)
    kh = kk.double().reshape(B, S_kv, heads, D).transpose(1, 2)
    vh = vv.double().reshape(B, S_kv, heads, D).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if mask is not None:
        s = s.masked_fill(mask.view(B, 1, 1, S_kv), float('-inf'))
    ref = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, L, E)
    got = S.ops.mha(q.cuda(), k.cuda(), v.cuda(), heads, None if mask is None else mask.to(torch.uint8).cuda(), exact=exact)
    assert float((got.cpu().double() - ref).abs().max()) < (2e-5 if exact else 1e-4)


@pytest.mark.parametrize('B,L,S_kv,qs', [(3, 1024, 1024, 1.0), (1, 256, 512, 1.0), (2, 1024, 1024, 6.0)])
def test_attn256_tensor_core_kernel(S, B, L, S_kv, qs):
    """AttnBlock attention on tcgen05 (csrc/attn256.cu) called directly; qs = 6 makes the scores spread over ~+-100 so that the lazy
    reference maximum is exceeded and the O rescale in tensor memory runs; q, k, v are column slices of one buffer like the qkv conv output."""
    qkv = rnd(B, max(L, S_kv), 768, seed=7)
    qkv[..., :256] *= qs
    q, k, v = qkv[:, :L, :256], qkv[:, :S_kv, 256:512], qkv[:, :S_kv, 512:]
    s = (q.double() @ k.double().transpose(-1, -2)) * (256 ** -0.5)
    ref = torch.softmax(s, -1) @ v.double()
    dev = qkv.cuda()
    n0 = S.ops.launch_count()
    got = S.ops.attn256(dev[:, :L, :256], dev[:, :S_kv, 256:512], dev[:, :S_kv, 512:], 256 ** -0.5)
    assert S.ops.launch_count() - n0 == 2                            # split + attention kernels
    assert float((got.cpu().double() - ref).abs().max()) < 1e-4
    with pytest.raises(RuntimeError):                                # L not a multiple of 128 -> status -2
        S.ops.attn256(dev[:, :100, :256], dev[:, :S_kv, 256:512], dev[:, :S_kv, 512:], 1.0)


def test_mha_all_keys_masked_gives_nan_like_reference(S):
    q, k, v = rnd(1, 1024, 256, seed=1).cuda(), rnd(1, 1024, 256, seed=2).cuda(), rnd(1, 1024, 256, seed=3).cuda()
    mask = torch.ones(1, 1024, dtype=torch.uint8, device='cuda')
    assert bool(torch.isnan(S.ops.mha(q, k, v, 8, mask)).all())


@pytest.mark.parametrize('E', [256, 32])
@pytest.mark.parametrize('init', ['normal', 'tiny'])
def test_vq_lookup_indices_bit_exact(S, E, init):
    g = torch.Generator().manual_seed(7)
    cb = torch.randn(1024, E, generator=g) if init == 'normal' else (torch.rand(1024, E, generator=g) * 2 - 1) / 1024
    z = torch.randn(8, E, 32, 32, generator=g) * (1.0 if init == 'normal' else 0.3)
    zf = z.permute(0, 2, 3, 1).reshape(-1, E).contiguous()
    for scale in (None, 0.25, 0.5, 0.75):
        zq_r, _, idx_r, _, _ = O.vq_lookup(cb, z, scale)
        n = 1024 if scale is None else int(scale * 1024)
        idx, zq, md = S.ops.vq_lookup(zf.cuda(), cb.cuda(), n)
        idx = idx.cpu()
        mism = int((idx != idx_r[:, 0]).sum())
        if mism:
            # a different summation order may move a row across an exact fp32 tie; the distances must still be
            # equal to the last bit of the reference's own formula for the index we picked
            d = (zf ** 2).sum(1, keepdim=True) + (cb[:n] ** 2).sum(1) - 2 * zf @ cb[:n].t()
            bad = (idx != idx_r[:, 0]).nonzero()[:, 0]
            assert mism <= 2 and bool((d[bad, idx[bad]] - d[bad, idx_r[bad, 0]]).abs().max() <= 4e-6 * d[bad].abs().max()), mism
        assert torch.equal(zq.cpu(), cb[idx])
        # idempotence: quantising code vectors returns the same codes
        idx2, _, _ = S.ops.vq_lookup(zq, cb.cuda(), n)
        assert torch.equal(idx2.cpu(), idx)


def test_vq_lookup_ties_lowest_index_and_ragged(S):
    g = torch.Generator().manual_seed(3)
    cb = torch.randn(16, 32, generator=g)
    cb[9] = cb[4]
    z = cb[4].view(1, 32).expand(13, 32).contiguous()          # ragged N (not a multiple of the 8-row block)
    idx, zq, md = S.ops.vq_lookup(z.cuda(), cb.cuda(), 16)
    assert idx.dtype == torch.int64 and idx.tolist() == [4] * 13
    idx, _, _ = S.ops.vq_lookup(z[:1].cuda(), cb.cuda(), 1)    # single code
    assert idx.tolist() == [0]


# ---------------------------------------------------------------------------------------------------
# motion estimator (KP detector, normalize_kp, dense motion) against the oracle and the reference fixtures
# ---------------------------------------------------------------------------------------------------
def test_kp_detector_and_dense_motion(S, nets, weights, clip, golden):
    g, me = nets
    P_g, P_me = weights
    src, drv = clip
    s1, d1, d0 = src.unsqueeze(0).cuda(), drv[1].unsqueeze(0).cuda(), drv[0].unsqueeze(0).cuda()
    kp_s, kp_d, kp_0 = me.estimate_kp(s1), me.estimate_kp(d1), me.estimate_kp(d0)
    assert float((kp_s['value'].cpu() - golden['kp_source_value']).abs().max()) < 1e-4
    assert float((kp_s['jacobian'].cpu() - golden['kp_source_jacobian']).abs().max()) < 1e-4
    assert float((kp_d['value'].cpu() - golden['kp_driving1_value']).abs().max()) < 1e-4
    assert float((kp_d['jacobian'].cpu() - golden['kp_driving1_jacobian']).abs().max()) < 1e-4
    # batched == per-frame
    kp_b = me.estimate_kp(torch.cat([d0, d1]))
    assert float((kp_b['value'][1] - kp_d['value'][0]).abs().max()) < 1e-5
    kpn = S.normalize_kp(kp_s, kp_d, kp_0, adapt_movement_scale=True, use_relative_movement=True, use_relative_jacobian=True)
    assert float((kpn['value'].cpu() - golden['kp_norm1_value']).abs().max()) < 2e-4
    assert float((kpn['jacobian'].cpu() - golden['kp_norm1_jacobian']).abs().max()) < 2e-3
    # dense motion from the reference's own normalised keypoints (isolates this stage)
    kpn_ref = {'value': golden['kp_norm1_value'].cuda(), 'jacobian': golden['kp_norm1_jacobian'].cuda()}
    kps_ref = {'value': golden['kp_source_value'].cuda(), 'jacobian': golden['kp_source_jacobian'].cuda()}
    dm = me.estimate_motion_w_kp(kp_source=kps_ref, kp_driving=kpn_ref, source_image=s1)
    assert float((dm['deformation'].cpu() - golden['deformation1']).abs().max()) < 1e-4
    assert float((dm['occlusion_map'].cpu() - golden['occlusion1']).abs().max()) < 1e-4
    assert float((dm['driving_kp_heatmap'].cpu()[:, :, ::4, ::4] - golden['driving_kp_heatmap1_s4']).abs().max()) < 1e-5
    assert set(dm) >= {'deformation', 'occlusion_map', 'driving_kp_heatmap', 'kp_driving', 'kp_source'}


def test_encoder_features(S, nets, weights, clip, golden):
    g, me = nets
    src, _ = clip
    feats = g.encode_source(src.unsqueeze(0).cuda())
    ref = O.encode_source(weights[0], src.unsqueeze(0))
    for s in (256, 128, 64, 32):
        got = nchw(feats[s])
        assert float((got - ref[str(s)]).abs().max()) < 2e-4 * max(1.0, float(ref[str(s)].abs().max())), s
    assert float((nchw(feats[32])[:, ::4, ::4, ::4] - golden['enc_feat32_s4']).abs().max()) < 5e-4
    nch = g.encode_driving(src.unsqueeze(0).cuda())
    assert set(nch) == {'256', '128', '64', '32'} and tuple(nch['32'].shape) == (1, 256, 32, 32)


def test_generator_forward_matches_reference_fixture(S, nets, golden, clip):
    """The north-star check: `net_g(source, dense_motion, w=1, inference=True)['out']` within 1e-3 max-abs of the
    reference's fp32 forward on identical tensors."""
    g, me = nets
    src, drv = clip
    dm = {'deformation': golden['deformation1'].cuda(), 'occlusion_map': golden['occlusion1'].cuda(),
          'driving_kp_heatmap': None}
    # the fixture stores the heat-map subsampled; rebuild it from the reference key-points (kp2gaussian)
    dm['driving_kp_heatmap'] = O.gaussian_heatmaps(golden['kp_norm1_value'], 64, 64).cuda()
    out = g(src.unsqueeze(0).cuda(), dm, w=1, inference=True)
    err = float((out['out'].cpu() - golden['out1']).abs().max())
    assert err < 1e-3, err
    for a, b in zip(out['out_occ'], golden['out_occ1']):
        assert float((a.cpu() - b).abs().max()) < 1e-4
    for a, b in zip(out['deformation_list'], golden['deformation_list1']):
        assert float((a.cpu() - b).abs().max()) < 1e-4
    assert float((out['lq_feat'].cpu()[:, ::4, ::4, ::4] - golden['lq_feat1_s4']).abs().max()) < 5e-4
    with pytest.raises(NotImplementedError):
        g(src.unsqueeze(0).cuda(), dm, w=1, inference=False)


def test_make_animation_matches_reference_clip(S, nets, golden, clip):
    g, me = nets
    src, drv = clip
    for batch in (1, 3):
        preds, drvs = S.make_animation(src, drv, g, me, relative=True, adapt_movement_scale=True, batch=batch)
        assert len(preds) == 3 and preds[0].shape == (256, 256, 3) and preds[0].dtype == np.uint8
        for p, r in zip(preds, golden['pred_uint8']):
            d = np.abs(p.astype(int) - r.numpy().astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 0.01, (batch, d.max(), (d > 0).mean())
        for d_, f in zip(drvs, drv):
            assert np.array_equal(d_, O.to_uint8(f))
    with pytest.raises(RuntimeError):
        S.make_animation(src, drv, g, me, cpu=True)
    # absolute (non-relative) mode and BGR twin
    pa, _ = S.make_animation(src, drv[:1], g, me, relative=False, adapt_movement_scale=False, batch=1)
    ra, _, _ = O.make_animation(*_weights_of(nets), src, drv[:1], False, False)
    assert np.abs(pa[0].astype(int) - ra[0].astype(int)).max() <= 1
    pb, _ = S.make_animation_model({'val': {'relative': True, 'adapt_scale': True, 'w': 1}}, g, me, src.unsqueeze(0),
                                   [f.unsqueeze(0) for f in drv[:1]])
    assert np.abs(pb[0][:, :, ::-1].astype(int) - golden['pred_uint8'][0].numpy().astype(int)).max() <= 1


def _weights_of(nets):
    g, me = nets
    return ({k: v.detach().cpu() for k, v in g.state_dict().items()}, {k: v.detach().cpu() for k, v in me.state_dict().items()})


def test_full_batch_is_frame_independent(S, nets):
    """BASELINE configs[1] size (64 driving frames): the batched path equals the per-frame path and a permutation of
    the driving frames permutes the output (no cross-frame state)."""
    g, me = nets
    src, drv = O.synthetic_frames(64, seed=99)
    p64, _ = S.make_animation(src, drv, g, me, batch=16)
    idx = [5, 40, 63]
    p1, _ = S.make_animation(src, [drv[0]] + [drv[i] for i in idx], g, me, batch=1)   # frame 0 fixes kp_driving_initial
    for j, i in enumerate(idx):
        assert np.abs(p64[i].astype(int) - p1[j + 1].astype(int)).max() <= 1
    assert all(np.isfinite(p.astype(np.float32)).all() for p in p64)


def test_caller_surface_plain_decoder_and_estimator_forward(S, nets, clip):
    """what AppMotionCompModel.test calls besides the forward (models/appmotioncomp_model.py:437-456):
    `motion_estimator(driving, source)` and `net_g.generator(lq_feat)`, against the fixtures of the live reference."""
    import os
    from conftest import GOLD
    fx = torch.load(os.path.join(GOLD, 'reference_callers.pt'))
    g, me = nets
    src, drv = clip
    lq = torch.randn(1, 256, 32, 32, generator=torch.Generator().manual_seed(fx['lq_seed'])) * 0.5
    recon = g.generator(lq.cuda()).cpu()
    assert recon.shape == (1, 3, 256, 256)
    assert float((recon[:, :, ::2, ::2] - fx['recon_s2']).abs().max()) < 1e-3
    dm = me(drv[1].unsqueeze(0).cuda(), src.unsqueeze(0).cuda())
    assert float((dm['deformation'].cpu() - fx['fwd_deformation']).abs().max()) < 1e-4
    assert float((dm['occlusion_map'].cpu() - fx['fwd_occlusion']).abs().max()) < 1e-4
    assert float((dm['kp_driving']['value'].cpu() - fx['fwd_kp_driving_value']).abs().max()) < 1e-4
    sd = g.state_dict()                                   # the callable container does not disturb the key inventory
    assert 'generator.blocks.0.weight' in sd and not any(k.startswith('generator._') for k in sd)


def test_to_uint8_round_half_even_bit_exact(S):
    x = torch.linspace(-1.2, 1.2, 256 * 256 * 3).view(1, 256, 256, 3)
    x[0, 0, 0, 0] = 1.0 / 255.0 * 2 * 0.5 - 1      # (v+1)/2*255 = 0.5 -> rounds to 0 (half to even)
    x[0, 0, 0, 1] = 1.0 / 255.0 * 2 * 1.5 - 1      # 1.5 -> 2
    got = S.ops.to_uint8(x.cuda(), False).cpu().numpy()[0]
    ref = O.to_uint8(x[0].permute(2, 0, 1))
    assert np.array_equal(got, ref)
    got = S.ops.to_uint8(x.cuda(), True).cpu().numpy()[0]
    assert np.array_equal(got, O.to_uint8(x[0].permute(2, 0, 1), bgr=True))


def test_launch_counter_counts_kernels(S):
    n0 = S.ops.launch_count()
    S.ops.to_uint8(torch.zeros(1, 4, 4, 3, device='cuda'))
    assert S.ops.launch_count() == n0 + 1
