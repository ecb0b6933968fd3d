"""CPU oracle for the per-driving-frame talking-head hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

A from-scratch functional restatement, in plain fp32 torch ops on the CPU, of what the reference
computes between `demo.make_animation` and the uint8 frame.  It exists only to check the CUDA path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it.  The product path (synergize-motion-appearance_b200/) never imports anything under oracle/.

Parity status: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against the *live reference modules* imported in the build container
(oracle/make_golden.py: identical seeded weights + inputs, outputs compared, fixtures written to
tests/golden/).  Everything here works on a flat {reference state_dict key: tensor} mapping, NCHW.

Reference citations are relative to /root/reference/basicsr/.
"""
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

T = torch.Tensor
SD = Dict[str, T]

# ----------------------------------------------------------------------------------------------
# small shared pieces
# ----------------------------------------------------------------------------------------------


def coord_grid(h: int, w: int, dtype=torch.float32, device=None) -> T:
    """(h,w,2) grid, x then y, 2*i/(n-1)-1.  utils/motion_estimator_util.py:56-72."""
    xs = 2 * (torch.arange(w, dtype=dtype, device=device) / (w - 1)) - 1
    ys = 2 * (torch.arange(h, dtype=dtype, device=device) / (h - 1)) - 1
    return torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=-1)


def resize_ac(x: T, size) -> T:
    """bilinear, align_corners=True (archs/appmotioncodebook_arch.py:354,360,390,414,418,488,571,671)."""
    if tuple(x.shape[-2:]) == tuple(size):
        return x
    return F.interpolate(x, size=tuple(size), mode='bilinear', align_corners=True)


def conv(P: SD, name: str, x: T, stride=1, padding=0) -> T:
    return F.conv2d(x, P[name + '.weight'], P[name + '.bias'], stride=stride, padding=padding)


def group_norm(P: SD, name: str, x: T) -> T:
    """GroupNorm(32 groups, eps=1e-6, affine).  archs/vqgan_arch.py:14-15."""
    return F.group_norm(x, 32, P[name + '.weight'], P[name + '.bias'], eps=1e-6)


def swish(x: T) -> T:
    """archs/vqgan_arch.py:18-20."""
    return x * torch.sigmoid(x)


def bn_eval(P: SD, name: str, x: T) -> T:
    """Eval-mode batch norm with running statistics.  sync_batchnorm/batchnorm.py:48-53."""
    return F.batch_norm(x, P[name + '.running_mean'], P[name + '.running_var'], P[name + '.weight'],
                        P[name + '.bias'], False, 0.1, 1e-5)


# ----------------------------------------------------------------------------------------------
# FOMM-style motion estimator: keypoint detector + dense motion
# ----------------------------------------------------------------------------------------------


def antialias_down(P: SD, name: str, x: T) -> T:
    """13x13 depthwise gaussian, zero pad 6, keep every 4th sample.
    utils/motion_estimator_util.py:599-645 (weight is the registered buffer `<name>.weight`)."""
    w = P[name + '.weight']
    k = w.shape[-1]
    pad = k // 2
    y = F.conv2d(F.pad(x, (pad, pad, pad, pad)), w, groups=w.shape[0])
    return y[:, :, ::4, ::4]


def hourglass(P: SD, name: str, x: T, num_blocks: int = 5) -> T:
    """Encoder (conv3x3+BN+ReLU+avgpool2) / decoder (nearest x2 + conv3x3+BN+ReLU, cat skip).
    utils/motion_estimator_util.py:214-231,363-380,440-492,551-563.  Returns the last decoder output."""
    skips = [x]
    for i in range(num_blocks):
        p = f'{name}.encoder.down_blocks.{i}'
        y = F.relu(bn_eval(P, p + '.norm', conv(P, p + '.conv', skips[-1], padding=1)))
        skips.append(F.avg_pool2d(y, 2))
    out = skips.pop()
    for i in range(num_blocks):
        p = f'{name}.decoder.up_blocks.{i}'
        out = F.interpolate(out, scale_factor=2)
        out = F.relu(bn_eval(P, p + '.norm', conv(P, p + '.conv', out, padding=1)))
        out = torch.cat([out, skips.pop()], dim=1)
    return out


def kp_detector(P: SD, x: T, temperature: float = 0.1, prefix: str = 'kp_detector') -> Dict[str, T]:
    """archs/keypoint_detector_arch.py:48-86."""
    feat = hourglass(P, prefix + '.predictor', antialias_down(P, prefix + '.down', x))
    pred = conv(P, prefix + '.kp', feat)                         # 7x7, pad 0 -> (B,15,58,58)
    b, k, h, w = pred.shape
    heat = F.softmax(pred.view(b, k, -1) / temperature, dim=2).view(b, k, h, w)
    grid = coord_grid(h, w, x.dtype, x.device)
    value = (heat.unsqueeze(-1) * grid.view(1, 1, h, w, 2)).sum(dim=(2, 3))
    jm = conv(P, prefix + '.jacobian', feat).reshape(b, k, 4, h, w)
    jac = (heat.unsqueeze(2) * jm).view(b, k, 4, -1).sum(-1).view(b, k, 2, 2)
    return {'value': value, 'jacobian': jac}


def gaussian_heatmaps(kp_value: T, h: int, w: int, var: float = 0.01) -> T:
    """utils/motion_estimator_util.py:11-32 -> (B,K,h,w)."""
    g = coord_grid(h, w, kp_value.dtype, kp_value.device).view(1, 1, h, w, 2)
    d = g - kp_value.view(*kp_value.shape[:2], 1, 1, 2)
    return torch.exp(-0.5 * (d ** 2).sum(-1) / var)


def dense_motion(P: SD, source: T, kp_driving: Dict[str, T], kp_source: Dict[str, T],
                 prefix: str = 'dense_motion_network') -> Dict[str, T]:
    """archs/dense_motion_arch.py:65-161 (single occlusion map, scale_factor 0.25)."""
    src = antialias_down(P, prefix + '.down', source)
    b, _, h, w = src.shape
    K = kp_driving['value'].shape[1]
    g_drv = gaussian_heatmaps(kp_driving['value'], h, w)
    g_src = gaussian_heatmaps(kp_source['value'], h, w)
    heat = torch.cat([torch.zeros(b, 1, h, w, dtype=src.dtype, device=src.device), g_drv - g_src], dim=1)      # (B,K+1,h,w)
    # sparse motions  T_{s<-d}(z) = J_s J_d^-1 (z - kp_d) + kp_s ; background = identity
    ident = coord_grid(h, w, src.dtype, src.device).view(1, 1, h, w, 2)
    z = ident - kp_driving['value'].view(b, K, 1, 1, 2)
    jac = torch.matmul(kp_source['jacobian'], torch.inverse(kp_driving['jacobian']))          # (B,K,2,2)
    z = torch.matmul(jac.view(b, K, 1, 1, 2, 2), z.unsqueeze(-1)).squeeze(-1)
    d2s = z + kp_source['value'].view(b, K, 1, 1, 2)
    sparse = torch.cat([ident.expand(b, 1, h, w, 2), d2s], dim=1)                            # (B,K+1,h,w,2)
    # deformed sources: grid_sample with the DEFAULT align_corners=False (dense_motion_arch.py:114)
    rep = src.unsqueeze(1).expand(b, K + 1, 3, h, w).reshape(b * (K + 1), 3, h, w)
    deformed = F.grid_sample(rep, sparse.reshape(b * (K + 1), h, w, 2), align_corners=False)
    deformed = deformed.view(b, K + 1, 3, h, w)
    inp = torch.cat([heat.unsqueeze(2), deformed], dim=2).view(b, (K + 1) * 4, h, w)
    feat = hourglass(P, prefix + '.hourglass', inp)
    mask = F.softmax(conv(P, prefix + '.mask', feat, padding=3), dim=1)                      # (B,K+1,h,w)
    deformation = (sparse.permute(0, 1, 4, 2, 3) * mask.unsqueeze(2)).sum(1).permute(0, 2, 3, 1)
    occlusion = torch.sigmoid(conv(P, prefix + '.occlusion', feat, padding=3))
    return {'deformation': deformation, 'occlusion_map': occlusion, 'driving_kp_heatmap': g_drv,
            'mask': mask, 'sparse_motion': sparse, 'sparse_deformed': deformed, 'source': src,
            'kp_heatmap': heat}


def hull_area(pts: np.ndarray) -> float:
    """Area of the 2-D convex hull (what scipy ConvexHull(...).volume returns; demo.py:26-28).
    Monotone chain + shoelace in float64."""
    p = sorted(map(tuple, np.asarray(pts, dtype=np.float64).tolist()))
    if len(p) < 3:
        return 0.0

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lo: List = []
    for q in p:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], q) <= 0:
            lo.pop()
        lo.append(q)
    up: List = []
    for q in reversed(p):
        while len(up) >= 2 and cross(up[-2], up[-1], q) <= 0:
            up.pop()
        up.append(q)
    hull = lo[:-1] + up[:-1]
    a = 0.0
    for i in range(len(hull)):
        x0, y0 = hull[i]
        x1, y1 = hull[(i + 1) % len(hull)]
        a += x0 * y1 - x1 * y0
    return abs(a) / 2.0


def normalize_kp(kp_source, kp_driving, kp_driving_initial, adapt_movement_scale=False,
                 use_relative_movement=False, use_relative_jacobian=False):
    """demo.py:24-44."""
    s = 1.0
    if adapt_movement_scale:
        s = math.sqrt(hull_area(kp_source['value'][0].cpu().numpy())) / \
            math.sqrt(hull_area(kp_driving_initial['value'][0].cpu().numpy()))
    out = dict(kp_driving)
    if use_relative_movement:
        out['value'] = (kp_driving['value'] - kp_driving_initial['value']) * s + kp_source['value']
        if use_relative_jacobian:
            jd = torch.matmul(kp_driving['jacobian'], torch.inverse(kp_driving_initial['jacobian']))
            out['jacobian'] = torch.matmul(jd, kp_source['jacobian'])
    return out


# ----------------------------------------------------------------------------------------------
# VQGAN blocks (archs/vqgan_arch.py)
# ----------------------------------------------------------------------------------------------


def res_block(P: SD, name: str, x: T) -> T:
    """archs/vqgan_arch.py:168-191."""
    h = conv(P, name + '.conv1', swish(group_norm(P, name + '.norm1', x)), padding=1)
    h = conv(P, name + '.conv2', swish(group_norm(P, name + '.norm2', h)), padding=1)
    if (name + '.conv_out.weight') in P:
        x = conv(P, name + '.conv_out', x)
    return x + h


def attn_block(P: SD, name: str, x: T) -> T:
    """Single-head spatial self-attention.  archs/vqgan_arch.py:194-253."""
    h = group_norm(P, name + '.norm', x)
    b, c, hh, ww = x.shape
    q = conv(P, name + '.q', h).reshape(b, c, -1).permute(0, 2, 1)
    k = conv(P, name + '.k', h).reshape(b, c, -1)
    v = conv(P, name + '.v', h).reshape(b, c, -1)
    a = F.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    o = torch.bmm(v, a.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + conv(P, name + '.proj_out', o)


def downsample(P: SD, name: str, x: T) -> T:
    """pad right/bottom 1, conv3x3 stride 2 pad 0.  archs/vqgan_arch.py:144-153."""
    return conv(P, name + '.conv', F.pad(x, (0, 1, 0, 1)), stride=2)


def upsample(P: SD, name: str, x: T) -> T:
    """nearest x2 then conv3x3.  archs/vqgan_arch.py:156-165."""
    return conv(P, name + '.conv', F.interpolate(x, scale_factor=2.0, mode='nearest'), padding=1)


# block lists for nf=64, ch_mult=[1,2,2,4], res_blocks=2, attn_resolutions=[32]
# (archs/vqgan_arch.py:256-350 instantiated by appmotioncodebook_arch.py:172,190)
ENCODER_BLOCKS = ['conv', 'res', 'res', 'down', 'res', 'res', 'down', 'res', 'res', 'down',
                  'res', 'attn', 'res', 'attn', 'res', 'attn', 'res', 'norm', 'conv']
GENERATOR_BLOCKS = ['conv', 'res', 'attn', 'res', 'res', 'attn', 'res', 'attn', 'up',
                    'res', 'res', 'up', 'res', 'res', 'up', 'res', 'res', 'norm', 'conv']


def run_block(P: SD, kind: str, name: str, x: T) -> T:
    if kind == 'conv':
        return conv(P, name, x, padding=1)
    if kind == 'res':
        return res_block(P, name, x)
    if kind == 'attn':
        return attn_block(P, name, x)
    if kind == 'down':
        return downsample(P, name, x)
    if kind == 'up':
        return upsample(P, name, x)
    if kind == 'norm':
        return group_norm(P, name, x)
    raise ValueError(kind)


def encode_source(P: SD, x: T) -> Dict[str, T]:
    """Encoder loop with taps after blocks 2/5/8 (archs/appmotioncodebook_arch.py:327,549-554).
    Depends on the source only."""
    feats = {}
    for i, kind in enumerate(ENCODER_BLOCKS):
        x = run_block(P, kind, f'encoder.blocks.{i}', x)
        if i in (2, 5, 8):
            feats[str(x.shape[-1])] = x
    feats[str(x.shape[-1])] = x          # the latent: '32' at 256x256, '64' for the 512x512 variant
    return feats


def encode_driving(P: SD, x: T) -> Dict[str, T]:
    """AppMotionCompFormer.encode_driving (archs/appmotioncodebook_arch.py:364-371; the same taps feed app_codebook_loss, :433-439): features after
    encoder blocks 2 / 5 / 8 and **11** - the '32' entry is the output of the first attention block at 32x32 (`fuse_encoder_block['32'] = 11`, :327),
    NOT the latent the forward warps (that is the output of the last block)."""
    feats = {}
    for i, kind in enumerate(ENCODER_BLOCKS):
        x = run_block(P, kind, f'encoder.blocks.{i}', x)
        if i in (2, 5, 8, 11):
            feats[str(x.shape[-1])] = x
    return feats


# ----------------------------------------------------------------------------------------------
# codebook transformer layer + vector quantizer
# ----------------------------------------------------------------------------------------------


def mha(P: SD, name: str, q: T, k: T, v: T, n_head: int, key_padding_mask: Optional[T] = None) -> T:
    """nn.MultiheadAttention forward restated (packed in_proj, heads = contiguous E/n slices,
    scale 1/sqrt(E/n), -inf on masked keys, fp32 softmax, out_proj).  q:(L,B,E) k,v:(S,B,E).
    archs/appmotioncodebook_arch.py:69-70,101,115."""
    L, B, E = q.shape
    S = k.shape[0]
    W, bias = P[name + '.in_proj_weight'], P[name + '.in_proj_bias']
    d = E // n_head
    qp = F.linear(q, W[:E], bias[:E]).view(L, B, n_head, d).permute(1, 2, 0, 3)
    kp = F.linear(k, W[E:2 * E], bias[E:2 * E]).view(S, B, n_head, d).permute(1, 2, 0, 3)
    vp = F.linear(v, W[2 * E:], bias[2 * E:]).view(S, B, n_head, d).permute(1, 2, 0, 3)
    s = torch.matmul(qp * (d ** -0.5), kp.transpose(-1, -2))                       # (B,H,L,S)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask.view(B, 1, 1, S), float('-inf'))
    o = torch.matmul(F.softmax(s, dim=-1), vp)                                     # (B,H,L,d)
    o = o.permute(2, 0, 1, 3).reshape(L, B, E)
    return F.linear(o, P[name + '.out_proj.weight'], P[name + '.out_proj.bias'])


def transformer_layer(P: SD, name: str, t: T, ctx: T, pos: T, n_head: int = 8,
                      key_padding_mask: Optional[T] = None) -> T:
    """archs/appmotioncodebook_arch.py:88-126.  t,pos:(L,B,E) with L = 1024 tokens on a 32x32 grid (4096 on 64x64 for the 512x512
    variant)  ctx:(K,B,E)."""
    L, B, E = t.shape
    tg = int(round(math.sqrt(L)))
    u = F.layer_norm(t, (E,), P[name + '.norm1.weight'], P[name + '.norm1.bias'])
    t = t + mha(P, name + '.self_attn', u + pos, u + pos, u, n_head, key_padding_mask)
    u = F.layer_norm(t, (E,), P[name + '.norm2.weight'], P[name + '.norm2.bias'])
    t = t + mha(P, name + '.cross_attn', u + pos, ctx, ctx, n_head)
    u = F.layer_norm(t, (E,), P[name + '.norm3.weight'], P[name + '.norm3.bias'])
    u = u.permute(1, 2, 0).reshape(B, E, tg, tg)
    u = conv(P, name + '.conv2', F.gelu(conv(P, name + '.conv1', u, padding=1)), padding=1)
    return t + u.reshape(B, E, L).permute(2, 0, 1)


def vq_lookup(codebook: T, z: T, scale: Optional[float] = None):
    """VectorQuantizer.forward restated (archs/vqgan_arch.py:33-93) for the shared-prefix split.
    z:(B,E,H,W) -> z_q (B,E,H,W), loss, indices (B*H*W,1) int64, mean_distance, perplexity."""
    n = codebook.shape[0] if scale is None else int(scale * codebook.shape[0])
    e = codebook[:n]
    zl = z.permute(0, 2, 3, 1).contiguous()
    zf = zl.view(-1, codebook.shape[1])
    d = (zf ** 2).sum(dim=1, keepdim=True) + (e ** 2).sum(1) - 2 * torch.matmul(zf, e.t())
    idx = torch.argmin(d, dim=1)
    zq = e[idx].view(zl.shape)
    loss = 0.25 * torch.mean((zq - zl) ** 2) + torch.mean((zq - zl) ** 2)
    onehot_mean = torch.bincount(idx, minlength=n).to(z.dtype) / idx.numel()
    perplexity = torch.exp(-torch.sum(onehot_mean * torch.log(onehot_mean + 1e-10)))
    return zq.permute(0, 3, 1, 2).contiguous(), loss, idx.unsqueeze(1), d.mean(), perplexity


def vq_forward(codebook: T, z: T, scale: Optional[float] = None, beta: float = 0.25):
    """What VectorQuantizer.forward RETURNS as its first two values (archs/vqgan_arch.py:76-80,88): the straight-through tensor
    z + (z_q - z) (its forward value differs from z_q by one rounding) and the loss beta * mean((z_q - z)^2) + mean((z_q - z)^2)."""
    zq, _, idx, _, _ = vq_lookup(codebook, z, scale)
    loss = beta * torch.mean((zq - z) ** 2) + torch.mean((zq - z) ** 2)
    return z + (zq - z), loss, idx


def to_motion(P: SD, x: T) -> T:
    """self.to_motion = Upsample, ResBlock, GroupNorm, conv3x3 32 -> 2 (archs/appmotioncodebook_arch.py:290-292)."""
    x = upsample(P, 'to_motion.0', x)
    x = res_block(P, 'to_motion.1', x)
    return conv(P, 'to_motion.3', group_norm(P, 'to_motion.2', x), padding=1)


def app_codebook_loss(P: SD, gt: T, beta: float = 0.25):
    """AppMotionCompFormer.app_codebook_loss (archs/appmotioncodebook_arch.py:429-469), split = 1, shared codebook prefixes: the driving ("gt")
    frame's encoder features are embedded to the 32x32 token grid, quantised against the first 256 k rows of the appearance codebook and mapped
    back.  -> ([[app_recon, app_feat_original, quant_app, app_feat, feat_com] per scale 32, 64, 128, 256], [loss per scale])."""
    feats = encode_driving(P, gt)
    recon, losses = [], []
    cb = P['quantize_app.embedding.weight']
    for w in (32, 64, 128, 256):
        f = feats[str(w)]
        b, c = f.shape[:2]
        if w == 32:
            app_feat = conv(P, 'app_feat_emb_32', f)
        else:
            p = w // 32
            x = f.view(b, c, 32, p, 32, p).permute(0, 2, 4, 3, 5, 1).reshape(b, 1024, p * p * c)          # Rearrange 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'
            app_feat = F.linear(x, P[f'app_feat_emb_{w}.1.weight'], P[f'app_feat_emb_{w}.1.bias'])        # (b, 1024, 256)
            app_feat = app_feat.permute(0, 2, 1).reshape(b, app_feat.shape[2], 32, 32)                    # Rearrange 'b n d -> b d n', then :452-453
        quant, loss, _ = vq_forward(cb, app_feat, SCALE_K[w] / 4.0, beta)

        def back(t):
            if w == 32:
                return conv(P, 'to_app_feat_32', t)
            y = F.linear(t.reshape(b, t.shape[1], 1024).permute(0, 2, 1), P[f'to_app_feat_{w}.0.weight'], P[f'to_app_feat_{w}.0.bias'])
            p = w // 32
            return y.view(b, 32, 32, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, w, w)
        recon.append([back(quant), back(app_feat), quant, app_feat, f])
        losses.append(loss)
    return recon, losses


# ----------------------------------------------------------------------------------------------
# AppMotionCompFormer forward (inference=True, w given)   archs/appmotioncodebook_arch.py:546-764
# ----------------------------------------------------------------------------------------------

SCALE_K = {32: 1, 64: 2, 128: 3, 256: 4}

# The 512x512 variant (BASELINE configs[3]; SURVEY.md 8d "Config 4").  The reference itself cannot run it: its 32x32 token grid,
# position_emb (1024,E) and resolution-keyed module names are hard-coded.  The variant defined there is the direct generalisation: the same
# layer graph and weight shapes except position_emb_* (4096,E); every spatial size doubles (token grid 64x64, flow / occlusion grid 128x128,
# feature scales 64..512) while module names and codebook prefixes keep their 256x256 ("nominal") scale s0 = s / R, R = image size / 256.
# Everything below infers R from the tensors it is given (flow grid = 64 R), so the 256x256 path is untouched.  PARITY UNPINNED BY THE
# REFERENCE at 512: this restatement is proven equal to the reference at 256 (oracle/make_golden.py) and is then run at 512.


def warp_ac(feat: T, deformation: T) -> T:
    """deform_input: flow resized (bilinear, align_corners=True) to the feature size, then
    grid_sample bilinear / zeros / align_corners=True.  appmotioncodebook_arch.py:349-356."""
    h, w = feat.shape[-2:]
    d = resize_ac(deformation.permute(0, 3, 1, 2), (h, w)).permute(0, 2, 3, 1)
    if feat.shape[0] != d.shape[0]:
        feat = feat.expand(d.shape[0], -1, -1, -1)
    return F.grid_sample(feat, d, mode='bilinear', padding_mode='zeros', align_corners=True)


def occlude(feat: T, occ: T) -> T:
    """appmotioncodebook_arch.py:358-362."""
    return feat * resize_ac(occ, feat.shape[-2:])


def motion_compensation(P: SD, flow_px: T, qfeat: T, warp0: T, s: int, collect: Optional[dict] = None) -> T:
    """motion_codebook_compensation, inference branch (appmotioncodebook_arch.py:373-427) with
    BasicMotionEncoder (:129-147) and RefineFlow (:150-168).  flow_px:(B,64,64,2) in pixels.
    Returns (B,64,64,3): delta-flow (px) and delta-occlusion logit."""
    b, h, w, _ = flow_px.shape
    m = flow_px.permute(0, 3, 1, 2)
    mf = conv(P, 'motion_emb.0', m, padding=1)
    mf = downsample(P, 'motion_emb.1', mf)
    mf = res_block(P, 'motion_emb.2', mf)                                          # (B,32,32,32)
    if collect is not None:
        collect[f'm_feat_{s}'] = mf
    q = conv(P, 'motion_query_enc_2', torch.cat([mf, resize_ac(qfeat, mf.shape[-2:])], dim=1))
    E = q.shape[1]
    R = h // 64                                                                    # 1 at 256x256, 2 for the 512x512 variant
    s0, tg = s // R, mf.shape[-1]                                                  # nominal scale (names, codebook prefix), token grid
    t = q.reshape(b, E, tg * tg).permute(2, 0, 1)
    pos = P['position_emb_motion'].unsqueeze(1).expand(-1, b, -1)
    ctx = P['quantize_motion.embedding.weight'][:256 * SCALE_K[s0]].unsqueeze(1).expand(-1, b, -1)
    for i in range(2):
        t = transformer_layer(P, f'motion_block.{i}', t, ctx, pos)
    mfeat = resize_ac(t.permute(1, 2, 0).reshape(b, E, tg, tg), (h, w))
    # BasicMotionEncoder
    cor = F.relu(conv(P, 'BasicMotionEncoder.convc1', mfeat))
    cor = F.relu(conv(P, 'BasicMotionEncoder.convc2', cor, padding=1))
    flo = F.relu(conv(P, 'BasicMotionEncoder.convf1', m, padding=3))
    flo = F.relu(conv(P, 'BasicMotionEncoder.convf2', flo, padding=1))
    mo = F.relu(conv(P, 'BasicMotionEncoder.conv', torch.cat([cor, flo], dim=1), padding=1))
    m_f = torch.cat([mo, m], dim=1)                                                # 128 ch
    ctxf = F.relu(conv(P, f'to_context.{int(math.log2(s0)) - 5}', warp0))
    ctxf = resize_ac(ctxf, (h, w))
    # RefineFlow
    c = F.relu(conv(P, 'refine.convc1', ctxf, padding=1))
    z = torch.cat([m_f, c], dim=1)
    dflow = conv(P, 'refine.conv2', F.relu(conv(P, 'refine.conv1', z, padding=1)), padding=1)
    docc = conv(P, 'refine.convo2', F.relu(conv(P, 'refine.convo1', z, padding=1)), padding=1)
    return torch.cat([dflow, docc], dim=1).permute(0, 2, 3, 1)


def app_compensation(P: SD, feat: T, m_com: T) -> T:
    """app_codebook_compensation (appmotioncodebook_arch.py:472-544), split=1, shared prefixes."""
    b, c, s, _ = feat.shape
    R = m_com.shape[1] // 64
    s0, tg = s // R, 32 * R                                                        # nominal scale, token grid (32x32; 64x64 at 512x512)
    L = tg * tg
    m32 = resize_ac(m_com.permute(0, 3, 1, 2), (tg, tg)).reshape(b, 2, L)
    ignore = ((m32 > 1) | (m32 < -1)).any(dim=1)                                   # (B,L) bool
    if s0 == 32:
        tok = conv(P, 'app_feat_emb_32', feat).reshape(b, 256, L).permute(2, 0, 1)
    else:
        p = s0 // 32
        x = feat.view(b, c, tg, p, tg, p).permute(0, 2, 4, 3, 5, 1).reshape(b, L, p * p * c)
        tok = F.linear(x, P[f'app_feat_emb_{s0}.1.weight'], P[f'app_feat_emb_{s0}.1.bias']).permute(1, 0, 2)
    pos = P['position_emb_app'].unsqueeze(1).expand(-1, b, -1)
    ctx = P['quantize_app.embedding.weight'][:256 * SCALE_K[s0]].unsqueeze(1).expand(-1, b, -1)
    tok = transformer_layer(P, 'app_block.0', tok, ctx, pos, key_padding_mask=ignore)
    tok = transformer_layer(P, 'app_block.1', tok, ctx, pos)
    if s0 == 32:
        return conv(P, 'to_app_feat_32', tok.permute(1, 2, 0).reshape(b, 256, tg, tg))
    y = F.linear(tok.permute(1, 0, 2), P[f'to_app_feat_{s0}.0.weight'], P[f'to_app_feat_{s0}.0.bias'])
    p = s0 // 32
    return y.view(b, tg, tg, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, s, s)


def sft_fuse(P: SD, name: str, enc: T, dec: T, w: float) -> T:
    """Fuse_sft_block.forward (appmotioncodebook_arch.py:43-52)."""
    e = res_block(P, name + '.encode_enc', torch.cat([enc, dec], dim=1))
    scale = conv(P, name + '.scale.2', F.leaky_relu(conv(P, name + '.scale.0', e, padding=1), 0.2), padding=1)
    shift = conv(P, name + '.shift.2', F.leaky_relu(conv(P, name + '.shift.0', e, padding=1), 0.2), padding=1)
    return dec + w * (dec * scale + shift)


def decode_plain(P: SD, x: T) -> T:
    """Generator.forward (vqgan_arch.py:347-350): the decoder blocks alone, no fusion - what
    `net_g.generator(lq_feat)` computes in AppMotionCompModel.test (models/appmotioncomp_model.py:453-454)."""
    for i, kind in enumerate(GENERATOR_BLOCKS):
        x = run_block(P, kind, f'generator.blocks.{i}', x)
    return x


def generator_forward(P: SD, src_feats: Dict[str, T], dm: Dict[str, T], w: float = 1.0,
                      collect: Optional[dict] = None) -> Dict[str, T]:
    """Everything in AppMotionCompFormer.forward after the (source-only) encoder loop,
    inference=True (appmotioncodebook_arch.py:556-764).  src_feats from encode_source (batch 1 or B)."""
    deformation = dm['deformation']
    b, hs, ws, _ = deformation.shape
    # the residual grid uses linspace (appmotioncodebook_arch.py:562-565), not the arithmetic grid
    xx = torch.linspace(-1., 1., hs, device=deformation.device)
    yy = torch.linspace(-1., 1., ws, device=deformation.device)
    gx, gy = torch.meshgrid(xx, yy, indexing='xy')
    grid = torch.stack([gx, gy], dim=-1).unsqueeze(0)
    half = (hs - 1.) / 2.
    motions = [deformation]
    occs: List[T] = []
    R = hs // 64
    tg = 32 * R
    kp_feat = F.relu(conv(P, 'driving_kp_enc', resize_ac(dm['driving_kp_heatmap'], (tg, tg))))

    def compensate(feat_s: T, s: int, occ_prev: T):
        m_prev = motions[-1]
        warp0 = warp_ac(feat_s, m_prev)
        ws_ = F.relu(conv(P, f'warped_source_enc_{s // R}', resize_ac(warp0, (tg, tg))))
        qf = conv(P, 'motion_query_enc_1', torch.cat([ws_, kp_feat], dim=1))
        res = motion_compensation(P, (m_prev - grid) * half, qf, warp0, s, collect)
        m_com = m_prev + res[..., 0:2] / half
        motions.append(m_com)
        occ = torch.sigmoid(occ_prev + res[..., 2:].permute(0, 3, 1, 2))
        occs.append(occ)
        warped = occlude(warp_ac(feat_s, m_com), occ)
        out = app_compensation(P, warped, m_com)
        if collect is not None:
            collect[f'warp0_{s}'] = warp0
            collect[f'res_{s}'] = res
            collect[f'warped_{s}'] = warped
            collect[f'app_{s}'] = out
        return out

    x = compensate(src_feats[str(tg)], tg, dm['occlusion_map'])
    lq_feat = x
    for i, kind in enumerate(GENERATOR_BLOCKS):
        x = run_block(P, kind, f'generator.blocks.{i}', x)
        if i in (9, 12, 15) and w > 0:
            s = x.shape[-1]
            enc = compensate(src_feats[str(s)], s, occs[-1])
            x = sft_fuse(P, f'fuse_convs_dict.{s // R}', enc, x, w)
            if collect is not None:
                collect[f'sft_{s}'] = x
            x = x + conv(P, f'fuse_ms_dict.{s // R}', enc, padding=1)
            if collect is not None:
                collect[f'fused_{s}'] = x
    return {'out': x, 'lq_feat': lq_feat, 'out_occ': occs, 'deformation_list': motions}


def generator_forward_train(P: SD, src_feats: Dict[str, T], dm: Dict[str, T], w: float = 1.0, gt: Optional[T] = None, beta: float = 0.25) -> Dict[str, T]:
    """The VALUES AppMotionCompFormer.forward adds with inference=False (archs/appmotioncodebook_arch.py:379-386,424-427,580-587,641-662,
    676-683,749-757): per scale the motion feature is quantised against the first 256 k rows of the motion codebook and decoded by `to_motion`
    (`motion_recon_list`, in grid units; `codebook_loss_motion_list`), the un-fused decoder runs on lq_feat (`out_lr`), and with a ground-truth frame
    `app_codebook_loss` (`app_recon_list`, `codebook_loss_app_list`).  Forward values only: gradients are outside the path."""
    collect: dict = {}
    out = generator_forward(P, src_feats, dm, w, collect)
    half = (dm['deformation'].shape[1] - 1.) / 2.
    R = dm['deformation'].shape[1] // 64
    cb = P['quantize_motion.embedding.weight']
    recon, losses = [], []
    for s0 in (32, 64, 128, 256):
        quant, loss, _ = vq_forward(cb, collect[f'm_feat_{s0 * R}'], SCALE_K[s0] / 4.0, beta)
        recon.append(to_motion(P, quant).permute(0, 2, 3, 1) / half)
        losses.append(loss)
    out.update(out_lr=[decode_plain(P, out['lq_feat'])], motion_recon_list=recon, codebook_loss_motion_list=losses)
    if gt is not None:
        out['app_recon_list'], out['codebook_loss_app_list'] = app_codebook_loss(P, gt, beta)
    return out


def to_uint8(x: T, bgr: bool = False) -> np.ndarray:
    """tensor2img for one (3,H,W) frame in [-1,1] (utils/img_util.py:42-98): clamp, ->[0,1],
    HWC, optional channel flip, x255, round half to even, uint8."""
    t = x.detach().float().clamp(-1, 1)
    t = (t - (-1)) / (1 - (-1))
    a = t.cpu().numpy().transpose(1, 2, 0)
    if bgr:
        a = a[:, :, ::-1]
    return (a * 255.0).round().astype(np.uint8)


def make_animation(P_g: SD, P_me: SD, source: T, driving: List[T], relative=True,
                   adapt_movement_scale=True, w: float = 1.0, bgr: bool = False, batch: int = 1):
    """demo.make_animation restated (demo.py:103-134): source (3,H,W), driving list of (3,H,W).
    Returns (list of HWC uint8 predictions, list of HWC uint8 driving frames, last fp32 out)."""
    with torch.no_grad():
        src = source.unsqueeze(0)
        kp_s = kp_detector(P_me, src)
        kp_0 = kp_detector(P_me, driving[0].unsqueeze(0))
        feats = encode_source(P_g, src)
        preds, drvs, outs = [], [], []
        for i0 in range(0, len(driving), batch):
            frames = torch.stack(driving[i0:i0 + batch])
            kp_d = kp_detector(P_me, frames)
            kp_n = normalize_kp(kp_s, kp_d, kp_0, adapt_movement_scale, relative, relative)
            nb = frames.shape[0]
            kp_sb = {k: v.expand(nb, *v.shape[1:]) for k, v in kp_s.items()}
            dm = dense_motion(P_me, src.expand(nb, -1, -1, -1), kp_n, kp_sb)
            out = generator_forward(P_g, feats, dm, w)['out']
            for j in range(nb):
                preds.append(to_uint8(out[j], bgr))
                drvs.append(to_uint8(frames[j], bgr))
                outs.append(out[j])
    return preds, drvs, outs


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (BASELINE.md "CPU-baseline plan" item 2, made
# reference-independent so they can be regenerated on the GPU box)
# ----------------------------------------------------------------------------------------------


def synthetic_state_dict(shapes: Dict[str, List[int]], seed: int = 0) -> SD:
    """Fill a {key: shape} inventory (tests/golden/state_keys.json) deterministically.
    Conv/linear weights ~ U(-a,a), a = sqrt(1/fan_in) (the bound of torch.nn default init); biases ~ N(0,0.05);
    norm gamma ~ 1+N(0,0.1), beta ~ N(0,0.1); BN running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5);
    position embeddings ~ N(0,0.02); codebooks ~ N(0,1); kp jacobian conv weight ~ N(0,1e-3) with
    bias [1,0,0,1]; antialias kernels are the fixed gaussian (computed, not random)."""
    g = torch.Generator().manual_seed(seed)
    out: SD = {}
    for k in sorted(shapes):
        shp = list(shapes[k])
        leaf = k.rsplit('.', 1)[-1]
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith('down.weight') and len(shp) == 4 and shp[1] == 1 and shp[2] == 13:
            ax = torch.arange(13, dtype=torch.float32)
            g1 = torch.exp(-(ax - 6.0) ** 2 / (2 * 1.5 ** 2))
            k2 = g1.view(13, 1) * g1.view(1, 13)
            out[k] = (k2 / k2.sum()).view(1, 1, 13, 13).repeat(shp[0], 1, 1, 1)
        elif k.startswith('position_emb'):
            out[k] = torch.randn(shp, generator=g) * 0.02
        elif k.endswith('embedding.weight'):
            out[k] = torch.randn(shp, generator=g)
        elif leaf == 'running_mean':
            out[k] = torch.randn(shp, generator=g) * 0.1
        elif leaf == 'running_var':
            out[k] = torch.rand(shp, generator=g) + 0.5
        elif 'norm' in k.rsplit('.', 2)[-2] and leaf == 'weight':
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif 'norm' in k.rsplit('.', 2)[-2] and leaf == 'bias':
            out[k] = 0.1 * torch.randn(shp, generator=g)
        elif k == 'to_motion.2.weight':
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k == 'to_motion.2.bias':
            out[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith('kp_detector.jacobian.weight'):
            out[k] = torch.randn(shp, generator=g) * 1e-3
        elif k.endswith('kp_detector.jacobian.bias'):
            out[k] = torch.tensor([1., 0., 0., 1.] * (shp[0] // 4))
        elif leaf in ('weight', 'in_proj_weight'):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            a = math.sqrt(1.0 / fan_in)   # = torch default kaiming_uniform(a=sqrt(5)) bound
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif leaf in ('bias', 'in_proj_bias'):
            out[k] = torch.randn(shp, generator=g) * 0.05
        else:
            raise KeyError(f'no synthetic rule for {k} {shp}')
    return out


def synthetic_frames(n_driving: int, seed: int = 1234, size: int = 256, smooth: bool = True):
    """Source + driving frames in [-1,1].  `smooth`: low-frequency random fields (so the keypoint
    soft-argmax and flows are well conditioned, like a face crop), else iid U(-1,1)."""
    g = torch.Generator().manual_seed(seed)

    def frame():
        if not smooth:
            return torch.rand(3, size, size, generator=g) * 2 - 1
        lo = torch.randn(1, 3, 8, 8, generator=g)
        hi = torch.randn(1, 3, 32, 32, generator=g) * 0.35
        f = F.interpolate(lo, size=(size, size), mode='bicubic', align_corners=False) + \
            F.interpolate(hi, size=(size, size), mode='bicubic', align_corners=False)
        return torch.tanh(f[0] * 0.8)
    src = frame()
    drv = [frame() for _ in range(n_driving)]
    return src, drv


def variant_shapes(shapes: Dict[str, List[int]], img_size: int = 256) -> Dict[str, List[int]]:
    """State-dict inventory of the `img_size` variant: identical to the reference's (tests/golden/state_keys.json) except the position
    embeddings, which have one row per token of the (img_size/8)^2 grid (Config 4)."""
    tg = img_size // 8
    out = {k: list(v) for k, v in shapes.items()}
    for k in ('position_emb_app', 'position_emb_motion'):
        if k in out:
            out[k] = [tg * tg, out[k][1]]
    return out
