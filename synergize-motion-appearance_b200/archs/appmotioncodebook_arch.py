"""AppMotionCompFormer on B200 kernels.

Drop-in for basicsr/archs/appmotioncodebook_arch.py:170-764 (generator of options/test.yml:8-45, built
on the VQGAN blocks of basicsr/archs/vqgan_arch.py): same constructor kwargs, same 472 state_dict keys,
same call `net_g(source, dense_motion, w=1, inference=True) -> dict` with the reference's full inference key set
(`out`, `lq_feat`, `out_occ`, `deformation_list`, `res_deform_list`, `deform_feat_list`, `app_comp_list`,
`app_before_comp_list`; `app_query_feat_list` / `app_comp_feat_list` with visualize_app_feat, `x_before_app_32` with
vis_app_before_comp: appmotioncodebook_arch.py:745-764), plus `encode_driving`, `app_codebook_loss` and the callable `generator`.
`inference=False` returns the forward VALUES of the training branch as well (`out_lr`, `motion_recon_list`, `codebook_loss_motion_list`, and with
`gt` `app_recon_list` / `codebook_loss_app_list`: the VQ lookup runs inside the forward); there is no autograd graph - backward is outside the path.

Design (B200-first, not a module-for-module port):
  * everything NHWC fp32; torch only allocates buffers; every op is a kernel from csrc/;
  * per-source work is hoisted and cached: the 19 encoder blocks run once per source, the codebook K/V
    projections of all cross-attentions once per weight load (the reference redoes both every frame,
    appmotioncodebook_arch.py:549-554,405,514);
  * B driving frames share one source: source features are passed with batch stride 0;
  * GroupNorm = statistics kernel + normalise/swish fused into the consuming conv's operand load;
    residual adds, activations, biases, nearest upsampling, (un)patchify fused into the conv kernel;
    torch.cat is replaced by writing producers into channel slices of one buffer;
  * the third (duplicate) warp per scale that only feeds `deform_feat_list` is not computed.
"""
import math
import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..registry import ARCH_REGISTRY
from .params import ParamModule

# block kinds for nf=64, ch_mult=[1,2,2,4], res_blocks=2, attn_resolutions=[32] (vqgan_arch.py:256-350)


def _encoder_layout(nf, ch_mult, res_blocks, resolution, attn_res) -> List[Tuple[str, int, int]]:
    blocks = [('conv', 3, nf)]
    cur = resolution
    in_mult = (1,) + tuple(ch_mult)
    cin = nf
    for i in range(len(ch_mult)):
        cin, cout = nf * in_mult[i], nf * ch_mult[i]
        for _ in range(res_blocks):
            blocks.append(('res', cin, cout))
            cin = cout
            if cur in attn_res:
                blocks.append(('attn', cin, cin))
        if i != len(ch_mult) - 1:
            blocks.append(('down', cin, cin))
            cur //= 2
    blocks += [('res', cin, cin), ('attn', cin, cin), ('res', cin, cin), ('norm', cin, cin)]
    return blocks, cin


def _generator_layout(nf, ch_mult, res_blocks, resolution, attn_res, emb_dim):
    cin = nf * ch_mult[-1]
    cur = resolution // 2 ** (len(ch_mult) - 1)
    blocks = [('conv', emb_dim, cin), ('res', cin, cin), ('attn', cin, cin), ('res', cin, cin)]
    for i in reversed(range(len(ch_mult))):
        cout = nf * ch_mult[i]
        for _ in range(res_blocks):
            blocks.append(('res', cin, cout))
            cin = cout
            if cur in attn_res:
                blocks.append(('attn', cin, cin))
        if i != 0:
            blocks.append(('up', cin, cin))
            cur *= 2
    blocks += [('norm', cin, cin), ('conv', cin, 3)]
    return blocks


class _PlainDecoder(nn.Module):
    """`net_g.generator(feat)`: the reference's Generator.forward (vqgan_arch.py:347-350), which AppMotionCompModel.test calls on `lq_feat`
    (models/appmotioncomp_model.py:453-454).  The parameter container named `generator` is given this class so that the attribute is callable."""

    def forward(self, x):
        return self._owner().decode_plain(x)


@ARCH_REGISTRY.register()
class AppMotionCompFormer(ParamModule):
    SCALES = (32, 64, 128, 256)

    def __init__(self, img_size=256, nf=64, ch_mult=[1, 2, 2, 4], res_blocks=2, attn_resolutions=[32],
                 quantizer_type='nearest', beta=0.25, codebook_size_motion=1024, embed_dim_motion=32,
                 codebook_size_app=1024, embed_dim_app=256, n_head=8, dim_embd_motion=32, n_layers_motion=2,
                 dim_embd_app=256, n_layers_app=2, split=1, num_kp=15, with_position_emb=True, warp_s_d_kp_query=True,
                 MRFA_motion_enc=True, motion_codebook_split=True, detach_motion_query=True,
                 multiscale_feature_fusion=True, multiscale_sft=True, app_codebook_split=True,
                 wo_motion_cdbk_share=False, wo_app_cdbk_share=False, connect_list=['64', '128', '256'],
                 connect_app_list=['32', '64', '128', '256'], fix_modules=[], ae_path=None):
        super().__init__()
        supported = (img_size in (256, 512) and nf == 64 and list(ch_mult) == [1, 2, 2, 4] and res_blocks == 2 and
                     list(attn_resolutions) == [32] and quantizer_type == 'nearest' and split == 1 and with_position_emb and
                     warp_s_d_kp_query and MRFA_motion_enc and motion_codebook_split and multiscale_feature_fusion and
                     multiscale_sft and app_codebook_split and not wo_motion_cdbk_share and not wo_app_cdbk_share and
                     list(connect_list) == ['64', '128', '256'] and list(connect_app_list) == ['32', '64', '128', '256'] and
                     embed_dim_motion == dim_embd_motion and embed_dim_app == dim_embd_app and n_layers_motion == 2 and
                     n_layers_app == 2 and dim_embd_app == 256 and dim_embd_motion == 32 and n_head == 8 and
                     codebook_size_app % 4 == 0 and codebook_size_motion % 4 == 0 and
                     codebook_size_app // 4 % 64 == 0 and codebook_size_motion // 4 % 8 == 0)
        if not supported:
            raise NotImplementedError('B200 AppMotionCompFormer implements the options/test.yml configuration (and its 512x512 variant) only')
        self.beta, self.n_head, self.num_kp = beta, n_head, num_kp
        self.Ea, self.Em = dim_embd_app, dim_embd_motion
        self.n_codes_app, self.n_codes_motion = codebook_size_app, codebook_size_motion
        # 512x512 variant (BASELINE configs[3]; SURVEY.md 8d "Config 4"; the reference itself crashes at 512): the same layer graph and weight
        # shapes except position_emb_* (one row per token), every spatial size scaled by R = img_size / 256 (token grid 32R, flow / occlusion
        # grid 64R, feature scales 32R..256R); module names and codebook prefixes keep their NOMINAL 256x256 scale s0 = s / R.
        self.img_size, self.R = img_size, img_size // 256
        self.tg, self.fg = 32 * self.R, 64 * self.R               # token grid, flow / occlusion grid
        self.L = self.tg * self.tg
        self.channels = {32: 256, 64: 128, 128: 128, 256: 64}      # keyed by nominal scale
        self.connect_list, self.connect_app_list = list(connect_list), list(connect_app_list)
        self.fuse_encoder_block = {'256': 2, '128': 5, '64': 8, '32': 11}
        self.fuse_generator_block = {'32': 6, '64': 9, '128': 12, '256': 15}
        self.enc_layout, latent_c = _encoder_layout(nf, ch_mult, res_blocks, 256, attn_resolutions)      # (attention blocks sit at the latent scale)
        self.enc_layout.append(('conv', latent_c, 256))
        self.gen_layout = _generator_layout(nf, ch_mult, res_blocks, 256, attn_resolutions, 256)
        self._declare_all()
        gen = self._modules['generator']
        gen.__class__ = _PlainDecoder
        object.__setattr__(gen, '_owner', weakref.ref(self))      # not a registered sub-module: no cycle in the module tree
        self._src_cache = None
        self._two_tensor_ok = os.environ.get('SMA_NO_TWO', '0') != '1'      # (env: A/B on one box)
        self._skip_fused_ok = os.environ.get('SMA_NO_SKIPFUSE', '0') != '1'
        self._split_ok = True
        self._skipfuse_min = int(os.environ.get('SMA_SKIPFUSE_MIN_COUT', '128'))
        if ae_path is not None:
            self.load_state_dict(torch.load(ae_path, map_location='cpu')['params_ema'])
        for module in (fix_modules or []):
            for n, p in self.named_parameters():
                if n == module or n.startswith(module + '.'):
                    p.requires_grad = False

    # ------------------------------------------------------------------------------------------
    # parameter inventory
    # ------------------------------------------------------------------------------------------
    def _declare_res(self, name, cin, cout):
        self.declare_norm(name + '.norm1', cin)
        self.declare_conv(name + '.conv1', cout, cin, 3)
        self.declare_norm(name + '.norm2', cout)
        self.declare_conv(name + '.conv2', cout, cout, 3)
        if cin != cout:
            self.declare_conv(name + '.conv_out', cout, cin, 1)

    def _declare_blocks(self, prefix, layout):
        for i, (kind, cin, cout) in enumerate(layout):
            n = f'{prefix}.blocks.{i}'
            if kind == 'conv':
                self.declare_conv(n, cout, cin, 3)
            elif kind == 'res':
                self._declare_res(n, cin, cout)
            elif kind == 'attn':
                self.declare_norm(n + '.norm', cin)
                for k in ('q', 'k', 'v', 'proj_out'):
                    self.declare_conv(f'{n}.{k}', cin, cin, 1)
            elif kind in ('down', 'up'):
                self.declare_conv(n + '.conv', cin, cin, 3)
            elif kind == 'norm':
                self.declare_norm(n, cin)

    def _declare_transformer(self, name, E):
        for a in ('self_attn', 'cross_attn'):
            b = 1.0 / math.sqrt(E)
            self.declare(f'{name}.{a}.in_proj_weight', (3 * E, E), lambda t: t.uniform_(-b, b))
            self.declare(f'{name}.{a}.in_proj_bias', (3 * E,), lambda t: t.zero_())
            self.declare_linear(f'{name}.{a}.out_proj', E, E)
        self.declare_conv(name + '.conv1', 2 * E, E, 3)
        self.declare_conv(name + '.conv2', E, 2 * E, 3)
        for k in ('norm1', 'norm2', 'norm3'):
            self.declare_norm(f'{name}.{k}', E)

    def _declare_all(self):
        Ea, Em = self.Ea, self.Em
        self._declare_blocks('encoder', self.enc_layout)
        self._declare_blocks('generator', self.gen_layout)
        self.declare_conv('app_feat_emb_32', Ea, 256, 1)
        self.declare_conv('to_app_feat_32', 256, Ea, 1)
        for s in (64, 128, 256):
            p, c = s // 32, self.channels[s]
            self.declare_linear(f'app_feat_emb_{s}.1', Ea, c * p * p)
            self.declare_linear(f'to_app_feat_{s}.0', c * p * p, Ea)
        self.declare('quantize_app.embedding.weight', (self.n_codes_app, Ea),
                     lambda t: t.uniform_(-1.0 / self.n_codes_app, 1.0 / self.n_codes_app))   # vqgan_arch.py:31
        for s in (64, 128, 256):
            c = self.channels[s]
            self._declare_res(f'fuse_convs_dict.{s}.encode_enc', 2 * c, c)
            for br in ('scale', 'shift'):
                self.declare_conv(f'fuse_convs_dict.{s}.{br}.0', c, c, 3)
                self.declare_conv(f'fuse_convs_dict.{s}.{br}.2', c, c, 3)
            self.declare_conv(f'fuse_ms_dict.{s}', c, c, 3)
        self.declare('position_emb_app', (self.L, Ea), lambda t: t.zero_())                      # :266-267 (1024 rows at 256x256)
        self.declare('position_emb_motion', (self.L, Em), lambda t: t.zero_())
        self.declare('quantize_motion.embedding.weight', (self.n_codes_motion, Em),
                     lambda t: t.uniform_(-1.0 / self.n_codes_motion, 1.0 / self.n_codes_motion))
        self.declare_conv('motion_emb.0', Em, 2, 3)
        self.declare_conv('motion_emb.1.conv', Em, Em, 3)
        self._declare_res('motion_emb.2', Em, Em)
        for i in range(2):
            self._declare_transformer(f'motion_block.{i}', Em)
        self.declare_conv('to_motion.0.conv', Em, Em, 3)        # training-only head (:290-292); kept for strict loading
        self._declare_res('to_motion.1', Em, Em)
        self.declare_norm('to_motion.2', Em)
        self.declare_conv('to_motion.3', 2, Em, 3)
        self.declare_conv('BasicMotionEncoder.convc1', 128, Em, 1)
        self.declare_conv('BasicMotionEncoder.convc2', 96, 128, 3)
        self.declare_conv('BasicMotionEncoder.convf1', 128, 2, 7)
        self.declare_conv('BasicMotionEncoder.convf2', 64, 128, 3)
        self.declare_conv('BasicMotionEncoder.conv', 126, 160, 3)
        for i, s in enumerate(self.SCALES):
            self.declare_conv(f'to_context.{i}', 192, self.channels[s], 1)
        self.declare_conv('refine.convc1', 128, 192, 3)
        self.declare_conv('refine.conv1', 128, 256, 3)
        self.declare_conv('refine.conv2', 2, 128, 3)
        self.declare_conv('refine.convo1', 128, 256, 3)
        self.declare_conv('refine.convo2', 1, 128, 3)
        for i in range(2):
            self._declare_transformer(f'app_block.{i}', Ea)
        for s in self.SCALES:
            self.declare_conv(f'warped_source_enc_{s}', Em, self.channels[s], 1)
        self.declare_conv('driving_kp_enc', Em, self.num_kp, 1)
        self.declare_conv('motion_query_enc_1', Em, 2 * Em, 1)
        self.declare_conv('motion_query_enc_2', Em, 2 * Em, 1)

    # ------------------------------------------------------------------------------------------
    # weight packing (once per load): kernel layouts, fused siblings, constant codebook K/V
    # ------------------------------------------------------------------------------------------
    def _weights(self) -> dict:
        if self._packed is not None:
            return self._packed
        T = self.tensors()
        W = self._restore_packed()
        if W is not None:                 # pre-packed blob of these exact weights: no pack kernel is launched
            self._packed, self._T, self._src_cache, self._pack_saved = W, T, None, True
            return W
        W: Dict[str, object] = {}

        def pc(name):
            W[name] = ops.pack_conv(T[name + '.weight'], T[name + '.bias'])

        def pres(name, cin, cout):
            pc(name + '.conv1'); pc(name + '.conv2')
            if cin != cout:
                pc(name + '.conv_out')

        for prefix, layout in (('encoder', self.enc_layout), ('generator', self.gen_layout)):
            for i, (kind, cin, cout) in enumerate(layout):
                n = f'{prefix}.blocks.{i}'
                if kind == 'conv':
                    pc(n)
                    if cout <= 4 and cin % 64 == 0 and T[n + '.weight'].shape[-1] == 3:      # the image head: its 9 taps as columns of a pointwise layer (ops.conv_tapsum)
                        W[n + '.tap'] = ops.pack_conv_tapcols(T[n + '.weight'])
                elif kind == 'res':
                    pres(n, cin, cout)
                elif kind == 'attn':
                    W[n + '.qkv'] = ops.pack_conv_cat([T[f'{n}.{k}.weight'] for k in 'qkv'], [T[f'{n}.{k}.bias'] for k in 'qkv'])
                    pc(n + '.proj_out')
                elif kind in ('down', 'up'):
                    pc(n + '.conv')
        pc('app_feat_emb_32'); pc('to_app_feat_32')
        for s in (64, 128, 256):
            pc(f'app_feat_emb_{s}.1'); pc(f'to_app_feat_{s}.0')
            W[f'app_feat_emb_{s}.1'] = W[f'app_feat_emb_{s}.1'].as_patch(s // 32)
            c = self.channels[s]
            pres(f'fuse_convs_dict.{s}.encode_enc', 2 * c, c)
            W[f'fuse_convs_dict.{s}.ss0'] = ops.pack_conv_cat(
                [T[f'fuse_convs_dict.{s}.{b}.0.weight'] for b in ('scale', 'shift')],
                [T[f'fuse_convs_dict.{s}.{b}.0.bias'] for b in ('scale', 'shift')])
            pc(f'fuse_convs_dict.{s}.scale.2'); pc(f'fuse_convs_dict.{s}.shift.2'); pc(f'fuse_ms_dict.{s}')
            # shift.2 over the SFT branch and fuse_ms over the compensated feature write the same tensor (w = 1): one conv over two input tensors
            W[f'fuse_convs_dict.{s}.shift2ms'] = ops.pack_conv(
                torch.cat([T[f'fuse_convs_dict.{s}.shift.2.weight'], T[f'fuse_ms_dict.{s}.weight']], dim=1),
                T[f'fuse_convs_dict.{s}.shift.2.bias'] + T[f'fuse_ms_dict.{s}.bias'])
        # the 2-channel pixel-unit flow is kept in a 32-channel zero-padded buffer: both convs reading it run on the tensor cores
        W['motion_emb.0'] = ops.pack_conv(T['motion_emb.0.weight'], T['motion_emb.0.bias'], pad_cin=32)
        # the 7x7 conv over the same 2 channels is unfolded instead (ops.im2col_small): one 1x1 conv of depth 128 instead of 49 taps x 32 padded channels
        W['BasicMotionEncoder.convf1'] = ops.pack_conv_unfolded(T['BasicMotionEncoder.convf1.weight'], T['BasicMotionEncoder.convf1.bias'], 128)
        pc('motion_emb.1.conv'); pres('motion_emb.2', self.Em, self.Em)
        # 160 input channels -> 192 (zero weights; the [cor | flo] buffer carries a zero tail): Cin % 64 == 0 puts the layer on the fp16 kernel
        # (K = 16 per tensor-core instruction instead of the tf32 kernel's 8)
        W['BasicMotionEncoder.conv'] = ops.pack_conv(T['BasicMotionEncoder.conv.weight'], T['BasicMotionEncoder.conv.bias'], pad_cin=192)
        for n in ('BasicMotionEncoder.convc1', 'BasicMotionEncoder.convc2',
                  'BasicMotionEncoder.convf2', 'refine.convc1',
                  'driving_kp_enc', 'motion_query_enc_1', 'motion_query_enc_2'):
            pc(n)
        # the two 3x3 output convs of RefineFlow read different halves of one buffer: one block-diagonal launch
        W['refine.conv2o2'] = ops.pack_conv_blockdiag([T['refine.conv2.weight'], T['refine.convo2.weight']],
                                                      [T['refine.conv2.bias'], T['refine.convo2.bias']])
        big = torch.zeros((3, 256, 3, 3), device=T['refine.conv2.weight'].device, dtype=torch.float32)
        big[0:2, :128] = T['refine.conv2.weight'].float(); big[2:3, 128:] = T['refine.convo2.weight'].float()
        W['refine.conv2o2.tap'] = ops.pack_conv_tapcols(big)
        W['refine.conv1o1'] = ops.pack_conv_cat([T['refine.conv1.weight'], T['refine.convo1.weight']],
                                                [T['refine.conv1.bias'], T['refine.convo1.bias']])
        for i, s in enumerate(self.SCALES):
            pc(f'to_context.{i}'); pc(f'warped_source_enc_{s}')
        for blk, E, cb, pe in (('motion_block', self.Em, 'quantize_motion.embedding.weight', 'position_emb_motion'),
                               ('app_block', self.Ea, 'quantize_app.embedding.weight', 'position_emb_app')):
            codes = T[cb].float().contiguous()
            pos64 = T[pe].double()
            for i in range(2):
                n = f'{blk}.{i}'
                W[n + '.self_in'] = ops.pack_conv(T[n + '.self_attn.in_proj_weight'], T[n + '.self_attn.in_proj_bias'])
                W[n + '.cross_in'] = ops.pack_conv(T[n + '.cross_attn.in_proj_weight'], T[n + '.cross_attn.in_proj_bias'])
                W[n + '.self_out'] = ops.pack_conv(T[n + '.self_attn.out_proj.weight'], T[n + '.self_attn.out_proj.bias'])
                W[n + '.cross_out'] = ops.pack_conv(T[n + '.cross_attn.out_proj.weight'], T[n + '.cross_attn.out_proj.bias'])
                pc(n + '.conv1'); pc(n + '.conv2')
                # (LN(x) + pos) W^T = LN(x) W^T + pos W^T: the positional term of the q / k projections is frame-invariant - one (L, 3E) per-token bias
                # [pos Wq^T | pos Wk^T | 0] added as a batch-stride-0 residual lets ONE linear produce q | k | v from LN(x) (and LayerNorm write one tensor)
                wi = T[n + '.self_attn.in_proj_weight'].double(); wc = T[n + '.cross_attn.in_proj_weight'].double()
                pq = torch.zeros((pos64.shape[0], 3 * E), device=codes.device, dtype=torch.float32)
                pq[:, :2 * E] = (pos64 @ wi[:2 * E].t()).float()
                W[n + '.pos_qkv'] = pq
                W[n + '.pos_q2'] = (pos64 @ wc[:E].t()).float().contiguous()
                # codebook keys/values are frame-invariant: project all rows once (prefix-sliceable for the split)
                kv = ops.linear(codes.view(1, -1, E), W[n + '.cross_in'].cols(E, 2 * E), exact=True)
                W[n + '.ctx_kv'] = kv[0]                              # (n_codes, 2E): K | V
                if E == 256 and codes.shape[0] % 256 == 0:            # ... and, for the tensor-core attention, their fp16 hi / lo operand images per prefix length
                    for n_ctx in range(codes.shape[0] // 4, codes.shape[0] + 1, codes.shape[0] // 4):
                        W[n + f'.ctx_kvimg.{n_ctx}'] = ops.attn_kv_images(kv[0][:n_ctx, :E], kv[0][:n_ctx, E:])
        self._packed = W
        self._T = T
        self._src_cache = None
        return W

    # ------------------------------------------------------------------------------------------
    # VQGAN blocks
    # ------------------------------------------------------------------------------------------
    def _gn(self, name, x):
        return ops.groupnorm_stats(x, self._T[name + '.weight'], self._T[name + '.bias'], 32, 1e-6)

    def _gnp(self, name):
        """(gamma, beta) of a GroupNorm: passed as conv2d(gn=...) so that the producing convolution's epilogue delivers the statistics."""
        return self._T[name + '.weight'], self._T[name + '.bias']

    # Every block takes `stats` (GroupNorm scale / shift of its input, already produced by the previous convolution's epilogue, or None) and `want`
    # ((gamma, beta) of the GroupNorm that will read its output, or None) and returns y, or (y, stats of y) when `want` is given.
    def _res(self, name, x, cin, cout, out=None, fast=False, stats=None, want=None):
        W = self._packed
        s1, h1 = stats if stats is not None else self._gn(name + '.norm1', x)
        h, (s2, h2) = ops.conv2d(x, W[name + '.conv1'], pad=1, pre=(s1, h1, 'swish'), fast=fast, gn=self._gnp(name + '.norm2'))
        # (cout = 64, the 256^2 scale, stays two launches: there the fused form is bound by the converters, which stage the 128-channel block input with a
        # full 3x3 halo for its centre tap alone - 2.59 ms against 1.30 + 0.87)
        if cin != cout and self._skip_fused_ok and cin % 64 == 0 and cout % 64 == 0 and self._skipfuse_min <= cout <= 128:
            # conv2 (3x3 over h, GroupNorm + swish prologue) and the 1x1 skip conv over the block input in ONE accumulator: the skip tensor is never written / re-read
            key = name + '.conv2+skip'
            if key not in W:
                W[key] = ops.pack_conv_plus_1x1(W[name + '.conv2'], W[name + '.conv_out'])
            try:
                return ops.conv2d(h, W[key], pad=1, pre=(s2, h2, 'swish'), x2=x, x2_1x1=True, out=out, fast=fast, gn=want)
            except ops._lib.SmaError:
                self._skip_fused_ok = False                  # (a layout the staged-input kernel declines: two convolutions from now on)
        skip = x if cin == cout else ops.conv2d(x, W[name + '.conv_out'], fast=fast)
        return ops.conv2d(h, W[name + '.conv2'], pad=1, pre=(s2, h2, 'swish'), res=skip, out=out, fast=fast, gn=want)

    def _attn(self, name, x, out=None, fast=False, stats=None, want=None):
        W = self._packed
        B, H, Wd, Cc = x.shape
        s, h = stats if stats is not None else self._gn(name + '.norm', x)
        o = None
        if Cc == 256 and (H * Wd) % 128 == 0 and ops.SPLIT_FUSE and ops.USE_TF32X3 and self._split_ok:
            # the qkv conv writes the attention kernel's fp16 hi / lo operand images in its epilogue (no fp32 q | k | v, no split pass)
            ws = ops.attn256_workspace(B, H * Wd, x.device)
            try:
                ops.conv2d(x, W[name + '.qkv'], pre=(s, h, 'none'), fast=fast, attn_split=(ws, float(int(Cc) ** (-0.5))))
                o = ops.attn256_presplit(ws, B, H * Wd)
            except ops._lib.SmaError:
                self._split_ok = False
        if o is None:
            qkv = ops.conv2d(x, W[name + '.qkv'], pre=(s, h, 'none'), fast=fast).view(B, H * Wd, 3 * Cc)
            o = ops.mha(qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:], heads=1, scale=float(int(Cc) ** (-0.5)))
        return ops.conv2d(o.view(B, H, Wd, Cc), W[name + '.proj_out'], res=x, out=out, fast=fast, gn=want)

    def _block(self, prefix, i, layout, x, out=None, fast=False, stats=None, want=None):
        kind, cin, cout = layout[i]
        n = f'{prefix}.blocks.{i}'
        W = self._packed
        if kind == 'conv':
            pre = None
            if i > 0 and layout[i - 1][0] == 'norm':          # GroupNorm (no activation) feeding the last conv
                s, h = stats if stats is not None else self._gn(f'{prefix}.blocks.{i - 1}', x)
                pre = (s, h, 'none')
            if (n + '.tap') in W and ops.TAPSUM and want is None and cout <= 4:
                return ops.conv_tapsum(ops.conv2d(x, W[n + '.tap'], pre=pre, fast=fast), W[n].bias, cout, 3, 1, out=out)
            return ops.conv2d(x, W[n], pad=1, pre=pre, out=out, fast=fast, gn=want)
        if kind == 'res':
            return self._res(n, x, cin, cout, out=out, fast=fast, stats=stats, want=want)
        if kind == 'attn':
            return self._attn(n, x, out=out, fast=fast, stats=stats, want=want)
        if kind == 'down':      # pad right/bottom by one, stride 2 (vqgan_arch.py:149-152)
            return ops.conv2d(x, W[n + '.conv'], stride=2, pad_tl=(0, 0), out_hw=(x.shape[1] // 2, x.shape[2] // 2), out=out, fast=fast, gn=want)
        if kind == 'up':
            return ops.conv2d(x, W[n + '.conv'], pad=1, upsample2=True, out=out, fast=fast, gn=want)
        raise ValueError(kind)

    def _want_after(self, prefix, layout, i):
        """(gamma, beta) of the GroupNorm that reads the output of block i (the first norm of block i + 1), or None."""
        if i + 1 >= len(layout):
            return None
        kind = layout[i + 1][0]
        n = f'{prefix}.blocks.{i + 1}'
        if kind == 'res':
            return self._gnp(n + '.norm1')
        if kind == 'attn':
            return self._gnp(n + '.norm')
        if kind == 'norm':
            return self._gnp(n)
        return None

    def _chain(self, prefix, layout, x, start, stop, stats=None, taps=None, fast=False):
        """Blocks [start, stop) in sequence, each convolution producing the GroupNorm statistics its successor needs.  -> (x, stats of x or None)"""
        for i in range(start, stop):
            if layout[i][0] == 'norm':                        # folded into the next conv's operand load; the statistics pass through
                continue
            want = self._want_after(prefix, layout, i)
            r = self._block(prefix, i, layout, x, fast=fast, stats=stats, want=want)
            x, stats = r if want is not None else (r, None)
            if taps is not None and i in taps:
                taps[i] = x
        return x, stats

    # ------------------------------------------------------------------------------------------
    # per-source work: encoder features (cached)
    # ------------------------------------------------------------------------------------------
    def clear_source_cache(self):
        self._src_cache = None

    @torch.no_grad()
    def encode_source(self, x: torch.Tensor) -> Dict[int, torch.Tensor]:
        """x (N,3,H,W) NCHW -> {H, H/2, H/4, H/8: NHWC features} (256, 128, 64, 32 at 256x256).  Cached on the tensor OBJECT (which the cache entry keeps
        alive, so that its address cannot be recycled for another clip's source) and its in-place version counter."""
        self._weights()
        c = self._src_cache
        if c is not None and c[0] is x and c[1] == x._version:
            return c[2]
        h = ops.nchw_to_nhwc(x.contiguous().float())
        taps = {2: None, 5: None, 8: None}
        h, _ = self._chain('encoder', self.enc_layout, h, 0, len(self.enc_layout), taps=taps)
        feats = {t.shape[2]: t for t in taps.values()}
        feats[h.shape[2]] = h
        self._src_cache = (x, x._version, feats)
        return feats

    @torch.no_grad()
    def decode_plain(self, x: torch.Tensor) -> torch.Tensor:
        """(B,256,32,32) NCHW latent -> (B,3,256,256): the 19 decoder blocks without the multi-scale fusion."""
        self._weights()
        h = ops.nchw_to_nhwc(x.contiguous().float())
        h, _ = self._chain('generator', self.gen_layout, h, 0, len(self.gen_layout))
        return ops.nhwc_to_nchw(h)

    @torch.no_grad()
    def _encode_taps(self, x: torch.Tensor) -> Dict[int, torch.Tensor]:
        """Encoder features after blocks 2 / 5 / 8 / 11 (NHWC), the taps of `encode_driving` and `app_codebook_loss`
        (appmotioncodebook_arch.py:327,364-371,433-439): the 32x32 entry is the output of block 11 - the first attention block at that scale -
        NOT the latent of the last block that `forward` warps."""
        self._weights()
        h = ops.nchw_to_nhwc(x.contiguous().float())
        taps = {2: None, 5: None, 8: None, 11: None}
        self._chain('encoder', self.enc_layout, h, 0, 12, taps=taps)
        return {t.shape[2]: t for t in taps.values()}

    def encode_driving(self, x):
        """Reference API (appmotioncodebook_arch.py:364-371): NCHW feature dict keyed by resolution string."""
        return {str(s): ops.nhwc_to_nchw(t) for s, t in self._encode_taps(x).items()}

    # ------------------------------------------------------------------------------------------
    # training-path forward values: VQ lookups inside the forward, `to_motion`, `app_codebook_loss`.  No gradients.
    # ------------------------------------------------------------------------------------------
    def _to_motion(self, x: torch.Tensor) -> torch.Tensor:
        """self.to_motion = Upsample, ResBlock, GroupNorm, conv3x3 Em -> 2 (appmotioncodebook_arch.py:290-292) on NHWC (B,tg,tg,Em) -> (B,fg,fg,2)."""
        W, T = self._packed, self.tensors()
        if 'to_motion.3' not in W:
            for n in ('to_motion.0.conv', 'to_motion.1.conv1', 'to_motion.1.conv2', 'to_motion.3'):
                W[n] = ops.pack_conv(T[n + '.weight'], T[n + '.bias'])
        x = ops.conv2d(x, W['to_motion.0.conv'], pad=1, upsample2=True)
        x = self._res('to_motion.1', x, self.Em, self.Em)
        sc, sh = self._gn('to_motion.2', x)
        return ops.conv2d(x, W['to_motion.3'], pad=1, pre=(sc, sh, 'none'))

    @torch.no_grad()
    def app_codebook_loss(self, x, weight=1, fuse_list=None):
        """Reference API (appmotioncodebook_arch.py:429-469), forward values: x = the driving ("gt") frames (B,3,H,W) ->
        ([[app_recon, app_feat_original, quant_app, app_feat, feat_com] per scale 32..256] (NCHW), [codebook loss per scale] (device scalars))."""
        if self.R != 1:
            raise NotImplementedError('app_codebook_loss: 256x256 only (the reference hard-codes the 32x32 token grid, :452-453)')
        W, T = self._weights(), self.tensors()
        feats = self._encode_taps(x)
        codes = T['quantize_app.embedding.weight'].float().contiguous()
        recon, losses = [], []
        for s0 in self.SCALES:
            f = feats[s0]
            B = f.shape[0]
            if s0 == 32:
                emb = ops.conv2d(f, W['app_feat_emb_32'])
            else:
                emb = ops.conv2d(f, W[f'app_feat_emb_{s0}.1'], stride=s0 // 32)
            quant, loss, _ = ops.vq_quantize(emb, codes, self._n_ctx(self.n_codes_app, s0), self.beta)

            def back(t):
                if s0 == 32:
                    return ops.conv2d(t, W['to_app_feat_32'])
                return ops.conv2d(t, W[f'to_app_feat_{s0}.0'], d2s=s0 // 32)
            recon.append([ops.nhwc_to_nchw(v) for v in (back(quant), back(emb), quant, emb, f)])
            losses.append(loss)
        return recon, losses

    # ------------------------------------------------------------------------------------------
    # codebook transformer layer (appmotioncodebook_arch.py:88-126) on (B,1024,E) tokens
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _n_ctx(codebook_size: int, s: int) -> int:
        """Codebook split = shared prefix of codebook_size//4 * k rows, k = 1..4 for scales 32..256 (appmotioncodebook_arch.py:374,405,473,514)."""
        return codebook_size // 4 * (int(math.log2(s)) - 4)

    def _transformer(self, name, t, E, n_ctx, pos, key_mask=None, fast=False):
        W, T = self._packed, self._T
        B = t.shape[0]
        L, tg = self.L, self.tg
        # E = 256: the projections write the attention kernels' fp16 hi / lo operand images in their epilogue (no fp32 q | k | v, no split pass)
        fuse = E == 256 and self.n_head == 8 and L % 128 == 0 and ops.SPLIT_FUSE and ops.USE_MH_F16 and ops.USE_TF32X3 and self._split_ok and (name + '.pos_qkv') in W
        a = None
        if fuse:
            u, _ = ops.layernorm(t, T[name + '.norm1.weight'], T[name + '.norm1.bias'])
            ws = ops.attn_workspace(B, B, L, L, t.device)
            try:
                ops.linear(u, W[name + '.self_in'], res=W[name + '.pos_qkv'].unsqueeze(0).expand(B, -1, -1), fast=fast, attn_split=(ws, 32 ** -0.5))
                a = ops.mha_presplit(ws, B, L, L, key_mask=key_mask)
            except ops._lib.SmaError:
                self._split_ok = False; fuse = False
        if a is not None:
            pass
        elif (name + '.pos_qkv') in W:
            u, _ = ops.layernorm(t, T[name + '.norm1.weight'], T[name + '.norm1.bias'])
            qkv = ops.linear(u, W[name + '.self_in'], res=W[name + '.pos_qkv'].unsqueeze(0).expand(B, -1, -1), fast=fast)
        else:       # (a pack cache written before the positional term was folded)
            u, uq = ops.layernorm(t, T[name + '.norm1.weight'], T[name + '.norm1.bias'], pos)
            qkv = torch.empty((B, L, 3 * E), device=t.device, dtype=torch.float32)
            ops.linear(uq, W[name + '.self_in'].cols(0, 2 * E), out=qkv[..., :2 * E], fast=fast)
            ops.linear(u, W[name + '.self_in'].cols(2 * E, E), out=qkv[..., 2 * E:], fast=fast)
        if a is None:
            a = ops.mha(qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:], heads=self.n_head, key_mask=key_mask, fast=fast)
        t = ops.linear(a, W[name + '.self_out'], res=t, fast=fast)
        kv = W[name + '.ctx_kv']
        a = None
        if fuse and n_ctx % 64 == 0:
            u2, _ = ops.layernorm(t, T[name + '.norm2.weight'], T[name + '.norm2.bias'])
            ws = ops.attn_workspace(B, 1, L, n_ctx, t.device)
            ops.linear(u2, W[name + '.cross_in'].cols(0, E), res=W[name + '.pos_q2'].unsqueeze(0).expand(B, -1, -1), fast=fast, attn_split=(ws, 32 ** -0.5))
            img = W.get(name + f'.ctx_kvimg.{n_ctx}')              # the codebook k / v operand images are frame-invariant: split once per weight load
            if img is not None:
                a = ops.mha_presplit(ws, B, L, n_ctx, kv_images=img)
            else:
                a = ops.mha_presplit(ws, B, L, n_ctx, k=kv[:n_ctx, :E], v=kv[:n_ctx, E:])
        elif (name + '.pos_q2') in W:
            u2, _ = ops.layernorm(t, T[name + '.norm2.weight'], T[name + '.norm2.bias'])
            qc = ops.linear(u2, W[name + '.cross_in'].cols(0, E), res=W[name + '.pos_q2'].unsqueeze(0).expand(B, -1, -1), fast=fast)
        else:
            _, uq = ops.layernorm(t, T[name + '.norm2.weight'], T[name + '.norm2.bias'], pos, want_y=False)
            qc = ops.linear(uq, W[name + '.cross_in'].cols(0, E), fast=fast)
        if a is None:
            a = ops.mha(qc, kv[:n_ctx, :E], kv[:n_ctx, E:], heads=self.n_head, fast=fast)
        t = ops.linear(a, W[name + '.cross_out'], res=t, fast=fast)
        u, _ = ops.layernorm(t, T[name + '.norm3.weight'], T[name + '.norm3.bias'])
        f = ops.conv2d(u.view(B, tg, tg, E), W[name + '.conv1'], pad=1, act='gelu', fast=fast)
        return ops.conv2d(f, W[name + '.conv2'], pad=1, res=t.view(B, tg, tg, E), fast=fast).view(B, L, E)

    # ------------------------------------------------------------------------------------------
    # stage 3m: motion codebook compensation (appmotioncodebook_arch.py:373-427, 129-168)
    # ------------------------------------------------------------------------------------------
    def _motion_comp(self, m_prev, occ_prev, warp0, qcat, s, collect=None, warp0g=None):
        W, T = self._packed, self._T
        B = m_prev.shape[0]
        dev = m_prev.device
        Em = self.Em
        fg, tg, s0 = self.fg, self.tg, s // self.R             # flow grid, token grid, nominal scale (module names, codebook prefix)
        fs = ops.fast('s3m')
        z = torch.empty((B, fg, fg, 256), device=dev, dtype=torch.float32)        # [BME out 126 | flow_px 2 | refine.convc1 128]
        ops.flow_to_px(m_prev, z[..., 126:128])
        flow_px = torch.zeros((B, fg, fg, 32), device=dev, dtype=torch.float32)      # [flow_px 2 | zero padding]
        ops.flow_to_px(m_prev, flow_px[..., 0:2])
        mf = ops.conv2d(flow_px, W['motion_emb.0'], pad=1, fast=fs)
        mf = ops.conv2d(mf, W['motion_emb.1.conv'], stride=2, pad_tl=(0, 0), out_hw=(tg, tg), fast=fs)
        ops_out = qcat[..., :Em]
        self._res('motion_emb.2', mf, Em, Em, out=ops_out, fast=fs)                        # qcat = [m_feat | query_feat]
        if collect is not None and collect.get('_train'):
            collect[f'm_feat_{s}'] = ops_out.contiguous()                                      # (the buffer is reused by the next scale)
        t = ops.conv2d(qcat, W['motion_query_enc_2'], fast=fs).view(B, self.L, Em)
        for i in range(2):
            t = self._transformer(f'motion_block.{i}', t, Em, self._n_ctx(self.n_codes_motion, s0), T['position_emb_motion'], fast=fs)
        mfeat = ops.resize_ac(t.view(B, tg, tg, Em), (fg, fg))
        cf = torch.empty((B, fg, fg, 192), device=dev, dtype=torch.float32)       # [cor 96 | flo 64 | zero 32]
        cf[..., 160:].zero_()
        cor = ops.conv2d(mfeat, W['BasicMotionEncoder.convc1'], act='relu', fast=fs)
        ops.conv2d(cor, W['BasicMotionEncoder.convc2'], pad=1, act='relu', out=cf[..., :96], fast=fs)
        flo = ops.conv2d(ops.im2col_small(flow_px, 2, 7, 3, 128), W['BasicMotionEncoder.convf1'], act='relu', fast=fs)
        ops.conv2d(flo, W['BasicMotionEncoder.convf2'], pad=1, act='relu', out=cf[..., 96:160], fast=fs)
        ops.conv2d(cf, W['BasicMotionEncoder.conv'], pad=1, act='relu', out=z[..., :126], fast=fs)
        wc = W[f'to_context.{int(math.log2(s0)) - 5}']
        if s >= 4 * fg:        # relu(conv1x1) then a 4x bilinear down-sampling: evaluate the pointwise layer only at the sampled neighbours (a quarter of the pixels), bit-identical
            g4 = warp0g if warp0g is not None else ops.gather_bil4(warp0, (fg, fg))
            ctx = ops.blend_bil4(ops.conv2d(g4, wc, act='relu', fast=fs), (s, s))
        else:
            ctx = ops.conv2d(warp0, wc, act='relu', fast=fs)
            if s != fg:
                ctx = ops.resize_ac(ctx, (fg, fg))
        ops.conv2d(ctx, W['refine.convc1'], pad=1, act='relu', out=z[..., 128:], fast=fs)
        f = ops.conv2d(z, W['refine.conv1o1'], pad=1, act='relu', fast=fs)                 # [flow branch 128 | occlusion branch 128]
        r = torch.empty((B, fg, fg, 4), device=dev, dtype=torch.float32)
        if ops.TAPSUM:         # 3 outputs: the 9 taps as 27 columns of a pointwise layer + a gather-sum (9x fewer tensor-core instructions)
            ops.conv_tapsum(ops.conv2d(f, W['refine.conv2o2.tap'], fast=fs), W['refine.conv2o2'].bias, 3, 3, 1, out=r[..., 0:3])
        else:
            ops.conv2d(f, W['refine.conv2o2'], pad=1, out=r[..., 0:3], fast=fs)            # [delta-flow 2 | delta-occlusion 1]
        return ops.flow_update(m_prev, occ_prev, r) + (r,)

    # ------------------------------------------------------------------------------------------
    # stage 3a: appearance codebook compensation (appmotioncodebook_arch.py:472-544)
    # ------------------------------------------------------------------------------------------
    def _app_comp(self, feat, m_com, s, out=None, gn=None):
        W, T = self._packed, self._T
        B = feat.shape[0]
        tg, s0 = self.tg, s // self.R
        mask = ops.motion_ignore_mask(m_com, (tg, tg))
        fa = ops.fast('s3a')
        if s0 == 32:
            tok = ops.conv2d(feat, W['app_feat_emb_32'], fast=fa)
        else:
            p = s0 // 32
            tok = ops.conv2d(feat, W[f'app_feat_emb_{s0}.1'], stride=p, fast=fa)
        tok = tok.view(B, self.L, self.Ea)
        n_ctx = self._n_ctx(self.n_codes_app, s0)
        tok = self._transformer('app_block.0', tok, self.Ea, n_ctx, T['position_emb_app'], key_mask=mask, fast=fa)
        tok = self._transformer('app_block.1', tok, self.Ea, n_ctx, T['position_emb_app'], fast=fa)
        tok = tok.view(B, tg, tg, self.Ea)
        if s0 == 32:
            return ops.conv2d(tok, W['to_app_feat_32'], out=out, fast=fa)
        r = ops.conv2d(tok, W[f'to_app_feat_{s0}.0'], d2s=s0 // 32, out=out, fast=fa, gn=gn)      # (gn: GroupNorm statistics of the un-patchified output, for Fuse_sft_block)
        return r[0] if gn is not None else r

    # ------------------------------------------------------------------------------------------
    # the per-driving-frame body (appmotioncodebook_arch.py:556-764, inference=True)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, feats: Dict[int, torch.Tensor], deformation: torch.Tensor, occlusion: torch.Tensor,
                 kp_heat_nhwc: torch.Tensor, w: float = 1.0, collect: Optional[dict] = None) -> dict:
        """feats: encode_source() output (batch 1 = shared by all frames, or batch B).
        deformation (B,64,64,2), occlusion (B,64,64) post-sigmoid, kp_heat_nhwc (B,64,64,15)   [128x128 grids for the 512x512 variant].
        Returns NHWC tensors: out (B,H,W,3), lq_feat (B,H/8,W/8,256), lists of motions / occlusions."""
        W = self._weights()
        B = deformation.shape[0]
        dev = deformation.device
        Em = self.Em

        def src(s):
            f = feats[s]
            return f if f.shape[0] == B else f.expand(B, -1, -1, -1)

        motions, occs, residuals = [deformation.contiguous()], [], []
        fg, tg, R = self.fg, self.tg, self.R
        occ_prev = occlusion.contiguous().view(B, fg, fg)
        qk = torch.empty((B, tg, tg, 2 * Em), device=dev, dtype=torch.float32)     # [warped-source query | driving kp feat]
        qcat = torch.empty((B, tg, tg, 2 * Em), device=dev, dtype=torch.float32)   # [motion feat | query feat]
        ops.conv2d(ops.resize_ac(kp_heat_nhwc, (tg, tg)), W['driving_kp_enc'], act='relu', out=qk[..., Em:])

        def compensate(s, out=None, gn=None):
            nonlocal occ_prev
            f = src(s)
            if collect is None and s >= 4 * fg:
                # the un-occluded query warp of this scale is only ever SAMPLED (32x32 query resize, to_context's 64x64 resize): evaluate it at those
                # samples' neighbours instead of materialising s x s pixels (bit-identical: the blend uses the resize kernel's weights and order)
                warp0 = None
                w32 = ops.blend_bil4(ops.warp_occlude_gather(f, motions[-1], None, (tg, tg)), (s, s))
                warp0g = ops.warp_occlude_gather(f, motions[-1], None, (fg, fg))
            else:
                warp0 = ops.warp_occlude(f, motions[-1], None)
                w32 = warp0 if s == tg else ops.resize_ac(warp0, (tg, tg))
                warp0g = None
            ops.conv2d(w32, W[f'warped_source_enc_{s // R}'], act='relu', out=qk[..., :Em])
            ops.conv2d(qk, W['motion_query_enc_1'], out=qcat[..., Em:])
            m_com, occ, r = self._motion_comp(motions[-1], occ_prev, warp0, qcat, s, collect, warp0g=warp0g)
            motions.append(m_com); occs.append(occ); residuals.append(r)
            occ_prev = occ
            warped = ops.warp_occlude(f, m_com, occ)
            enc = self._app_comp(warped, m_com, s, out=out, gn=gn)
            if collect is not None:
                collect[f'warp0_{s}'], collect[f'warped_{s}'], collect[f'app_{s}'] = warp0, warped, enc
            return enc

        x = compensate(tg)
        lq_feat = x
        fuse_at = {9: 64, 12: 128, 15: 256}
        cat = None
        fgen = ops.fast('gen')
        stats = None                                     # GroupNorm scale / shift of x, when the convolution that produced x delivered them
        for i in range(len(self.gen_layout)):
            if self.gen_layout[i][0] == 'norm':
                continue
            want = self._want_after('generator', self.gen_layout, i)
            if i in fuse_at and w > 0:
                s0 = fuse_at[i]                      # nominal scale: names; s: the actual feature size
                s = s0 * R
                c = self.channels[s0]
                cat = torch.empty((B, s, s, 2 * c), device=dev, dtype=torch.float32)   # [enc | dec] for Fuse_sft_block
                n = f'fuse_convs_dict.{s0}'
                # GroupNorm(32) over [enc | dec] = 16 groups per half, none straddling: each half's statistics come out of the epilogue of the
                # convolution that writes it (the decoder block's last conv; the un-patchifying conv of the appearance compensation)
                g1, b1 = self._gnp(n + '.encode_enc.norm1')
                sc_cat = torch.empty((B, 2 * c), device=dev, dtype=torch.float32); sh_cat = torch.empty((B, 2 * c), device=dev, dtype=torch.float32)
                x = self._block('generator', i, self.gen_layout, x, out=cat[..., c:], fast=fgen, stats=stats,
                                want=(g1[c:], b1[c:], 16, sc_cat[:, c:], sh_cat[:, c:]))[0]
                enc = compensate(s, out=cat[..., :c], gn=(g1[:c], b1[:c], 16, sc_cat[:, :c], sh_cat[:, :c]))
                fsft = ops.fast('sft')
                e = self._res(n + '.encode_enc', cat, 2 * c, c, fast=fsft, stats=(sc_cat, sh_cat))
                ss = ops.conv2d(e, W[n + '.ss0'], pad=1, act='leaky', fast=fsft)                  # [scale.0 | shift.0]
                scale = ops.conv2d(ss[..., :c], W[n + '.scale.2'], pad=1, fast=fsft)
                # dec + w * (dec * scale + shift) in the epilogue of the `shift.2` conv; the decoder half is read in place (channel slice of `cat`)
                r = None
                if collect is None and float(w) == 1.0 and self._two_tensor_ok and fsft == ops.fast('ms') and (n + '.shift2ms') in W:
                    # dec + (dec * scale + shift) + fuse_ms(enc) with shift + fuse_ms in ONE accumulator: the SFT result is never written and re-read
                    try:
                        r = ops.conv2d(ss[..., c:], W[n + '.shift2ms'], pad=1, res=x, sft=(scale, 1.0), fast=fsft, x2=enc, gn=want)
                    except ops._lib.SmaError:
                        self._two_tensor_ok = False                                    # (layout the staged-input kernel declines: two convolutions from now on)
                if r is None:
                    xf = ops.conv2d(ss[..., c:], W[n + '.shift.2'], pad=1, res=x, sft=(scale, float(w)), fast=fsft)
                    if collect is not None:
                        collect[f'sft_{s}'] = xf.clone()                               # (the next conv accumulates in place)
                    r = ops.conv2d(enc, W[f'fuse_ms_dict.{s0}'], pad=1, res=xf, out=xf, fast=ops.fast('ms'), gn=want)
                x, stats = r if want is not None else (r, None)
                if collect is not None:
                    collect[f'fused_{s}'] = x
            else:
                r = self._block('generator', i, self.gen_layout, x, fast=fgen, stats=stats, want=want)
                x, stats = r if want is not None else (r, None)
        return {'out': x, 'lq_feat': lq_feat, 'out_occ': occs, 'deformation_list': motions, 'residuals': residuals}

    # ------------------------------------------------------------------------------------------
    # reference call surface
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, dense_motion, w=1, inference=False, vis_app_before_comp=False, gt=None, visualize_app_feat=False):
        if isinstance(dense_motion['occlusion_map'], list):
            raise NotImplementedError('multi-mask occlusion lists are not part of options/test.yml')
        feats = self.encode_source(x)
        deformation = dense_motion['deformation'].float()
        B = deformation.shape[0]
        heat = dense_motion.get('_driving_kp_heatmap_nhwc')
        if heat is None:
            heat = ops.nchw_to_nhwc(dense_motion['driving_kp_heatmap'].contiguous().float())
        occ = dense_motion['occlusion_map'].contiguous().float().view(B, self.fg, self.fg)
        collect = {} if inference else {'_train': True}
        r = self.generate(feats, deformation, occ, heat, float(w), collect=collect)
        half = (deformation.shape[1] - 1.0) / 2.0
        scales = [s * self.R for s in self.SCALES if f'warped_{s * self.R}' in collect]
        before = [ops.nhwc_to_nchw(collect[f'warped_{s}']) for s in scales]
        after = [ops.nhwc_to_nchw(collect[f'app_{s}']) for s in scales]
        out = {
            'out': ops.nhwc_to_nchw(r['out']),
            '_out_nhwc': r['out'],
            'lq_feat': ops.nhwc_to_nchw(r['lq_feat']),
            'out_occ': [o.view(B, 1, self.fg, self.fg) for o in r['out_occ']],
            'deformation_list': r['deformation_list'],
            'res_deform_list': [q[..., 0:2] / half for q in r['residuals']],
            # the reference computes the occluded final warp a third time for this list (:604,609-615,700,705-719): same values
            'deform_feat_list': before,
            'app_comp_list': after,
            'app_before_comp_list': before,
        }
        if visualize_app_feat:      # vis_original_size=True: the query is the warped feature, the result the compensated one (:497-500,534-536)
            out['app_query_feat_list'], out['app_comp_feat_list'] = before, after
        if vis_app_before_comp:     # the plain decoder run on the un-compensated 32x32 feature (:654-655,661-662)
            out['x_before_app_32'] = self.decode_plain(before[0])
        if not inference:
            # forward VALUES of the training branch (:379-386,424-427,580-587,641-662,676-683,749-757); gradients are outside the path (no autograd graph)
            codes = self._T['quantize_motion.embedding.weight'].float().contiguous()
            recon, losses = [], []
            for s0 in self.SCALES:
                quant, loss, _ = ops.vq_quantize(collect[f'm_feat_{s0 * self.R}'], codes, self._n_ctx(self.n_codes_motion, s0), self.beta)
                recon.append(self._to_motion(quant) / half)
                losses.append(loss)
            out['out_lr'] = [self.decode_plain(out['lq_feat'])]
            out['motion_recon_list'] = recon
            out['codebook_loss_motion_list'] = losses
            if gt is not None:
                out['app_recon_list'], out['codebook_loss_app_list'] = self.app_codebook_loss(gt)
        return out
