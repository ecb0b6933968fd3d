/*
 * sma_b200.h - C ABI of the B200-native per-driving-frame talking-head path.
 *
 * Drop-in boundary.  The reference is a PyTorch program whose hot path is
 * stock ATen ops; its only native-op slot is basicsr/ops/ (pybind module per op, caller-allocated
 * outputs, errors surfaced as exceptions, launches on the caller's current stream:
 * basicsr/ops/dcn/src/deform_conv_ext.cpp:150-164, basicsr/ops/dcn/deform_conv.py:51-64).  This
 * library occupies that slot: every entry point takes raw device pointers, explicit
 * dims/strides, a cudaStream_t and caller-provided workspaces, and returns an int status.
 * It never allocates, never synchronises and never aborts.  The only process-wide state is a launch
 * counter and per-device one-time attributes (shared-memory opt-in, SM count), kept per device
 * ordinal so that one process may drive several GPUs; entry points act on the CURRENT device, which
 * must be the device that owns the pointers and the stream.
 *
 * All activations are fp32, channels-last (NHWC): element (b,y,x,c) of a tensor with pixel
 * stride `ld` floats and batch stride `bstride` floats lives at  p[b*bstride + (y*W+x)*ld + c].
 * A batch stride of 0 means "shared by every frame of the batch" (per-source cached tensors).
 * Each function's comment cites the reference code it replaces (paths relative to
 * /root/reference/basicsr/).
 */
#ifndef SMA_B200_H
#define SMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sma_stream_t; /* cudaStream_t */

enum {
  SMA_OK = 0,
  SMA_ERR_BAD_ARG = -1,      /* null pointer, non-positive dim, misaligned pointer            */
  SMA_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels implement                      */
  SMA_ERR_CUDA = -3,         /* launch failed; see cudaGetLastError in the caller             */
  SMA_ERR_NO_DEVICE = -4     /* device is not sm_100                                          */
};

/* Arithmetic of a dense contraction.  Every mode accumulates in fp32 (TMEM / registers).
 *   EXACT  fp32 FFMA on the CUDA cores.
 *   TF32X3 error-compensated split on tcgen05 kind::tf32: x = hi + lo, a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (fp32-faithful).
 *   F16X3  the same split with fp16 halves on kind::f16 (same 11-bit significands, twice the MACs per tensor cycle); weights are
 *          pre-scaled per output channel by a power of two, activations saturate at +-65504 and lose their lo half below 2^-25.
 *   TF32 / F16  single pass (hi*hi only): only for stages whose contribution to the 1e-3 output budget was measured negligible.
 *   F16X2  weights hi + lo, activations hi only (one MMA per k-step on the fused [hi | lo] weight tile); measured too coarse for every generator
 *          stage (profiles/r2_two_product_policy.md): no stage of the shipped policy uses it. */
enum { SMA_PREC_EXACT = 0, SMA_PREC_TF32X3 = 1, SMA_PREC_TF32 = 2, SMA_PREC_F16X3 = 3, SMA_PREC_F16 = 4, SMA_PREC_F16X2 = 5 };

enum { SMA_ACT_NONE = 0, SMA_ACT_RELU = 1, SMA_ACT_LEAKY02 = 2, SMA_ACT_GELU = 3, SMA_ACT_SIGMOID = 4,
       SMA_ACT_SWISH = 5 };

/* library / device info */
int sma_abi_version(void);                 /* bumps when a signature changes */
const char* sma_status_string(int status);
int sma_device_check(int device);          /* SMA_OK iff compute capability 10.x */
int sma_kernel_launch_count(void);         /* launches issued by this library in this process (monotonic, relaxed atomic) */

/* ---------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM (replaces every nn.Conv2d / nn.Linear on the path:
 * archs/vqgan_arch.py:144-191, archs/appmotioncodebook_arch.py:28-52,72-74,129-168,219-240,
 * utils/motion_estimator_util.py:214-231,363-380, archs/dense_motion_arch.py:26,56,
 * archs/keypoint_detector_arch.py:27-31).
 *   out[b,oy,ox,n] = act( bias[n] + sum_{ky,kx,c} W[(ky*kw+kx)*Cin+c][n] * pre(in[b, iy, ix, c]) ) (+ res)
 *   iy = oy*stride - pad_t + ky, ix = ox*stride - pad_l + kx on the (optionally nearest-x2
 *   upsampled) input; out-of-range taps contribute 0 AFTER the prologue (zero padding of the
 *   normalised tensor, as F.conv2d(pad=..) after GroupNorm does).
 *   pre(v) = act_pre(v * pre_scale[b*Cin+c] + pre_shift[b*Cin+c])  (GroupNorm-apply + swish fused
 *   into the operand load; pre_scale == NULL disables it).
 * Weights are pre-packed by sma_pack_conv_weight: row k=(ky*kw+kx)*Cin+c, `ldw` floats per row.
 * ------------------------------------------------------------------------------------------- */
typedef struct sma_conv_desc {
  const float* x;  int B, Hi, Wi, Cin;  int64_t in_bstride;  int in_ld;
  const float* w;  int ldw;  const float* bias;
  int Cout, kh, kw, stride, pad_t, pad_l;
  int upsample2;                 /* 1: conv runs on nearest-x2 upsampled input (Hi,Wi are the stored dims) */
  const float* pre_scale; const float* pre_shift; int pre_act;
  float* y;  int Ho, Wo;  int64_t out_bstride;  int out_ld;
  int act;
  const float* res;  int64_t res_bstride;  int res_ld;   /* added after the activation; may be NULL */
  int d2s;                       /* depth-to-space factor p (<=1 off): column n=(p1*p+p2)*C+c goes to pixel
                                    (oy*p+p1, ox*p+p2) channel c of a (B,Ho*p,Wo*p,C) tensor (un-patchify,
                                    appmotioncodebook_arch.py:223,230,237) */
  int out_nchw;                  /* 1: y is (B,Cout,Ho,Wo) contiguous (API-facing final image) */
  int precision;                 /* SMA_PREC_*: arithmetic of the contraction (the kernels fall back towards tf32 / exact when the
                                    shape or the available weight images do not allow the requested one) */
  const float* w_tc;             /* tf32 tensor-core weight image from sma_pack_conv_weight_tc (may be NULL) */
  const float* w_tc16;           /* fp16 tensor-core weight image from sma_pack_conv_weight_tc16 (may be NULL) */
  const float* w_ts;             /* fp16 tensor-memory-operand weight image from sma_pack_conv_weight_ts (may be NULL) */
  int tc_variant;                /* 0: library picks (tensor-memory-operand kernel when the weights of a 128-channel block fit in tensor memory, else
                                    the persistent halo kernel for stride 1, else the gather kernel); bit 0: force the gather kernel; bits 1-3:
                                    timing experiments (results invalid): no weight loads / no halo loads / no epilogue; bit 4: tensor-memory-
                                    operand kernel streams its weights through the ring even when they would fit; bit 5: allow its 128-pixel streaming ring; bit 6:
                                    force that ring; bit 7: halo kernel issues three separate MMAs per k-step instead of the fused [hi | lo] weight tile;
                                    bit 8: no accumulation-bias correction in the fp16 halo kernel's epilogue */
  int kernel_used;               /* OUT: 0 CUDA-core FFMA kernel, 1 tcgen05 tf32 gather kernel, 2 tcgen05 tf32 persistent halo kernel,
                                    3 tcgen05 fp16 persistent halo kernel, 4 tcgen05 fp16 kernel with the weights in tensor memory */
  int w_tc_nt;                   /* IN: output-channel tile the tf32 image `w_tc` was packed with (0 = the default, min(256, Cout rounded up to 16));
                                    OUT in plan mode: the tile the planned kernel wants (the gather kernel narrows it when there are few rows) */
  int plan_only;                 /* 1: launch nothing; set kernel_used / w_tc_nt to what this call would use if every weight image were present
                                    (lets the binding pack only the image a layer needs) */
  /* SFT epilogue (Fuse_sft_block tail, appmotioncodebook_arch.py:50-51), aux != NULL: y = res + sft_w * (res * aux + v) with v = act(conv + bias):
   * this conv is `shift.2`, aux the output of `scale.2`, res the decoder feature.  Persistent tensor-core kernel only (SMA_ERR_UNSUPPORTED
   * otherwise: use sma_sft_combine). */
  const float* aux;  int64_t aux_bstride;  int aux_ld;  float sft_w;
  /* GroupNorm statistics of the OUTPUT fused into the epilogue (the next layer's GroupNorm would otherwise re-read the tensor from HBM): with
     gn_want != 0 the persistent tensor-core kernel writes per (frame, 32-row chunk, channel pair) partial sums {sum, sum of squares} into
     gn_partial (B * gn_chunks * Cout floats) where the 256-bit epilogue runs; gn_chunks is an OUTPUT (also of a plan_only call): 0 = this launch
     cannot produce them (other kernel / layout): the caller runs sma_groupnorm_stats on y instead.  Finish with sma_groupnorm_finalize_pairs. */
  int gn_want;  float* gn_partial;  int gn_chunks;
  /* second input tensor (may be NULL): input channels [Cin1, Cin) are read from x2 (same B, Hi, Wi; row pitch in2_ld, batch stride in2_bstride), i.e.
     conv(x, w[:, :Cin1]) + conv(x2, w[:, Cin1:]) in one accumulator - Fuse_sft_block's shift conv and the fuse_ms conv of the same scale
     (appmotioncodebook_arch.py:50-51,737-738).  TMA-staged fp16 kernel only (else SMA_ERR_UNSUPPORTED: run the two convolutions). */
  const float* x2;  int64_t in2_bstride;  int in2_ld;  int Cin1;
  /* split_ws != NULL: a flat 1x1 layer producing q (Cout = 256) or q | k | v (Cout = 768) of an E = 256 attention writes the attention kernels' fp16 hi / lo
     operand images (the layout of sma_mha_e256_fwd / sma_attn256_fwd's workspace; q pre-scaled by split_qscale = softmax scale * log2 e) instead of fp32 rows:
     y is not written; call the attention with presplit = 1 (q) or 2 (q, k, v).  Persistent tensor-core kernel only (else SMA_ERR_UNSUPPORTED). */
  void* split_ws;  float split_qscale;
  int x2_k1;                     /* 1: x2 enters through a 1x1 conv (the centre tap only): a ResBlock's conv2 (k x k over x, with its GroupNorm prologue, which then
                                    applies to x alone) + its 1x1 skip conv over the block input (archs/vqgan_arch.py:185-191).  The weight image is then packed
                                    as a 1x1 conv over the rows [x: channel chunk outer, tap inner, 64 channels][x2: 64-channel chunks] */
} sma_conv_desc;

int sma_conv2d_fwd(sma_conv_desc* d, sma_stream_t stream);
int sma_sizeof_conv_desc(void);            /* sizeof(sma_conv_desc) as compiled: bindings assert their mirror of the struct against it */
/* OIHW (or (N,K) Linear) fp32 weight on the device -> packed [kh*kw*Cin][ldw] ; optional BatchNorm(eval)
 * fold: w' = w*g/sqrt(var+eps), b' = (b-mean)*g/sqrt(var+eps)+beta (sync_batchnorm/batchnorm.py:48-53). */
int sma_pack_conv_weight(const float* w_oihw, const float* bias, int Cout, int Cin, int kh, int kw,
                         const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var,
                         float bn_eps, float* w_packed, int ldw, float* bias_out, sma_stream_t stream);

/* Tensor-core weight image: per (N-tile, 32-wide K-chunk) the tf32 "hi" part and the fp32 residual "lo" part of the
 * packed weight, each laid out as the K-major SWIZZLE_128B shared-memory tile tcgen05.mma reads, so that the kernel
 * fetches it with one bulk copy per chunk.  Needs Cin % 32 == 0.  sma_conv_weight_tc_floats gives the buffer size
 * in floats (0 when the shape is not eligible). */
int64_t sma_conv_weight_tc_floats(int Cout, int Cin, int kh, int kw, int nt /* N tile, 0 = default */);
int sma_pack_conv_weight_tc(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, int nt, float* w_tc, sma_stream_t stream);
/* fp16 variant (Cin % 64 == 0): [un-scaling factors per output column][per (N-tile, 64-channel chunk, tap): fp16 hi image | lo image
 * of w * 2^-e(column)], same SWIZZLE_128B tiles. */
int64_t sma_conv_weight_tc16_floats(int Cout, int Cin, int kh, int kw);
int sma_pack_conv_weight_tc16(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, float* w_tc16, sma_stream_t stream);
/* Tensor-memory-operand variant (Cin % 64 == 0; stride-1 convs): the weights are the A operand of the MMA and live in tensor memory,
 * the pixels are the B operand in shared memory (csrc/conv_ts.cu).  Image: [un-scaling factor per output channel, padded to a multiple
 * of 128][units of 128 rows x 64 fp16: per (block of 128 channels, 64-channel chunk, tap) the hi and the lo unit; Cout <= 64: one unit
 * whose rows 0-63 are the hi and rows 64-127 the lo halves]. */
int64_t sma_conv_weight_ts_floats(int Cout, int Cin, int kh, int kw);
int sma_pack_conv_weight_ts(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, float* w_ts, sma_stream_t stream);
/* profiling aid: {SM cycles, nanoseconds} the MMA warp of CTA 0 spent in the tile loop of the last tensor-memory-operand conv launch
 * (synchronises the device) */
int sma_debug_conv_ts_prof(long long* cycles_ns /* 8 values: cycles, ns, cycles waiting for accumulator / halo / weights, 3 spare */);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm(32, eps) statistics -> per-(b,c) scale/shift consumed by the conv prologue
 * (archs/vqgan_arch.py:14-15).  partial: workspace of B*nchunk*C*2 floats, nchunk = ceil(HW/256).
 * ------------------------------------------------------------------------------------------- */
int sma_groupnorm_stats(const float* x, int B, int HW, int C, int64_t bstride, int ld, int groups, float eps,
                        const float* gamma, const float* beta, float* partial, float* scale, float* shift,
                        sma_stream_t stream);
/* second half of the fused form: partial sums written by sma_conv2d_fwd (sma_conv_desc.gn_partial, nchunk = gn_chunks) -> scale/shift rows of pitch out_ld
 * (>= C: the statistics of one half of a concatenated tensor land in their columns of the consumer's (B, 2C) buffers - the groups of GroupNorm(32) over
 * [enc | dec] do not straddle the halves).  fold > 1: the producer was a depth-to-space convolution with fold = d2s^2 sub-pixel column blocks of C channels
 * (its Cout = fold * C; HW = its own output rows). */
int sma_groupnorm_finalize_pairs(const float* partial, int B, int nchunk, int C, int groups, int HW, int fold, float eps,
                                 const float* gamma, const float* beta, float* scale, float* shift, int out_ld, sma_stream_t stream);
/* y = act(x*scale[b,c]+shift[b,c]) elementwise (used where no conv follows, e.g. AttnBlock input) */
int sma_affine_act(const float* x, int B, int HW, int C, int64_t bstride, int ld, const float* scale,
                   const float* shift, int act, float* y, int64_t y_bstride, int y_ld, sma_stream_t stream);
/* LayerNorm over the last dim of (rows,E) tokens; y = LN(x)*g+b ; yq = y + pos[row % pos_rows] (optional)
 * (archs/appmotioncodebook_arch.py:76-78,97-99,109-113,119). */
int sma_layernorm(const float* x, int rows, int E, const float* gamma, const float* beta, float eps,
                  const float* pos, int pos_rows, float* y, float* yq, sma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2: bilinear warp fused with the occlusion multiply (deform_input + occlude_input,
 * archs/appmotioncodebook_arch.py:349-362).  flow (B,hf,wf,2) and occ (B,hf,wf) are resized to
 * (H,W) on the fly (bilinear, align_corners=True); sampling is grid_sample bilinear / zeros /
 * align_corners=True.  occ may be NULL (query warp).  feat batch stride may be 0.
 * ------------------------------------------------------------------------------------------- */
int sma_warp_occlude_fwd(const float* feat, int64_t feat_bstride, int B, int H, int W, int C,
                         const float* flow, const float* occ, int hf, int wf, float* out, sma_stream_t stream);
/* the same warp evaluated only at the pixels a following bilinear (align_corners=True) down-sampling to (Hg, Wg) reads: out (B, 2Hg, 2Wg, C),
 * out[2i+a][2j+b] = warp at pixel (i_a(i), j_b(j)) - the layout sma_blend_bilinear4 consumes (Hg = Wg = 0: the full warp).  At the 256x256 scale
 * the un-occluded query warp is only ever sampled (by the 32x32 query resize and by to_context's 64x64 resize): it is never materialised. */
int sma_warp_occlude_gather_fwd(const float* feat, int64_t feat_bstride, int B, int H, int W, int C, const float* flow, const float* occ,
                                int hf, int wf, int Hg, int Wg, float* out, sma_stream_t stream);
/* bilinear align_corners=True resize of an NHWC tensor (F.interpolate call sites :390,414,418,571,671) */
int sma_resize_bilinear_ac(const float* x, int B, int Hi, int Wi, int C, int64_t in_bstride, int in_ld,
                           float* y, int Ho, int Wo, int64_t out_bstride, int out_ld, sma_stream_t stream);
/* A pointwise layer followed by a bilinear (align_corners=True) down-sampling, evaluated only where it is sampled: gather the 4 neighbours of every
 * output sample into g (B, 2Ho, 2Wo, C) [g[2i+a][2j+b] = x[i_a(i)][j_b(j)]], run the layer on g, blend (B, 2Ho, 2Wo, C') -> y (B, Ho, Wo, C') with the weights and
 * arithmetic order of sma_resize_bilinear_ac for an (Hi, Wi) source: bit-identical to layer-then-resize for a 1x1 conv + activation
 * (to_context at the 256x256 scale, archs/appmotioncodebook_arch.py:416-418: a quarter of the pixels). */
int sma_gather_bilinear4(const float* x, int B, int Hi, int Wi, int C, int64_t in_bstride, int in_ld, float* g, int Ho, int Wo, sma_stream_t stream);
int sma_blend_bilinear4(const float* g, int B, int Hi, int Wi, int C, float* y, int Ho, int Wo, int64_t out_bstride, int out_ld, sma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3: multi-head attention core softmax(Q K^T * scale [+mask]) V on projected tensors
 * (nn.MultiheadAttention inside TransformerLayer, appmotioncodebook_arch.py:97-116, and the
 * single-head AttnBlock, archs/vqgan_arch.py:233-248).  q:(B,L,*) k,v:(B or shared,S,*) ;
 * head h occupies columns [h*D,(h+1)*D).  key_mask (B,S) uint8, 1 = ignore (may be NULL).
 * D in {4,32,256}.  D in {4,32} with L % 128 == 0 and S % 64 == 0 runs on the tensor cores (tcgen05, 3xTF32 for both
 * contractions); flags bit 0 forces the exact-fp32 CUDA-core kernel; flags bit 1 allows, for D = 4 without a mask, the register-level mma form
 * (fp32-faithful hi/lo scores, P rounded to fp16: ~3e-4 absolute on unit-scale values; for stages whose error budget allows it, here S3m).
 * ------------------------------------------------------------------------------------------- */
int sma_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                int64_t kv_bstride, int B, int L, int S, int heads, int D, float scale,
                const uint8_t* key_mask, float* out, int ldo, int flags, sma_stream_t stream);

/* Single-head attention with head dim 256 (the AttnBlock, archs/vqgan_arch.py:233-248) on the tensor cores (csrc/attn256.cu): a split
 * pass writes fp16 hi / lo tile images of q (pre-scaled), k, v into `workspace` (sma_attn256_workspace_bytes), the attention kernel
 * streams them with bulk copies.  L % 128 == 0, S % 64 == 0; q:(B,L,256) k,v:(B,S,256) views with row strides ld*. */
int64_t sma_attn256_workspace_bytes(int B, int L, int S);
int sma_attn256_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int64_t q_bstride, int64_t kv_bstride,
                    int B, int L, int S, float scale, void* workspace, float* out, int ldo,
                    int presplit /* != 0: q, k, v images already in workspace (sma_conv_desc.split_ws with Cout = 768); S == L */, sma_stream_t stream);

/* 8-head attention with E = 256 (head dim 32: the appearance TransformerLayer) on the tensor cores (csrc/attn_mh.cu), same scheme as
 * sma_attn256_fwd; kv_bstride == 0: k, v (the codebook projections) are shared by every frame.  key_mask (B,S) uint8 or NULL. */
int64_t sma_mha_e256_workspace_bytes(int B, int kvB, int L, int S);
int sma_mha_e256_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int64_t q_bstride, int64_t kv_bstride,
                     int B, int L, int S, float scale, const uint8_t* key_mask, void* workspace, float* out, int ldo,
                     int presplit /* 0: split q,k,v here; 1: q images already in workspace (sma_conv_desc.split_ws); 2: q,k,v images already there;
                                     3: q images in workspace and `k` = the images sma_attn_split_kv wrote for k, v shared by all frames (kv_bstride 0) */,
                     sma_stream_t stream);
/* The operand images [k hi | k lo | v hi | v lo] (4 * S * 256 fp16) of k, v (S,256) shared by every frame - the frame-invariant codebook projections of the
 * cross-attention (archs/appmotioncodebook_arch.py:109-116) - written once per weight load; S % 64 == 0. */
int sma_attn_split_kv(const float* k, int ldk, const float* v, int ldv, int S, void* images, sma_stream_t stream);

/* VectorQuantizer lookup (archs/vqgan_arch.py:33-73): d = fl(fl(|z|^2+|e|^2) - 2 z.e), argmin with
 * lowest-index ties -> idx (int64), zq = e[idx].  z:(N,E) row-major, codebook (n_codes,E).  workspace: n_codes floats (|e|^2) or NULL; with it,
 * N >= 512 rows run as an exact-fp32 register-tiled GEMM (bit-identical distances and indices, ~20x the warp-per-row form). */
int sma_vq_lookup_fwd(const float* z, int N, int E, const float* codebook, int n_codes, int64_t* idx,
                      float* zq, float* min_dist, float* workspace, sma_stream_t stream);
/* Forward values of VectorQuantizer.forward's other returns (archs/vqgan_arch.py:76-80,88), the training-path pieces of SURVEY 8f(4):
 * zq_st = z + (zq - z) (the straight-through tensor, may be NULL) and loss[0] = beta * mean((zq - z)^2) + mean((zq - z)^2) over n elements
 * (a device scalar).  Deterministic two-stage sum; workspace: sma_vq_workspace_floats() floats. */
int sma_vq_workspace_floats(void);
int sma_vq_commit_fwd(const float* z, const float* zq, int64_t n, float beta, float* zq_st, float* workspace, float* loss, sma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Motion-estimator heads and glue
 * ------------------------------------------------------------------------------------------- */
/* AntiAliasInterpolation2d (utils/motion_estimator_util.py:599-645): x (B,3,H,W) NCHW -> y (B,H/4,W/4,3) NHWC */
int sma_antialias_down4(const float* x_nchw, int B, int C, int H, int W, const float* kernel13, float* y_nhwc,
                        int y_ld, sma_stream_t stream);
int sma_avgpool2(const float* x, int B, int H, int W, int C, float* y, int y_ld, sma_stream_t stream);
/* KPDetector tail (archs/keypoint_detector_arch.py:48-86): pred (B,h,w,ld>=5K) with columns [0,K) = kp logits,
 * [K,5K) = jacobian maps -> value (B,K,2), jacobian (B,K,2,2) */
int sma_kp_head_fwd(const float* pred, int B, int h, int w, int ld, int K, float temperature, float* value,
                    float* jacobian, sma_stream_t stream);
/* normalize_kp relative mode (demo.py:24-44): value = kp_src + s*(kp_drv-kp_drv0); jac = J_drv J_drv0^-1 J_src.
 * The movement scale s is `scale`, or *scale_dev when scale_dev != NULL (a device scalar written by sma_hull_scale: no host round trip). */
int sma_normalize_kp(const float* src_v, const float* src_j, const float* drv_v, const float* drv_j,
                     const float* drv0_v, const float* drv0_j, int B, int K, float scale, const float* scale_dev, int relative,
                     float* out_v, float* out_j, sma_stream_t stream);
/* adapt_movement_scale of normalize_kp (demo.py:26-29): sqrt(area of the convex hull of the K source key-points) /
 * sqrt(area of the hull of the K initial driving key-points), fp64 on the device, written to scale_out[0].  Replaces
 * scipy.spatial.ConvexHull(...).volume and the two .cpu() syncs per frame.  3 <= K <= 64. */
int sma_hull_scale(const float* src_v, const float* drv0_v, int K, float* scale_out, sma_stream_t stream);
/* DenseMotionNetwork input (archs/dense_motion_arch.py:65-130): heat-map differences, sparse motions and the
 * 16 warped copies of the down-sampled source (grid_sample align_corners=False) interleaved k-major into
 * hg_in (B,h,w,4*(K+1)); also the driving heat-maps drv_heat (B,h,w,K). src64 is (h,w,3) NHWC shared. */
int sma_dense_motion_prep(const float* src64, int h, int w, const float* kp_src_v, const float* kp_src_j,
                          const float* kp_drv_v, const float* kp_drv_j, int B, int K, float kp_variance,
                          float* hg_in, int hg_ld, float* drv_heat, sma_stream_t stream);
/* DenseMotionNetwork head (:134-159): logits (B,h,w,ld>=K+2): softmax over first K+1 columns, deformation =
 * sum mask_k * sparse_motion_k (recomputed), occlusion = sigmoid(column K+1). */
int sma_dense_motion_head(const float* logits, int ld, int h, int w, const float* kp_src_v, const float* kp_src_j,
                          const float* kp_drv_v, const float* kp_drv_j, int B, int K, float* deformation,
                          float* occlusion, float* mask_out, sma_stream_t stream);
/* im2col of a few-channel NHWC map (row pitch ld): out (B,H,W,Kp) with out[..][(ky*k+kx)*C + c] = x[.., y+ky-pad, x+kx-pad, c], zeros outside
 * the image and for columns >= k*k*C; turns the 7x7 conv over the 2-channel flow (archs/appmotioncodebook_arch.py:136,142) into one 1x1 conv of
 * depth Kp = 128 instead of 49 taps of a zero-padded 32-channel chunk. */
int sma_im2col_small(const float* x, int B, int H, int W, int ld, int C, int k, int pad, float* out, int Kp, sma_stream_t stream);
/* A k x k convolution with C <= 4 outputs (the image head, archs/vqgan_arch.py:349-352 last block; RefineFlow's conv2 / convo2, archs/appmotioncodebook_arch.py:163-175) as a
 * pointwise layer whose k*k*C columns are (tap, c) - run by sma_conv2d_fwd over the un-shifted input - followed by this gather-sum:
 * out[b,y,x,c] = bias[c] + sum_taps P[b, y+ky-pad, x+kx-pad, (ky*k+kx)*C + c], taps outside the map skipped (the conv's zero padding). */
int sma_conv_tapsum(const float* P, int ldp, int B, int H, int W, int C, int k, int pad, const float* bias, float* out, int out_ld, sma_stream_t stream);
/* flow glue of AppMotionCompFormer.forward (archs/appmotioncodebook_arch.py:562-601,689-710) */
int sma_flow_to_px(const float* m, int B, int h, int w, float* flow_px, int out_ld, sma_stream_t stream);
/* res: (B,h,w,res_ld) with columns 0,1 = delta-flow in pixels, 2 = delta-occlusion logit */
int sma_flow_update(const float* m_prev, const float* occ_prev, const float* res, int res_ld, int B, int h, int w,
                    float* m_new, float* occ_new, sma_stream_t stream);
/* key-padding mask of app_codebook_compensation (:487-492): m (B,h,w,2) resized to 32x32, 1 where |.|>1 */
int sma_motion_ignore_mask(const float* m, int B, int h, int w, int ht, int wt, uint8_t* mask, sma_stream_t stream);
/* Fuse_sft_block tail (:50-51): out = dec + w*(dec*scale + shift) */
int sma_sft_combine(const float* dec, const float* scale, const float* shift, float w, int64_t n, float* out,
                    sma_stream_t stream);
/* tensor2img (utils/img_util.py:42-98): NHWC fp32 in [-1,1] -> HWC uint8 (round half even), optional BGR */
int sma_to_uint8(const float* x_nhwc, int B, int H, int W, int C, int ld, int bgr, uint8_t* out, sma_stream_t stream);
/* the reference's host-side frame preparation on the device (demo.py:177-185, utils/img_util.py:13-39): uint8 HWC frames ->
 * fp32 NCHW, v = (float(u8)/255 - 0.5)/0.5, optional BGR->RGB; bit-exact with astype(float32)/255. + normalize(0.5, 0.5).
 * A quarter of the PCIe bytes of uploading fp32 frames. */
int sma_u8hwc_to_f32nchw(const uint8_t* x_hwc, int B, int H, int W, int C, int swap_rb, float* y_nchw, sma_stream_t stream);
/* (B,C,H,W) <-> (B,H,W,C) */
int sma_nchw_to_nhwc(const float* x, int B, int C, int H, int W, float* y, int y_ld, sma_stream_t stream);
int sma_nhwc_to_nchw(const float* x, int B, int C, int H, int W, int x_ld, float* y, sma_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMA_B200_H */
