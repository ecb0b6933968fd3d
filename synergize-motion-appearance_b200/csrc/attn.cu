// Stage 3: softmax(Q K^T * scale [+ key mask]) V on already-projected tokens, fp32, flash-style
// (scores never leave the SM).  Three head sizes occur on the path:
//   D = 32  : appearance TransformerLayer (E=256, 8 heads)        appmotioncodebook_arch.py:97-116
//   D = 4   : motion TransformerLayer (E=32, 8 heads)
//   D = 256 : single-head AttnBlock over the 32x32 latent         vqgan_arch.py:233-248
// The codebook K/V (cross attention) are shared by every frame: kv batch stride 0.
// An all-masked row yields NaN, exactly like softmax over an all -inf row in the reference.
#include "sma_common.cuh"
#include <math_constants.h>

namespace {

// ---------------------------------------------------------------------------------------------
// D in {32, 256}: CTA = 64 queries of one (b, head); 256 threads as 16(ty) x 16(tx);
// thread owns S rows {ty+16i} x cols {tx+16j} (i,j<4) and O rows {ty+16i} x D/16 columns.
// ---------------------------------------------------------------------------------------------
template <int D, int BKV>
__global__ void __launch_bounds__(256) mha_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                  const float* __restrict__ v, int ldv, long long kv_bs, int L, int S, float scale,
                                                  const uint8_t* __restrict__ mask, float* __restrict__ out, int ldo) {
  constexpr int BQ = 64;
  constexpr int QS = D + 4;                 // padded row stride (floats), keeps 16B alignment, conflict-free LDS.128
  constexpr int PS = BKV + 4;
  constexpr int NJ = BKV / 16;              // score columns per thread
  constexpr int DV = D / 16;                // output columns per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                         // [64][QS]
  float* Ks = Qs + BQ * QS;                 // [BKV][QS]
  float* Vs = Ks + BKV * QS;                // [BKV][D]
  float* Ps = Vs + BKV * D;                 // [64][PS]
  __shared__ float mk[BKV];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const float* qb = q + ((long long)b * L + q0) * ldq + h * D;
  const float* kb = k + (long long)b * kv_bs + h * D;
  const float* vb = v + (long long)b * kv_bs + h * D;

  // load Q tile (pre-scaled, as the reference scales q before the product)
  for (int f = tid; f < BQ * (D / 4); f += 256) {
    int r = f / (D / 4), c4 = f % (D / 4);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < L) t = __ldg(reinterpret_cast<const float4*>(qb + (long long)r * ldq + c4 * 4));
    t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
    *reinterpret_cast<float4*>(Qs + r * QS + c4 * 4) = t;
  }

  float o[4][DV];
  float mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    mrow[i] = -CUDART_INF_F; lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DV; j++) o[i][j] = 0.f;
  }

  for (int s0 = 0; s0 < S; s0 += BKV) {
    __syncthreads();                         // previous tile fully consumed (also orders the Q store)
    for (int f = tid; f < BKV * (D / 4); f += 256) {
      int r = f / (D / 4), c4 = f % (D / 4);
      float4 tk = make_float4(0.f, 0.f, 0.f, 0.f), tv = tk;
      if (s0 + r < S) {
        tk = __ldg(reinterpret_cast<const float4*>(kb + (long long)(s0 + r) * ldk + c4 * 4));
        tv = __ldg(reinterpret_cast<const float4*>(vb + (long long)(s0 + r) * ldv + c4 * 4));
      }
      *reinterpret_cast<float4*>(Ks + r * QS + c4 * 4) = tk;
      *reinterpret_cast<float4*>(Vs + r * D + c4 * 4) = tv;
    }
    if (tid < BKV) mk[tid] = (s0 + tid >= S || (mask && mask[(long long)b * S + s0 + tid])) ? 1.f : 0.f;
    __syncthreads();

    // scores
    float sc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < NJ; j++) sc[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < D; d += 4) {
      float4 qa[4], ka[NJ];
#pragma unroll
      for (int i = 0; i < 4; i++) qa[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * QS + d);
#pragma unroll
      for (int j = 0; j < NJ; j++) ka[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * QS + d);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
          sc[i][j] = fmaf(qa[i].x, ka[j].x, sc[i][j]); sc[i][j] = fmaf(qa[i].y, ka[j].y, sc[i][j]);
          sc[i][j] = fmaf(qa[i].z, ka[j].z, sc[i][j]); sc[i][j] = fmaf(qa[i].w, ka[j].w, sc[i][j]);
        }
    }
    // online softmax (row statistics shared by the 16 tx lanes of a half-warp)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < NJ; j++) { if (mk[tx + 16 * j] != 0.f) sc[i][j] = -CUDART_INF_F; mx = fmaxf(mx, sc[i][j]); }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float mnew = fmaxf(mrow[i], mx);
      float corr = (mnew == -CUDART_INF_F) ? 1.f : expf(mrow[i] - mnew);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        float pv = (mnew == -CUDART_INF_F) ? 0.f : expf(sc[i][j] - mnew);
        ps += pv; Ps[(ty + 16 * i) * PS + tx + 16 * j] = pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      lrow[i] = lrow[i] * corr + ps; mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < DV; j++) o[i][j] *= corr;
    }
    __syncthreads();
    // O += P V
#pragma unroll 2
    for (int c = 0; c < BKV; c += 4) {
      float4 pa[4];
#pragma unroll
      for (int i = 0; i < 4; i++) pa[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * PS + c);
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        float vv[DV];
        if (DV >= 4) {
#pragma unroll
          for (int g = 0; g < DV / 4; g++) {
            float4 t = *reinterpret_cast<const float4*>(Vs + (c + cc) * D + g * 64 + tx * 4);
            vv[g * 4 + 0] = t.x; vv[g * 4 + 1] = t.y; vv[g * 4 + 2] = t.z; vv[g * 4 + 3] = t.w;
          }
        } else {
          float2 t = *reinterpret_cast<const float2*>(Vs + (c + cc) * D + tx * 2);
          vv[0] = t.x; vv[1] = t.y;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float pw = cc == 0 ? pa[i].x : (cc == 1 ? pa[i].y : (cc == 2 ? pa[i].z : pa[i].w));
#pragma unroll
          for (int j = 0; j < DV; j++) o[i][j] = fmaf(pw, vv[j], o[i][j]);
        }
      }
    }
  }
  // normalise and store (head-concatenated)
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int r = q0 + ty + 16 * i;
    if (r >= L) continue;
    float inv = 1.f / lrow[i];               // 0 -> inf -> 0*inf = NaN for an all-masked row (reference behaviour)
    float* ob = out + ((long long)b * L + r) * ldo + h * D;
    if (DV >= 4) {
#pragma unroll
      for (int g = 0; g < DV / 4; g++)
        *reinterpret_cast<float4*>(ob + g * 64 + tx * 4) = make_float4(o[i][g * 4] * inv, o[i][g * 4 + 1] * inv, o[i][g * 4 + 2] * inv, o[i][g * 4 + 3] * inv);
    } else {
      *reinterpret_cast<float2*>(ob + tx * 2) = make_float2(o[i][0] * inv, o[i][1] * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// D = 4: one thread per (query, head); K/V head slices staged through shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mha_d4_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                     const float* __restrict__ v, int ldv, long long kv_bs, int L, int S, float scale,
                                                     const uint8_t* __restrict__ mask, float* __restrict__ out, int ldo) {
  constexpr int TK = 256;
  __shared__ float4 Ks[TK], Vs[TK];
  __shared__ float mk[TK];
  const int h = blockIdx.y, b = blockIdx.z;
  const int r = blockIdx.x * 128 + threadIdx.x;
  float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < L) qv = __ldg(reinterpret_cast<const float4*>(q + ((long long)b * L + r) * ldq + h * 4));
  qv.x *= scale; qv.y *= scale; qv.z *= scale; qv.w *= scale;
  float m = -CUDART_INF_F, l = 0.f; float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = 0; s0 < S; s0 += TK) {
    __syncthreads();
    for (int f = threadIdx.x; f < TK; f += 128) {
      bool ok = s0 + f < S;
      Ks[f] = ok ? __ldg(reinterpret_cast<const float4*>(k + (long long)b * kv_bs + (long long)(s0 + f) * ldk + h * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      Vs[f] = ok ? __ldg(reinterpret_cast<const float4*>(v + (long long)b * kv_bs + (long long)(s0 + f) * ldv + h * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      mk[f] = (!ok || (mask && mask[(long long)b * S + s0 + f])) ? 1.f : 0.f;
    }
    __syncthreads();
    // tile maximum first, then one rescale per tile
    float mx = m;
    for (int c = 0; c < TK; c++) {
      float4 kk = Ks[c];
      float s = fmaf(qv.x, kk.x, fmaf(qv.y, kk.y, fmaf(qv.z, kk.z, qv.w * kk.w)));
      if (mk[c] == 0.f) mx = fmaxf(mx, s);
    }
    if (mx == -CUDART_INF_F) continue;
    float corr = expf(m - mx);               // m = -inf -> 0
    l *= corr; o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr; m = mx;
    for (int c = 0; c < TK; c++) {
      if (mk[c] != 0.f) continue;
      float4 kk = Ks[c];
      float s = fmaf(qv.x, kk.x, fmaf(qv.y, kk.y, fmaf(qv.z, kk.z, qv.w * kk.w)));
      float p = expf(s - m); float4 vv = Vs[c];
      l += p; o.x = fmaf(p, vv.x, o.x); o.y = fmaf(p, vv.y, o.y); o.z = fmaf(p, vv.z, o.z); o.w = fmaf(p, vv.w, o.w);
    }
  }
  if (r < L) {
    float inv = 1.f / l;
    *reinterpret_cast<float4*>(out + ((long long)b * L + r) * ldo + h * 4) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// D = 4, no mask (motion TransformerLayer): CUDA cores beat the tensor cores here - a head is 4 values, so a score costs 4 FFMA and its
// P.V contribution 4 FFMA, while tcgen05 needs K >= 16 and N >= 16 (>= 75 % padding) plus the softmax round trip through tensor memory.
// One thread = one query row of one head; the head's K and V (S x 16 B each) sit in shared memory (broadcast LDS.128); scores are
// computed directly relative to a lazy reference maximum (s - m_ref in the FFMA chain, p = ex2 of it), which only moves when a chunk
// exceeds it by more than 2^8: ~13 issue slots per (32 queries x 1 key).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// QPT query rows per thread: the broadcast K / V shared-memory loads are shared by QPT independent dot-product / exp / accumulate chains
// (ncu, round 2: the single-row version was latency-bound - short-scoreboard stalls 4.5 per issue at 35 % occupancy - not issue-bound)
template <int QPT>
__global__ void __launch_bounds__(128) mha_d4_fast_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                          const float* __restrict__ v, int ldv, long long kv_bs, int L, int S, float qscale,
                                                          float* __restrict__ out, int ldo) {
  extern __shared__ float4 kv_s[];
  float4* Ks = kv_s; float4* Vs = kv_s + S;
  const int h = blockIdx.y, b = blockIdx.z;
  const int r0 = blockIdx.x * (128 * QPT) + threadIdx.x;
  for (int f = threadIdx.x; f < S; f += 128) {
    Ks[f] = __ldg(reinterpret_cast<const float4*>(k + (long long)b * kv_bs + (long long)f * ldk + h * 4));
    Vs[f] = __ldg(reinterpret_cast<const float4*>(v + (long long)b * kv_bs + (long long)f * ldv + h * 4));
  }
  float4 qv[QPT];
#pragma unroll
  for (int j = 0; j < QPT; j++) {
    const int r = r0 + j * 128;
    qv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < L) qv[j] = __ldg(reinterpret_cast<const float4*>(q + ((long long)b * L + r) * ldq + h * 4));
    qv[j].x *= qscale; qv[j].y *= qscale; qv[j].z *= qscale; qv[j].w *= qscale;        // log2 units
  }
  __syncthreads();
  constexpr float LAZY = 8.f;
  float m_ref[QPT], l[QPT]; float4 o[QPT];
#pragma unroll
  for (int j = 0; j < QPT; j++) {
    m_ref[j] = -CUDART_INF_F; l[j] = 0.f; o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) { float4 kk = Ks[i]; m_ref[j] = fmaxf(m_ref[j], fmaf(qv[j].w, kk.w, fmaf(qv[j].z, kk.z, fmaf(qv[j].y, kk.y, qv[j].x * kk.x)))); }
  }
  for (int c0 = 0; c0 < S; c0 += 8) {
    float t[QPT][8]; float tmax[QPT];
#pragma unroll
    for (int j = 0; j < QPT; j++) tmax[j] = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 kk = Ks[c0 + i];
#pragma unroll
      for (int j = 0; j < QPT; j++) {
        t[j][i] = fmaf(qv[j].w, kk.w, fmaf(qv[j].z, kk.z, fmaf(qv[j].y, kk.y, fmaf(qv[j].x, kk.x, -m_ref[j]))));
        tmax[j] = fmaxf(tmax[j], t[j][i]);
      }
    }
    bool any = false;
#pragma unroll
    for (int j = 0; j < QPT; j++) any |= tmax[j] > LAZY;
    if (__any_sync(0xffffffffu, any)) {        // rare after the first chunks
#pragma unroll
      for (int j = 0; j < QPT; j++) {
        if (tmax[j] > LAZY) {
          const float corr = ex2f_(-tmax[j]);
          l[j] *= corr; o[j].x *= corr; o[j].y *= corr; o[j].z *= corr; o[j].w *= corr; m_ref[j] += tmax[j];
#pragma unroll
          for (int i = 0; i < 8; i++) t[j][i] -= tmax[j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 vv = Vs[c0 + i];
#pragma unroll
      for (int j = 0; j < QPT; j++) {
        const float pr = ex2f_(t[j][i]);
        l[j] += pr; o[j].x = fmaf(pr, vv.x, o[j].x); o[j].y = fmaf(pr, vv.y, o[j].y); o[j].z = fmaf(pr, vv.z, o[j].z); o[j].w = fmaf(pr, vv.w, o[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < QPT; j++) {
    const int r = r0 + j * 128;
    if (r < L) {
      const float inv = 1.f / l[j];
      *reinterpret_cast<float4*>(out + ((long long)b * L + r) * ldo + h * 4) = make_float4(o[j].x * inv, o[j].y * inv, o[j].z * inv, o[j].w * inv);
    }
  }
}

template <int D, int BKV>
int launch_mha(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, long long kv_bs, int B, int L, int S, int heads,
               float scale, const uint8_t* mask, float* out, int ldo, cudaStream_t st) {
  constexpr int smem = (64 * (D + 4) + BKV * (D + 4) + BKV * D + 64 * (BKV + 4)) * (int)sizeof(float);
  static SmaDevOnce once;           // per device: the opt-in is a (kernel, device) attribute
  if (int rc = sma_opt_in_smem(once, mha_kernel<D, BKV>, smem)) return rc;
  mha_kernel<D, BKV><<<dim3(cdiv(L, 64), heads, B), 256, smem, st>>>(q, ldq, k, ldk, v, ldv, kv_bs, L, S, scale, mask, out, ldo);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

}  // namespace

int sma_mha_tc_try(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, long long kv_bs, int B, int L, int S, int heads, int D,
                   float scale, const uint8_t* mask, float* out, int ldo, cudaStream_t st);   // attn_tc.cu

extern "C" int sma_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int64_t kv_bstride, int B, int L, int S,
                           int heads, int D, float scale, const uint8_t* key_mask, float* out, int ldo, int flags, sma_stream_t stream) {
  if (!q || !k || !v || !out || B <= 0 || L <= 0 || S <= 0 || heads <= 0) return SMA_ERR_BAD_ARG;
  if ((ldq | ldk | ldv | ldo) & 3) return SMA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)
    return SMA_ERR_UNSUPPORTED;
  if (kv_bstride & 3) return SMA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (D == 4 && !key_mask && !(flags & 1) && (S % 8) == 0 && S * 32 <= 192 * 1024) {       // K / V of one head in shared memory: 4096 keys (512x512 variant) = 128 KB
    static SmaDevOnce once2, once4;
    if (L >= 1024) {      // four rows per thread (measured: 1024 tokens 3.44 -> 2.75 ms per step with two, 512x512 variant 96 -> 45 ms with four)
      if (int rc = sma_opt_in_smem(once4, mha_d4_fast_kernel<4>, 192 * 1024)) return rc;
      mha_d4_fast_kernel<4><<<dim3(cdiv(L, 512), heads, B), 128, S * 32, st>>>(q, ldq, k, ldk, v, ldv, kv_bstride, L, S, scale * 1.4426950408889634f, out, ldo);
    } else {
      if (int rc = sma_opt_in_smem(once2, mha_d4_fast_kernel<2>, 192 * 1024)) return rc;
      mha_d4_fast_kernel<2><<<dim3(cdiv(L, 256), heads, B), 128, S * 32, st>>>(q, ldq, k, ldk, v, ldv, kv_bstride, L, S, scale * 1.4426950408889634f, out, ldo);
    }
    SMA_LAUNCH_CHECK();
    return SMA_OK;
  }
  if (!(flags & 1)) {
    int r = sma_mha_tc_try(q, ldq, k, ldk, v, ldv, kv_bstride, B, L, S, heads, D, scale, key_mask, out, ldo, st);
    if (r != SMA_ERR_UNSUPPORTED) return r;
  }
  if (D == 4) {
    mha_d4_kernel<<<dim3(cdiv(L, 128), heads, B), 128, 0, st>>>(q, ldq, k, ldk, v, ldv, kv_bstride, L, S, scale, key_mask, out, ldo);
    SMA_LAUNCH_CHECK();
    return SMA_OK;
  }
  if (D == 32) return launch_mha<32, 64>(q, ldq, k, ldk, v, ldv, kv_bstride, B, L, S, heads, scale, key_mask, out, ldo, st);
  if (D == 256) return launch_mha<256, 32>(q, ldq, k, ldk, v, ldv, kv_bstride, B, L, S, heads, scale, key_mask, out, ldo, st);
  return SMA_ERR_UNSUPPORTED;
}
