// Stage 3: softmax(Q K^T * scale [+ key mask]) V on already-projected tokens, fp32, flash-style
// (scores never leave the SM).  Three head sizes occur on the path:
//   D = 32  : appearance TransformerLayer (E=256, 8 heads)        appmotioncodebook_arch.py:97-116
//   D = 4   : motion TransformerLayer (E=32, 8 heads)
//   D = 256 : single-head AttnBlock over the 32x32 latent         vqgan_arch.py:233-248
// The codebook K/V (cross attention) are shared by every frame: kv batch stride 0.
// An all-masked row yields NaN, exactly like softmax over an all -inf row in the reference.
#include "sma_common.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>

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


// ---------------------------------------------------------------------------------------------
// D = 4, no mask, on the tensor cores at REGISTER level (round 2b).  The CUDA-core kernel above spends ~12.5 issue slots per score (5 FFMA for
// the dot product relative to the reference maximum, FMNMX, MUFU, FADD, 4 FFMA for P.V) and is issue-bound (ncu: issue 63 %, FMA 40 %, XU 36 %).
// tcgen05 is a poor fit for a head of 4 values (M = 128 tiles through tensor memory: the softmax threads would pay a tcgen05.ld / st round trip
// per score, which is what bounds attn_mh), but the warp-level mma keeps S and P in registers, and its K = 16 has room for the whole hi / lo split:
//   S = Q K^T : ONE m16n8k16 per 8 keys with A row = [q_hi | q_hi | q_lo | 0] and B column = [k_hi | k_lo | k_hi | 0] (4 head dims each), i.e.
//               q_hi k_hi + q_hi k_lo + q_lo k_hi: fp32-faithful scores;
//   P.V       : ONE m16n8k16 per 16 keys with B columns = [v_hi (4 dims) | v_lo (4 dims)]; the accumulator fragment of S IS the A fragment of P
//               (packed to fp16 in place); the hi and lo halves of O are added at the end (lanes tig and tig + 2).
// 12 warp-level mma per 1024 scores and ~3 further issue slots per score (FADD, MUFU, FADD, half a pack, a quarter FMNMX).  Only P is rounded to
// fp16 (relative 2^-12, random, averaged over the keys): measured 3e-4 max / 9e-6 mean absolute against fp64 on N(0,1) inputs.
// A warp owns 32 query rows (two m16 tiles) of one head, a CTA 256; K (hi / lo, 2 words per key) and V^T (hi / lo, [4][S + 8] halfs) of the head
// are converted once per CTA into shared memory.  Lazy reference maximum per row, kept equal across the four threads of a row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void split_h(float x, float& h, float& l) { h = __half2float(__float2half_rn(x)); l = x - h; }
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) mha_d4_mma_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                         const float* __restrict__ v, int ldv, long long kv_bs, int L, int S, float qscale,
                                                         float* __restrict__ out, int ldo) {
  extern __shared__ uint32_t sm32[];
  const int SP = S + 8;                                      // V^T row pitch in halfs: the 8 rows (4 dims x hi / lo) fall into different banks
  uint32_t* Kh = sm32; uint32_t* Kl = Kh + 2 * S + 16;       // [S][2] words {d0,d1}, {d2,d3}; (+16 words: the hi and lo words of a key in different banks)
  __half* Vt = reinterpret_cast<__half*>(Kl + 2 * S);         // [8][SP]: rows 0-3 v_hi dim 0..3, rows 4-7 v_lo
  const int h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  for (int f = tid; f < S; f += 256) {
    const float4 kk = __ldg(reinterpret_cast<const float4*>(k + (long long)b * kv_bs + (long long)f * ldk + h * 4));
    const float4 vv = __ldg(reinterpret_cast<const float4*>(v + (long long)b * kv_bs + (long long)f * ldv + h * 4));
    float h0, h1, h2, h3, l0, l1, l2, l3;
    split_h(kk.x, h0, l0); split_h(kk.y, h1, l1); split_h(kk.z, h2, l2); split_h(kk.w, h3, l3);
    Kh[2 * f] = pack_h2(h0, h1); Kh[2 * f + 1] = pack_h2(h2, h3); Kl[2 * f] = pack_h2(l0, l1); Kl[2 * f + 1] = pack_h2(l2, l3);
    split_h(vv.x, h0, l0); split_h(vv.y, h1, l1); split_h(vv.z, h2, l2); split_h(vv.w, h3, l3);
    Vt[0 * SP + f] = __float2half_rn(h0); Vt[1 * SP + f] = __float2half_rn(h1); Vt[2 * SP + f] = __float2half_rn(h2); Vt[3 * SP + f] = __float2half_rn(h3);
    Vt[4 * SP + f] = __float2half_rn(l0); Vt[5 * SP + f] = __float2half_rn(l1); Vt[6 * SP + f] = __float2half_rn(l2); Vt[7 * SP + f] = __float2half_rn(l3);
  }
  // Q fragments of the packed A row [q_hi | q_hi | q_lo | 0]: a0 / a1 (k = 2 tig, 2 tig + 1; rows g / g + 8) = q_hi pair (tig & 1) for every tig;
  // a2 / a3 (k = 2 tig + 8 ..) = q_lo pair tig for tig < 2, zero else.  Pre-scaled to log2 units.
  const int rbase = blockIdx.x * 256 + warp * 32;
  uint32_t qa[2][2], qc[2][2];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      qa[t][u] = 0u; qc[t][u] = 0u;
      const int r = rbase + t * 16 + u * 8 + g;
      if (r < L) {
        const float2 x = __ldg(reinterpret_cast<const float2*>(q + ((long long)b * L + r) * ldq + h * 4 + 2 * (tig & 1)));
        float a0, a1, c0, c1; split_h(x.x * qscale, a0, c0); split_h(x.y * qscale, a1, c1);
        qa[t][u] = pack_h2(a0, a1);
        if (tig < 2) qc[t][u] = pack_h2(c0, c1);
      }
    }
  __syncthreads();
  constexpr float LAZY = 8.f;
  float m_ref[2][2], lsum[2][2], o[2][4];
#pragma unroll
  for (int t = 0; t < 2; t++) { m_ref[t][0] = m_ref[t][1] = -CUDART_INF_F; lsum[t][0] = lsum[t][1] = 0.f; o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f; }
  // packed B column [k_hi | k_lo | k_hi | 0]: b0 (k = 2 tig ..) = k_hi pair tig (tig < 2) or k_lo pair tig - 2; b1 (k = 2 tig + 8 ..) = k_hi pair tig (tig < 2) or 0
  const uint32_t* kw = (tig < 2 ? Kh : Kl) + (tig & 1);
  const __half* vr = Vt + g * SP;                            // B column g of P.V: v_hi dim g (g < 4) or v_lo dim g - 4
  for (int kb = 0; kb < S; kb += 32) {
    float sc[2][4][4];                                      // [row tile][8-key sub-tile][c0..c3]
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t b0 = kw[2 * (kb + j * 8 + g)], b1 = tig < 2 ? b0 : 0u;
#pragma unroll
      for (int t = 0; t < 2; t++) {
        sc[t][j][0] = sc[t][j][1] = sc[t][j][2] = sc[t][j][3] = 0.f;
        mma_16816(sc[t][j], qa[t][0], qa[t][1], qc[t][0], qc[t][1], b0, b1);
      }
    }
    // block maxima of the four rows this thread touches, made equal across the row's four threads
    float mx[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
      mx[t][0] = fmaxf(fmaxf(fmaxf(sc[t][0][0], sc[t][0][1]), fmaxf(sc[t][1][0], sc[t][1][1])), fmaxf(fmaxf(sc[t][2][0], sc[t][2][1]), fmaxf(sc[t][3][0], sc[t][3][1])));
      mx[t][1] = fmaxf(fmaxf(fmaxf(sc[t][0][2], sc[t][0][3]), fmaxf(sc[t][1][2], sc[t][1][3])), fmaxf(fmaxf(sc[t][2][2], sc[t][2][3]), fmaxf(sc[t][3][2], sc[t][3][3])));
#pragma unroll
      for (int u = 0; u < 2; u++) {
        mx[t][u] = fmaxf(mx[t][u], __shfl_xor_sync(0xffffffffu, mx[t][u], 1));
        mx[t][u] = fmaxf(mx[t][u], __shfl_xor_sync(0xffffffffu, mx[t][u], 2));
      }
    }
    bool any = false;
#pragma unroll
    for (int t = 0; t < 2; t++) any |= (mx[t][0] > m_ref[t][0] + LAZY) | (mx[t][1] > m_ref[t][1] + LAZY);
    if (__any_sync(0xffffffffu, any)) {                      // rare after the first blocks
#pragma unroll
      for (int t = 0; t < 2; t++)
#pragma unroll
        for (int u = 0; u < 2; u++)
          if (mx[t][u] > m_ref[t][u] + LAZY) {
            const float corr = ex2f_(m_ref[t][u] - mx[t][u]);          // m_ref = -inf (first block) -> 0
            lsum[t][u] *= corr; o[t][2 * u] *= corr; o[t][2 * u + 1] *= corr; m_ref[t][u] = mx[t][u];
          }
    }
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {                         // two 16-key steps of P.V
      const int kk0 = kb + jj * 16 + 2 * tig;
      const uint32_t v0 = *reinterpret_cast<const uint32_t*>(vr + kk0), v1 = *reinterpret_cast<const uint32_t*>(vr + kk0 + 8);
#pragma unroll
      for (int t = 0; t < 2; t++) {
        float p[2][4];
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const float* s4 = sc[t][jj * 2 + x];
          p[x][0] = ex2f_(s4[0] - m_ref[t][0]); p[x][1] = ex2f_(s4[1] - m_ref[t][0]);
          p[x][2] = ex2f_(s4[2] - m_ref[t][1]); p[x][3] = ex2f_(s4[3] - m_ref[t][1]);
          lsum[t][0] += p[x][0] + p[x][1]; lsum[t][1] += p[x][2] + p[x][3];
        }
        mma_16816(o[t], pack_h2(p[0][0], p[0][1]), pack_h2(p[0][2], p[0][3]), pack_h2(p[1][0], p[1][1]), pack_h2(p[1][2], p[1][3]), v0, v1);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      float ls = lsum[t][u];
      ls += __shfl_xor_sync(0xffffffffu, ls, 1); ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      float o0 = o[t][2 * u], o1 = o[t][2 * u + 1];          // columns 2 tig, 2 tig + 1: the v_hi part for tig < 2, the v_lo part of dims 2 (tig - 2) .. for tig >= 2
      o0 += __shfl_xor_sync(0xffffffffu, o0, 2); o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
      const int r = rbase + t * 16 + u * 8 + g;
      if (tig < 2 && r < L) {
        const float inv = 1.f / ls;
        *reinterpret_cast<float2*>(out + ((long long)b * L + r) * ldo + h * 4 + 2 * tig) = make_float2(o0 * inv, o1 * inv);
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
  if (D == 4 && !key_mask && !(flags & 1) && (flags & 2) && (S % 32) == 0 && (L % 16) == 0 && (ldq & 1) == 0 && (ldo & 1) == 0) {
    // register-level tensor-core form (fp32-faithful scores, P in fp16): K hi/lo 16 B + V^T hi/lo 16 B per key in shared memory
    const int smem = (4 * S + 16) * 4 + 8 * (S + 8) * 2;
    if (smem <= 200 * 1024) {
      static SmaDevOnce oncem;
      if (int rc = sma_opt_in_smem(oncem, mha_d4_mma_kernel, 200 * 1024)) return rc;
      mha_d4_mma_kernel<<<dim3(cdiv(L, 256), heads, B), 256, smem, st>>>(q, ldq, k, ldk, v, ldv, kv_bstride, L, S, scale * 1.4426950408889634f, out, ldo);
      SMA_LAUNCH_CHECK();
      return SMA_OK;
    }
  }
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
