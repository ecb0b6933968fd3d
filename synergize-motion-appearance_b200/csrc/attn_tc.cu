// Stage 3 on the tensor cores: flash-style softmax(Q K^T * scale [+ key mask]) V with both contractions on tcgen05
// (fp32-faithful 3xTF32: x = hi + lo, hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) for the codebook transformer
// layers (head dim 32: appearance, head dim 4: motion; appmotioncodebook_arch.py:97-116).
//
// CTA = 128 queries of one (frame, head); key blocks of 64.  13 warps:
//   warps 0-3 and 8-11  softmax: query row r (= TMEM lane r) is shared by two threads, one per 32-key half of the block
//              (a warp may only touch TMEM lanes 32*(warp%4)..+31).  Per key block: tcgen05.ld the scores (already in
//              log2 units), row max exchanged between the halves through shared memory, p = 2^(s - m) with ex2.approx,
//              tf32 hi / lo split, store K-major SWIZZLE_128B into one of two P buffers (A operand of P.V).  Only when a
//              row's running max moved is the O accumulator rescaled in TMEM (ld, mul, st) - after the first blocks that
//              is rare, so the softmax runs a block ahead of the tensor core.  At the end: O / l -> global.
//   warp 4     MMA issue: S_j = Q K_j^T (M128 x N64 x K=D) into one of two TMEM score buffers, O += P_j V_j
//              (M128 x N=D x K64).  The score MMA of block j+1 is issued before the P.V MMA of block j so that the
//              tensor core works while the softmax warps are busy.
//   warps 5-7,12  producers: Q once (pre-scaled like nn.MultiheadAttention scales q), then per key block K_j (K-major rows)
//              and V_j^T (head-dim rows x 64 keys, K-major) split into hi / lo, plus the additive key mask.
// All operand tiles use the layout already proven by the convolution kernels (K-major, 128-byte rows, SWIZZLE_128B).
// An all-masked row gives NaN, like softmax over an all -inf row in the reference.
#include "sma_common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int BQ = 128, BKV = 64;
constexpr int ATT_THREADS = 416;         // 13 warps: 0-3 softmax A | 4 MMA | 5-7,12 producers | 8-11 softmax B

struct AttP {
  const float* q; const float* k; const float* v; const uint8_t* mask; float* out;
  long long kv_bs;
  int ldq, ldk, ldv, ldo, L, S;
  float scale;
};

// byte offset of (row, 16-byte chunk) inside a K-major SWIZZLE_128B tile whose base is 1 KB aligned
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// round-to-nearest tf32 of a finite value (the cvt.rna instruction costs two more instructions for its inf/nan handling)
__device__ __forceinline__ float tf32_rn_finite(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1) mha_tc_kernel(const AttP p) {
  constexpr int DO = D < 16 ? 16 : D;                   // N of the P.V product (UMMA needs N >= 16 at M = 128)
  constexpr int KSTEPS_QK = (D + 7) / 8;                // k-steps of the score product (zero padded to a multiple of 8)
  constexpr uint32_t Q_IMG = BQ * 128, K_IMG = BKV * 128, VT_IMG = DO * 128, P_IMG = BQ * 128;
  constexpr uint32_t OFF_Q = 0;                                         // [hi | lo]
  constexpr uint32_t OFF_K = OFF_Q + 2 * Q_IMG;                         // 2 stages x [hi | lo]
  constexpr uint32_t OFF_V = OFF_K + 2 * 2 * K_IMG;                     // 2 stages x 2 key chunks x [hi | lo]   (V^T: DO rows x 32 keys per chunk)
  constexpr uint32_t OFF_P = OFF_V + 2 * 2 * 2 * ((VT_IMG + 1023u) & ~1023u);   // 2 buffers x 2 key chunks x [hi | lo]
  constexpr uint32_t VT_PAD = (VT_IMG + 1023u) & ~1023u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // 224 KB of tiles + < 3 KB static = the 227 KB opt-in limit exactly
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  __shared__ uint8_t s_mask[4][BKV];                       // 4 deep: block j+4 overwrites block j's row only after softmax(j) has finished
  __shared__ float s_red[2][2][BQ];                     // [block parity][key half][row]: row max (and finally row sum) exchange
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - smem_u32(smem_raw));            // generic pointer to the aligned base (zero fill only)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nblk = p.S / BKV;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, k_full0 = bar0 + 8, k_empty0 = bar0 + 24, s_full0 = bar0 + 40, p_full0 = bar0 + 56, pv_done0 = bar0 + 72,
                 v_full0 = bar0 + 88, v_empty0 = bar0 + 104;
  constexpr uint32_t TMEM_COLS = 256;                   // 2 x 64 score columns + DO output columns
  constexpr uint32_t O_COL = 128;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 128);
    for (int s = 0; s < 2; s++) {
      mbar_init(k_full0 + 8 * s, 128); mbar_init(k_empty0 + 8 * s, 1); mbar_init(v_full0 + 8 * s, 128); mbar_init(v_empty0 + 8 * s, 1);
      mbar_init(s_full0 + 8 * s, 1);
      mbar_init(p_full0 + 8 * s, 256); mbar_init(pv_done0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  if (D < 32) {   // padded operand tiles: everything the producers never write must read as zero
    const uint32_t total16 = (OFF_P) / 16;
    for (uint32_t i = threadIdx.x; i < total16; i += ATT_THREADS) reinterpret_cast<float4*>(gbase)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4 || (warp >= 8 && warp < 12)) {
    // =============================== softmax / correction / epilogue ===============================
    const int half = warp >= 8 ? 1 : 0;                  // which 32 keys of every 64-key block this thread handles
    const int r = (warp & 3) * 32 + lane;                // query row of this thread = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float m_run = -CUDART_INF_F, l_run = 0.f;
    for (int j = 0; j < nblk; j++) {
      const int sb = j & 1;
      mbar_wait(s_full0 + 8 * sb, (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld32(lane_addr + (uint32_t)(sb * BKV + half * 32), sv);
      tmem_ld_wait();
      float mloc = -CUDART_INF_F;
      if (p.mask) {
#pragma unroll
        for (int i = 0; i < 32; i++) { float s_ = s_mask[j & 3][half * 32 + i] ? -CUDART_INF_F : __uint_as_float(sv[i]); sv[i] = __float_as_uint(s_); }
      }
#pragma unroll
      for (int i = 0; i < 32; i++) mloc = fmaxf(mloc, __uint_as_float(sv[i]));
      s_red[sb][half][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");     // the two halves of every row exchange their block maxima
      const float mx = fmaxf(fmaxf(m_run, mloc), s_red[sb][half ^ 1][r]);
      const float base = mx == -CUDART_INF_F ? 0.f : mx; // every key so far masked: p = 2^(-inf) = 0, no NaN from inf - inf
      const float alpha = ex2_approx(m_run - base);      // m_run = -inf -> 0
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) { float pv = ex2_approx(__uint_as_float(sv[i]) - base); psum += pv; sv[i] = __float_as_uint(pv); }
      l_run = l_run * alpha + psum;
      const bool moved = mx != m_run;
      m_run = mx;
      if (j >= 2) mbar_wait(pv_done0 + 8 * sb, ((j >> 1) - 1) & 1);    // P buffer `sb` was last read by P.V of block j-2
      {
        const uint32_t img_hi = sbase + OFF_P + (uint32_t)sb * 4 * P_IMG + (uint32_t)half * 2 * P_IMG, img_lo = img_hi + P_IMG;
#pragma unroll
        for (int q4 = 0; q4 < 8; q4++) {
          float x0 = __uint_as_float(sv[q4 * 4]), x1 = __uint_as_float(sv[q4 * 4 + 1]), x2 = __uint_as_float(sv[q4 * 4 + 2]), x3 = __uint_as_float(sv[q4 * 4 + 3]);
          float h0 = tf32_rn_finite(x0), h1 = tf32_rn_finite(x1), h2 = tf32_rn_finite(x2), h3 = tf32_rn_finite(x3);
          const uint32_t off = sw_off(r, q4);
          sts128(img_hi + off, h0, h1, h2, h3);
          sts128(img_lo + off, x0 - h0, x1 - h1, x2 - h2, x3 - h3);
        }
      }
      if (half == 0 && j > 0 && __any_sync(0xffffffffu, moved)) {     // warp-uniform: tcgen05.ld/st are warp collectives
        mbar_wait(pv_done0 + 8 * ((j - 1) & 1), ((j - 1) >> 1) & 1);   // O must be stable: P.V of block j-1 retired
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32(lane_addr + O_COL, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(lane_addr + O_COL, o);
        tmem_st_wait();
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(p_full0 + 8 * sb);
    }
    // final: O / l  (row sum = sum of the two halves; they followed the same running max)
    s_red[nblk & 1][half][r] = l_run;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (half == 0) {
      const float l_tot = l_run + s_red[nblk & 1][1][r];
      mbar_wait(pv_done0 + 8 * ((nblk - 1) & 1), ((nblk - 1) >> 1) & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(lane_addr + O_COL, o);
      tmem_ld_wait();
      const float inv = 1.f / l_tot;                     // l = 0 (all keys masked): 0 * inf = NaN, the reference's behaviour
      if (q0 + r < p.L) {
        float* ob = p.out + ((long long)b * p.L + q0 + r) * p.ldo + h * D;
#pragma unroll
        for (int i = 0; i < D; i += 4)
          *reinterpret_cast<float4*>(ob + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, __uint_as_float(o[i + 2]) * inv,
                                                           __uint_as_float(o[i + 3]) * inv);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =============================== MMA issue ===============================
    const uint32_t idesc_s = idesc_tf32_m128(BKV), idesc_o = idesc_tf32_m128(DO);
    const uint64_t dq_hi = make_desc(sbase + OFF_Q), dq_lo = make_desc(sbase + OFF_Q + Q_IMG);
    auto issue_scores = [&](int j) {
      const int st = j & 1;
      mbar_wait(k_full0 + 8 * st, (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dk_hi = make_desc(sbase + OFF_K + (uint32_t)st * 2 * K_IMG), dk_lo = make_desc(sbase + OFF_K + (uint32_t)st * 2 * K_IMG + K_IMG);
      const uint32_t d_s = tmem_base + (uint32_t)(st * BKV);
      if (elect_one_sync()) {
#pragma unroll
        for (int k4 = 0; k4 < KSTEPS_QK; k4++) {
          const uint64_t ko = (uint64_t)(k4 * 2);
          tc_mma_tf32(d_s, dq_lo + ko, dk_hi + ko, idesc_s, k4 != 0);
          tc_mma_tf32(d_s, dq_hi + ko, dk_lo + ko, idesc_s, 1u);
          tc_mma_tf32(d_s, dq_hi + ko, dk_hi + ko, idesc_s, 1u);
        }
        tc_commit(s_full0 + 8 * st);
        tc_commit(k_empty0 + 8 * st);                     // K stage reusable as soon as the score product retires
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_scores(0);
    for (int j = 0; j < nblk; j++) {
      if (j + 1 < nblk) issue_scores(j + 1);             // its score buffer was drained by softmax(j-1), which p_full(j-1) confirmed
      const int st = j & 1;
      mbar_wait(v_full0 + 8 * st, (j >> 1) & 1);
      mbar_wait(p_full0 + 8 * st, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const uint32_t p_hi = sbase + OFF_P + (uint32_t)st * 4 * P_IMG + (uint32_t)c * 2 * P_IMG;
          const uint32_t v_hi = sbase + OFF_V + (uint32_t)(st * 2 + c) * 2 * VT_PAD;
          const uint64_t dp_hi = make_desc(p_hi), dp_lo = make_desc(p_hi + P_IMG), dv_hi = make_desc(v_hi), dv_lo = make_desc(v_hi + VT_PAD);
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t ko = (uint64_t)(k4 * 2);
            tc_mma_tf32(tmem_base + O_COL, dp_lo + ko, dv_hi + ko, idesc_o, (j | c | k4) != 0);
            tc_mma_tf32(tmem_base + O_COL, dp_hi + ko, dv_lo + ko, idesc_o, 1u);
            tc_mma_tf32(tmem_base + O_COL, dp_hi + ko, dv_hi + ko, idesc_o, 1u);
          }
        }
        tc_commit(pv_done0 + 8 * st);
        tc_commit(v_empty0 + 8 * st);
      }
      __syncwarp();
    }
  } else {
    // =============================== producers ===============================
    const int pt = (warp == 12 ? 96 : (warp - 5) * 32) + lane;   // 0..127 (warps 5,6,7,12)
    const float* qb = p.q + ((long long)b * p.L + q0) * p.ldq + h * D;
    const float* kb = p.k + (long long)b * p.kv_bs + h * D;
    const float* vb = p.v + (long long)b * p.kv_bs + h * D;
    const float qs = p.scale * 1.4426950408889634f;      // scores in log2 units: softmax uses ex2 directly
    if (D == 32) {
      const int cq = pt & 7, prow = pt >> 3;             // 16 rows per pass, 8 lanes per 128-byte row
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int row = u * 16 + prow;
        t[u] = (q0 + row < p.L) ? __ldg(reinterpret_cast<const float4*>(qb + (long long)row * p.ldq + cq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int row = u * 16 + prow;
        float x0 = t[u].x * qs, x1 = t[u].y * qs, x2 = t[u].z * qs, x3 = t[u].w * qs;
        float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
        const uint32_t off = sw_off(row, cq);
        sts128(sbase + OFF_Q + off, h0, h1, h2, h3);
        sts128(sbase + OFF_Q + Q_IMG + off, x0 - h0, x1 - h1, x2 - h2, x3 - h3);
      }
    } else {
      const int row = pt;                                // one float4 (= the whole head) per row
      float4 t = (q0 + row < p.L) ? __ldg(reinterpret_cast<const float4*>(qb + (long long)row * p.ldq)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float x0 = t.x * qs, x1 = t.y * qs, x2 = t.z * qs, x3 = t.w * qs;
      float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
      const uint32_t off = sw_off(row, 0);
      sts128(sbase + OFF_Q + off, h0, h1, h2, h3);
      sts128(sbase + OFF_Q + Q_IMG + off, x0 - h0, x1 - h1, x2 - h2, x3 - h3);
    }
    fence_async_smem();
    mbar_arrive(q_full);
    for (int j = 0; j < nblk; j++) {
      const int st = j & 1; const int s0 = j * BKV;
      const uint32_t k_hi = sbase + OFF_K + (uint32_t)st * 2 * K_IMG, k_lo = k_hi + K_IMG;
      if (D == 32) {
        const int cq = pt & 7, prow = pt >> 3;
        float4 tk[4], tv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int key = s0 + u * 16 + prow;
          tk[u] = __ldg(reinterpret_cast<const float4*>(kb + (long long)key * p.ldk + cq * 4));
          tv[u] = __ldg(reinterpret_cast<const float4*>(vb + (long long)key * p.ldv + cq * 4));
        }
        mbar_wait(k_empty0 + 8 * st, ((j >> 1) & 1) ^ 1u);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int kr = u * 16 + prow;                  // key row inside the block
          float h0 = tf32_rna(tk[u].x), h1 = tf32_rna(tk[u].y), h2 = tf32_rna(tk[u].z), h3 = tf32_rna(tk[u].w);
          const uint32_t off = sw_off(kr, cq);
          sts128(k_hi + off, h0, h1, h2, h3);
          sts128(k_lo + off, tk[u].x - h0, tk[u].y - h1, tk[u].z - h2, tk[u].w - h3);
        }
        if (pt < BKV) s_mask[j & 3][pt] = p.mask ? p.mask[(long long)b * p.S + s0 + pt] : (uint8_t)0;
        fence_async_smem();
        mbar_arrive(k_full0 + 8 * st);
        mbar_wait(v_empty0 + 8 * st, ((j >> 1) & 1) ^ 1u);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int kr = u * 16 + prow;
          // V^T: row d = 4*cq + i, column = key (chunk kr/32, position kr%32)
          const uint32_t vt_hi = sbase + OFF_V + (uint32_t)(st * 2 + (kr >> 5)) * 2 * VT_PAD, vt_lo = vt_hi + VT_PAD;
          const int kc = kr & 31;
          const float vv[4] = {tv[u].x, tv[u].y, tv[u].z, tv[u].w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int d = cq * 4 + i;
            const uint32_t off2 = sw_off(d, kc >> 2) + (uint32_t)(kc & 3) * 4u;
            float hv = tf32_rna(vv[i]);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vt_hi + off2), "f"(hv) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vt_lo + off2), "f"(vv[i] - hv) : "memory");
          }
        }
      } else {
        float4 tk = make_float4(0.f, 0.f, 0.f, 0.f), tv = tk;
        if (pt < BKV) {
          tk = __ldg(reinterpret_cast<const float4*>(kb + (long long)(s0 + pt) * p.ldk));
          tv = __ldg(reinterpret_cast<const float4*>(vb + (long long)(s0 + pt) * p.ldv));
        }
        mbar_wait(k_empty0 + 8 * st, ((j >> 1) & 1) ^ 1u);
        if (pt < BKV) {
          const int kr = pt;
          float h0 = tf32_rna(tk.x), h1 = tf32_rna(tk.y), h2 = tf32_rna(tk.z), h3 = tf32_rna(tk.w);
          const uint32_t off = sw_off(kr, 0);
          sts128(k_hi + off, h0, h1, h2, h3);
          sts128(k_lo + off, tk.x - h0, tk.y - h1, tk.z - h2, tk.w - h3);
          s_mask[j & 3][pt] = p.mask ? p.mask[(long long)b * p.S + s0 + pt] : (uint8_t)0;
        }
        fence_async_smem();
        mbar_arrive(k_full0 + 8 * st);
        mbar_wait(v_empty0 + 8 * st, ((j >> 1) & 1) ^ 1u);
        if (pt < BKV) {
          const int kr = pt;
          const uint32_t vt_hi = sbase + OFF_V + (uint32_t)(st * 2 + (kr >> 5)) * 2 * VT_PAD, vt_lo = vt_hi + VT_PAD;
          const int kc = kr & 31;
          const float vv[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t off2 = sw_off(i, kc >> 2) + (uint32_t)(kc & 3) * 4u;
            float hv = tf32_rna(vv[i]);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vt_hi + off2), "f"(hv) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(vt_lo + off2), "f"(vv[i] - hv) : "memory");
          }
        }
      }
      fence_async_smem();
      mbar_arrive(v_full0 + 8 * st);
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int D>
int launch_att(const AttP& p, int B, int heads, cudaStream_t st) {
  constexpr int DO = D < 16 ? 16 : D;
  constexpr int VT_PAD = (DO * 128 + 1023) & ~1023;
  constexpr int smem = 2 * BQ * 128 + 4 * BKV * 128 + 8 * VT_PAD + 8 * BQ * 128;
  static SmaDevOnce once;
  if (int rc = sma_opt_in_smem(once, mha_tc_kernel<D>, smem)) return rc;
  mha_tc_kernel<D><<<dim3(p.L / BQ, heads, B), ATT_THREADS, smem, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

}  // namespace

// returns SMA_ERR_UNSUPPORTED when the shape is not eligible (the caller then uses the CUDA-core kernel in attn.cu)
int sma_mha_tc_try(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, long long kv_bs, int B, int L, int S, int heads, int D,
                   float scale, const uint8_t* mask, float* out, int ldo, cudaStream_t st) {
  if ((D != 32 && D != 4) || (L % BQ) || (S % BKV) || S < BKV) return SMA_ERR_UNSUPPORTED;
  AttP p;
  p.q = q; p.k = k; p.v = v; p.mask = mask; p.out = out; p.kv_bs = kv_bs; p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.L = L; p.S = S;
  p.scale = scale;
  return D == 32 ? launch_att<32>(p, B, heads, st) : launch_att<4>(p, B, heads, st);
}
