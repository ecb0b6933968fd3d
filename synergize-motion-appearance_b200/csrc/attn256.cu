// Single-head attention with head dim 256 (the VQGAN AttnBlock over the 32x32 latent, vqgan_arch.py:233-248) on tcgen05:
// softmax(Q K^T * scale) V, fp32-faithful through the fp16 hi/lo split (a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32
// accumulation in tensor memory), flash-style (scores never leave the SM).
//
// Two kernels:
//  1. attn256_split_kernel: q (pre-scaled by scale*log2 e), k, v -> fp16 hi / lo images already in the shared-memory tile layout the MMAs
//     read (128-byte rows of 64 head-dim values, SWIZZLE_128B), so that the attention kernel fetches whole tiles with bulk copies:
//       q_hi / k / v : [frame][row block][4 chunks of 64 head-dim values][rows][128 B]   (row block = 128 queries / 64 keys)
//       q_lo         : [frame][row block][4 chunks][8 pieces of 16 B][128 rows]           (what tcgen05.st lanes load, coalesced)
//     K (rows = keys, 128 B along the contraction) and V (rows = keys = contraction, 128 B along N) share one layout: K is read
//     through a K-major descriptor, V through an MN-major one - no transpose anywhere.
//  2. attn256_kernel: CTA = 128 queries of one frame, key blocks of 64; 10 warps:
//       warps 0-7  softmax, two threads per query row (32 keys each; a warp may only touch TMEM lanes 32*(warp%4)..+31):
//                  tcgen05.ld the scores (log2 units), p = 2^(s - m_ref) with a LAZY reference maximum (m_ref only moves - and O is only
//                  rescaled in tensor memory - when the row maximum exceeds it by more than 8, i.e. p stays below 2^8: rescaling 256
//                  columns costs as much as a whole key block and would otherwise fire on almost every block), fp16 hi/lo split of p,
//                  swizzled store as the A operand of P.V; finally O / l -> global
//       warp 8     MMA issue (one elected lane): S = Q K_j^T as q_lo (TENSOR MEMORY operand) x k_hi + q_hi x k_lo + q_hi x k_hi
//                  (M128 x N64 x K256), O += P V_j (M128 x N256 x K64).  S_{j+1} is issued before P_j V_j.
//       warp 9     loader: one bulk copy for q_hi, then per key block the k and v images (hi | lo adjacent: 64 KB each)
//     Shared memory: q_hi 64 KB + k 64 KB + v 64 KB + p 32 KB; tensor memory: S 64 + O 256 + q_lo 128 columns.
#include "sma_common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

namespace {

constexpr int A2_BQ = 128, A2_BKV = 64, A2_D = 256;
constexpr int A2_THREADS = 320;
constexpr uint32_t A2_OFF_QH = 0, A2_OFF_K = 65536, A2_OFF_V = 131072, A2_OFF_P = 196608, A2_SMEM = 229376;
constexpr uint32_t A2_S_COL = 0, A2_O_COL = 64, A2_QL_COL = 320;
constexpr float A2_LAZY = 8.f;                         // log2 units

__device__ __forceinline__ uint32_t a2_sw_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ float a2_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void a2_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// 1. split: one thread per (frame, row, group of 8 head-dim values) of q, k and v
// ---------------------------------------------------------------------------------------------------------------
__global__ void attn256_split_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk, const float* __restrict__ v, int ldv,
                                     long long bs_q, long long bs_kv, int B, int kvB, int L, int S, float qscale, uint16_t* __restrict__ ws) {
  const long long nq = q ? (long long)B * L * 32 : 0, nk = (long long)kvB * S * 32;      // q == nullptr: the q images are already in the workspace
  const long long img_q = (long long)B * L * A2_D, img_k = (long long)kvB * S * A2_D;     // halfs per image
  uint16_t* qh = ws; uint16_t* ql = qh + img_q; uint16_t* kh = ql + img_q; uint16_t* kl = kh + img_k; uint16_t* vh = kl + img_k; uint16_t* vl = vh + img_k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nq + 2 * nk; i += (long long)gridDim.x * blockDim.x) {
    const int which = i < nq ? 0 : (i < nq + nk ? 1 : 2);
    const long long t = which == 0 ? i : (which == 1 ? i - nq : i - nq - nk);
    const int g = (int)(t & 31); const long long rowi = t >> 5;
    const int R = which == 0 ? L : S;
    const int b = (int)(rowi / R), row = (int)(rowi - (long long)b * R);
    const float* src = which == 0 ? q + b * bs_q + (long long)row * ldq + g * 8
                                  : (which == 1 ? k + b * bs_kv + (long long)row * ldk + g * 8 : v + b * bs_kv + (long long)row * ldv + g * 8);
    float4 x0 = __ldg(reinterpret_cast<const float4*>(src)), x1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    if (which == 0) { x0.x *= qscale; x0.y *= qscale; x0.z *= qscale; x0.w *= qscale; x1.x *= qscale; x1.y *= qscale; x1.z *= qscale; x1.w *= qscale; }
    uint4 hi, lo;
    split_f16x2(x0.x, x0.y, hi.x, lo.x); split_f16x2(x0.z, x0.w, hi.y, lo.y); split_f16x2(x1.x, x1.y, hi.z, lo.z); split_f16x2(x1.z, x1.w, hi.w, lo.w);
    const int RB = which == 0 ? A2_BQ : A2_BKV;                       // rows per tile
    const int rb = row / RB, r = row - rb * RB, c4 = g >> 3, ch = g & 7;
    // tile (frame, row block): 4 chunk images of RB rows x 128 B
    const long long tile = ((long long)b * (R / RB) + rb) * (4LL * RB * 64);
    const long long off = tile + (long long)c4 * RB * 64 + (a2_sw_off(r, ch) >> 1);
    if (which == 0) {
      *reinterpret_cast<uint4*>(qh + off) = hi;
      *reinterpret_cast<uint4*>(ql + tile + ((long long)(c4 * 8 + ch) * A2_BQ + r) * 8) = lo;      // piece-major for the TMEM loaders
    } else if (which == 1) {
      *reinterpret_cast<uint4*>(kh + off) = hi; *reinterpret_cast<uint4*>(kl + off) = lo;
    } else {
      *reinterpret_cast<uint4*>(vh + off) = hi; *reinterpret_cast<uint4*>(vl + off) = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. attention
// ---------------------------------------------------------------------------------------------------------------
struct A2P { const uint16_t* ws; float* out; int ldo, B, L, S; };

__global__ void __launch_bounds__(A2_THREADS, 1) attn256_kernel(const A2P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_red[2][A2_BQ];                     // [key half][row]: row max (and finally row sum) exchange
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y;
  const int nqt = p.L / A2_BQ, nblk = p.S / A2_BKV;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, k_full = bar0 + 8, k_empty = bar0 + 16, v_full = bar0 + 24, v_empty = bar0 + 32, s_full = bar0 + 40, s_free = bar0 + 48,
                 p_full = bar0 + 56, pv_done = bar0 + 64;
  const long long img_q = (long long)p.B * p.L * A2_D, img_k = (long long)p.B * p.S * A2_D;
  const uint16_t* qh = p.ws; const uint16_t* ql = qh + img_q; const uint16_t* kh = ql + img_q; const uint16_t* kl = kh + img_k;
  const uint16_t* vh = kl + img_k; const uint16_t* vl = vh + img_k;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1 + 4); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 8) {
    // =============================== softmax / correction / epilogue ===============================
    const int half = warp >> 2;                          // which 32 keys of every 64-key block (and which 128 output columns)
    const int r = (warp & 3) * 32 + lane;                // query row = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
      // q_lo -> tensor memory (A operand of the first score product): 4 chunks of 32 columns
      const uint4* src = reinterpret_cast<const uint4*>(ql + ((long long)b * nqt + qt) * (4LL * A2_BQ * 64)) + r;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; c4++) {
        uint32_t t[32];
#pragma unroll
        for (int i = 0; i < 8; i++) { uint4 x = __ldg(src + (c4 * 8 + i) * A2_BQ); t[4 * i] = x.x; t[4 * i + 1] = x.y; t[4 * i + 2] = x.z; t[4 * i + 3] = x.w; }
        tmem_st32(lane_addr + A2_QL_COL + (uint32_t)(c4 * 32), t);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    float m_ref = -CUDART_INF_F, l_run = 0.f;
    for (int j = 0; j < nblk; j++) {
      const uint32_t par = (uint32_t)j & 1u;
      mbar_wait(s_full, par);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld32(lane_addr + A2_S_COL + (uint32_t)(half * 32), sv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);                // the score buffer may be overwritten by S_{j+1}
      float mloc = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; i++) mloc = fmaxf(mloc, __uint_as_float(sv[i]));
      s_red[half][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");     // the two halves of every row exchange their block maxima
      const float mx = fmaxf(mloc, s_red[half ^ 1][r]);
      asm volatile("bar.sync 2, 256;" ::: "memory");     // ... before the next block overwrites them
      const bool need = mx > m_ref + A2_LAZY;            // identical in both halves of the row
      const float m_new = need ? mx : m_ref;
      const float alpha = need ? a2_ex2(m_ref - m_new) : 1.f;      // m_ref = -inf (first block) -> 0
      m_ref = m_new;
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) { float pv = a2_ex2(__uint_as_float(sv[i]) - m_new); psum += pv; sv[i] = __float_as_uint(pv); }
      l_run = l_run * alpha + psum;
      if (j > 0) {
        mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);      // P.V of block j-1 retired: O is stable and the P buffer is free
        if (__any_sync(0xffffffffu, need)) {             // warp-uniform: tcgen05.ld / st are warp collectives
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; c++) {
            uint32_t o[32];
            const uint32_t addr = lane_addr + A2_O_COL + (uint32_t)(half * 128 + c * 32);
            tmem_ld32(addr, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(addr, o);
          }
          tmem_st_wait();
        }
      }
      {
        const uint32_t p_hi = sbase + A2_OFF_P, p_lo = p_hi + 16384u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
          split_f16x2(__uint_as_float(sv[8 * c]), __uint_as_float(sv[8 * c + 1]), h0, l0);
          split_f16x2(__uint_as_float(sv[8 * c + 2]), __uint_as_float(sv[8 * c + 3]), h1, l1);
          split_f16x2(__uint_as_float(sv[8 * c + 4]), __uint_as_float(sv[8 * c + 5]), h2, l2);
          split_f16x2(__uint_as_float(sv[8 * c + 6]), __uint_as_float(sv[8 * c + 7]), h3, l3);
          const uint32_t off = a2_sw_off(r, half * 4 + c);
          sts128u(p_hi + off, h0, h1, h2, h3);
          sts128u(p_lo + off, l0, l1, l2, l3);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // final: O / l  (row sum = sum of the two halves; they followed the same reference maximum)
    s_red[half][r] = l_run;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float inv = 1.f / (l_run + s_red[half ^ 1][r]);
    mbar_wait(pv_done, (uint32_t)(nblk - 1) & 1u);
    tc_fence_after();
    float* ob = p.out + ((long long)b * p.L + (long long)qt * A2_BQ + r) * p.ldo + half * 128;
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
      uint32_t o[32];
      tmem_ld32(lane_addr + A2_O_COL + (uint32_t)(half * 128 + c * 32), o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(ob + c * 32 + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                                                  __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
    }
    tc_fence_before();
  } else if (warp == 8) {
    // =============================== MMA issue (one elected lane runs the whole loop) ===============================
    if (elect_one_sync()) {
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(A2_BKV >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);                 // K-major A and B
      const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(A2_D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // B (= V) MN-major
      const uint32_t s_tmem = tmem_base + A2_S_COL, o_tmem = tmem_base + A2_O_COL, ql_tmem = tmem_base + A2_QL_COL;
      // V block, MN-major SWIZZLE_128B: 64 head-dim values (128 B) contiguous per key row, 8-key groups 1 KB apart (SBO), 64-value groups
      // 8 KB apart (LBO)
      const uint64_t v_desc_bits = ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      auto issue_scores = [&](int j) {
        mbar_wait(k_full, (uint32_t)j & 1u);
        if (j > 0) mbar_wait(s_free, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 16; ks++) {
          const uint32_t c4 = ks >> 2; const uint64_t w2 = (uint64_t)((ks & 3) * 2);
          const uint64_t dq = make_desc(sbase + A2_OFF_QH + c4 * 16384u) + w2;
          const uint64_t dkh = make_desc(sbase + A2_OFF_K + c4 * 8192u) + w2, dkl = make_desc(sbase + A2_OFF_K + 32768u + c4 * 8192u) + w2;
          a2_mma_ts(s_tmem, ql_tmem + (uint32_t)(ks * 8), dkh, idesc_s, ks != 0);      // q_lo * k_hi
          tc_mma_f16(s_tmem, dq, dkl, idesc_s, 1u);                                    // q_hi * k_lo
          tc_mma_f16(s_tmem, dq, dkh, idesc_s, 1u);                                    // q_hi * k_hi
        }
        tc_commit(s_full);
        tc_commit(k_empty);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_scores(0);
      for (int j = 0; j < nblk; j++) {
        if (j + 1 < nblk) issue_scores(j + 1);
        mbar_wait(v_full, (uint32_t)j & 1u);
        mbar_wait(p_full, (uint32_t)j & 1u);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {                 // 16 keys per k-step: 32 bytes of a P row, two 8-key row groups of V
          const uint64_t dph = make_desc(sbase + A2_OFF_P) + (uint64_t)(ks * 2), dpl = make_desc(sbase + A2_OFF_P + 16384u) + (uint64_t)(ks * 2);
          const uint64_t dvh = v_desc_bits | (uint64_t)(((sbase + A2_OFF_V + (uint32_t)ks * 2048u) & 0x3FFFFu) >> 4);
          const uint64_t dvl = v_desc_bits | (uint64_t)(((sbase + A2_OFF_V + 32768u + (uint32_t)ks * 2048u) & 0x3FFFFu) >> 4);
          tc_mma_f16(o_tmem, dpl, dvh, idesc_o, (j | ks) != 0);
          tc_mma_f16(o_tmem, dph, dvl, idesc_o, 1u);
          tc_mma_f16(o_tmem, dph, dvh, idesc_o, 1u);
        }
        tc_commit(pv_done);
        tc_commit(v_empty);
      }
    }
    __syncwarp();
  } else {
    // =============================== loader ===============================
    if (lane == 0) {
      mbar_expect_tx(q_full, 65536u);
      bulk_g2s(sbase + A2_OFF_QH, qh + ((long long)b * nqt + qt) * (4LL * A2_BQ * 64), 65536u, q_full);
      for (int j = 0; j < nblk; j++) {
        const long long tile = ((long long)b * nblk + j) * (4LL * A2_BKV * 64);
        mbar_wait(k_empty, ((uint32_t)j & 1u) ^ 1u);
        mbar_expect_tx(k_full, 65536u);
        bulk_g2s(sbase + A2_OFF_K, kh + tile, 32768u, k_full);
        bulk_g2s(sbase + A2_OFF_K + 32768u, kl + tile, 32768u, k_full);
        mbar_wait(v_empty, ((uint32_t)j & 1u) ^ 1u);
        mbar_expect_tx(v_full, 65536u);
        bulk_g2s(sbase + A2_OFF_V, vh + tile, 32768u, v_full);
        bulk_g2s(sbase + A2_OFF_V + 32768u, vl + tile, 32768u, v_full);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// shared with attn_mh.cu: fp16 hi / lo tile images of q (B frames, pre-scaled by qscale), k and v (kvB frames: 1 = shared by all frames)
int sma_attn_split_launch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, long long q_bs, long long kv_bs, int B, int kvB,
                          int L, int S, float qscale, void* workspace, cudaStream_t st) {
  const long long items = (q ? (long long)B * L * 32 : 0) + 2LL * kvB * S * 32;
  int blocks = (int)((items + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
  attn256_split_kernel<<<blocks, 256, 0, st>>>(q, ldq, k, ldk, v, ldv, q_bs, kv_bs, B, kvB, L, S, qscale, reinterpret_cast<uint16_t*>(workspace));
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int64_t sma_attn256_workspace_bytes(int B, int L, int S) {
  if (B <= 0 || L <= 0 || S <= 0) return 0;
  return 2LL * A2_D * 2 * ((long long)B * L + 2LL * B * S);      // fp16 hi + lo images of q, k, v
}

extern "C" int sma_attn256_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int64_t q_bstride, int64_t kv_bstride,
                               int B, int L, int S, float scale, void* workspace, float* out, int ldo, int presplit, sma_stream_t stream) {
  // presplit != 0: the q | k | v projection's epilogue has written the operand images into `workspace` (sma_conv_desc.split_ws); q, k, v are then unused
  if (!out || !workspace || B <= 0 || L <= 0 || S <= 0 || (!presplit && (!q || !k || !v))) return SMA_ERR_BAD_ARG;
  if (presplit) { q = k = v = out; if (S != L) return SMA_ERR_BAD_ARG; }
  if ((L % A2_BQ) || (S % A2_BKV) || B > 65535) return SMA_ERR_UNSUPPORTED;
  if (((ldq | ldk | ldv | ldo) & 3) || ((q_bstride | kv_bstride) & 3)) return SMA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(workspace)) & 15)
    return SMA_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  if (!presplit) {
    int rs = sma_attn_split_launch(q, ldq, k, ldk, v, ldv, q_bstride, kv_bstride, B, B, L, S, scale * 1.4426950408889634f, workspace, st);
    if (rs != SMA_OK) return rs;
  }
  static SmaDevOnce once;
  if (int rc = sma_opt_in_smem(once, attn256_kernel, (int)A2_SMEM + 1024)) return rc;
  A2P p; p.ws = reinterpret_cast<const uint16_t*>(workspace); p.out = out; p.ldo = ldo; p.B = B; p.L = L; p.S = S;
  attn256_kernel<<<dim3(L / A2_BQ, B), A2_THREADS, A2_SMEM + 1024, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
