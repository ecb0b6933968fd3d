// 8-head attention of the appearance TransformerLayer (E = 256, head dim 32; appmotioncodebook_arch.py:97-116) on tcgen05 with the
// fp16 hi/lo split, flash-style.  Same building blocks as attn256.cu: pre-split tile images fetched with bulk copies, q_lo as a
// tensor-memory operand, V through an MN-major descriptor, lazy reference maximum.
//
// A 128-byte row of the tile images holds 64 head-dim values = TWO heads, so a CTA owns one (128-query tile, head PAIR, frame) and runs
// the two heads one after the other over the K / V tile stream: head x reads bytes 64x..64x+63 of each row (descriptor start + 64 B; the
// swizzle is a function of the absolute address, so a shifted start stays consistent).  Per 64-key block: S = q k^T (2 k-steps x 3
// products, N = 64), O += P V (4 k-steps x 3 products, N = 32, P read from TENSOR MEMORY).  K / V are double buffered.  80 KB of shared
// memory and 256 tensor-memory columns per CTA: two CTAs per SM.
//   warps 0-7  softmax (two threads per query row, 32 keys each; both heads), O rescale, epilogue
//   warp 8     MMA issue (one elected lane)          warp 9   loader (bulk copies)
// An all-masked row yields NaN like the reference (l = 0 -> 0 * inf).
#include "sma_common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>

int sma_attn_split_launch(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, long long q_bs, long long kv_bs, int B, int kvB,
                          int L, int S, float qscale, void* workspace, cudaStream_t st);   // attn256.cu

namespace {

constexpr int MH_BQ = 128, MH_BKV = 64, MH_E = 256;
constexpr int MH_THREADS = 320;
constexpr uint32_t MH_OFF_QH = 0, MH_OFF_K = 16384, MH_OFF_V = 49152, MH_SMEM = 81920;
constexpr uint32_t MH_S_COL = 0, MH_O_COL = 64, MH_QL_COL = 96, MH_P_COL = 128;      // S 0 | O 64 | q_lo 96 (two heads) | P hi 128, lo 160 ; 192 of 256 columns
constexpr float MH_LAZY = 8.f;

__device__ __forceinline__ float mh_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void mh_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a_tmem), "l"(b),
               "r"(idesc), "r"(acc)
               : "memory");
}

struct MhP { const uint16_t* ws; const uint16_t* kv; const uint8_t* mask; float* out; int ldo, B, kvB, L, S; };      // kv != nullptr: the k / v images live there, not behind the q images

// Two CTAs per SM (80 KB of shared memory, 256 tensor-memory columns each): the softmax of one CTA overlaps the MMA waits of the other.  The two
// heads of the pair run one after the other over the same K / V tile stream (block index g = head * nblk + j drives every barrier phase).
__global__ void __launch_bounds__(MH_THREADS, 2) attn_mh_kernel(const MhP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_red[2][MH_BQ];                     // [key half][row]
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, hp = blockIdx.y, b = blockIdx.z;
  const int nqt = p.L / MH_BQ, nblk = p.S / MH_BKV, ntot = 2 * nblk;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, s_full = bar0 + 8, s_free = bar0 + 16, p_full = bar0 + 24, pv_done = bar0 + 32;
  auto k_full = [&](int s) { return bar0 + 8u * (5 + s); };   auto k_empty = [&](int s) { return bar0 + 8u * (7 + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (9 + s); };   auto v_empty = [&](int s) { return bar0 + 8u * (11 + s); };
  const long long img_q = (long long)p.B * p.L * MH_E, img_k = (long long)p.kvB * p.S * MH_E;
  const uint16_t* qh = p.ws; const uint16_t* ql = qh + img_q; const uint16_t* kh = p.kv ? p.kv : ql + img_q; const uint16_t* kl = kh + img_k;
  const uint16_t* vh = kl + img_k; const uint16_t* vl = vh + img_k;
  const int kvb = p.kvB == 1 ? 0 : b;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1 + 4); mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(pv_done, 1);
    for (int s = 0; s < 2; s++) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 8) {
    // =============================== softmax / correction / epilogue ===============================
    const int half = warp >> 2;                          // which 32 keys of every 64-key block; also which 16 columns of O this warp rescales / stores
    const int r = (warp & 3) * 32 + lane;                // query row = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
      // q_lo of the head pair -> tensor memory: 64 fp16 = 32 columns
      const uint4* src = reinterpret_cast<const uint4*>(ql + ((long long)b * nqt + qt) * (4LL * MH_BQ * 64) + (long long)hp * (MH_BQ * 64)) + r;
      uint32_t t[32];
#pragma unroll
      for (int i = 0; i < 8; i++) { uint4 x = __ldg(src + i * MH_BQ); t[4 * i] = x.x; t[4 * i + 1] = x.y; t[4 * i + 2] = x.z; t[4 * i + 3] = x.w; }
      tmem_st32(lane_addr + MH_QL_COL, t);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    float m_ref = -CUDART_INF_F, l_run = 0.f;
    for (int g = 0; g < ntot; g++) {
      const int x = g >= nblk ? 1 : 0, j = g - x * nblk;
      const uint32_t par = (uint32_t)g & 1u;
      uint32_t sv[32];
      uint4 mk0 = make_uint4(0, 0, 0, 0), mk1 = mk0;
      if (p.mask) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (long long)b * p.S + j * MH_BKV + half * 32);
        mk0 = __ldg(mp); mk1 = __ldg(mp + 1);
      }
      mbar_wait(s_full, par);
      tc_fence_after();
      tmem_ld32(lane_addr + MH_S_COL + (uint32_t)(half * 32), sv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);                  // S may be overwritten by the next block's scores
      if (p.mask) {
        const uint32_t mw[8] = {mk0.x, mk0.y, mk0.z, mk0.w, mk1.x, mk1.y, mk1.z, mk1.w};
#pragma unroll
        for (int i = 0; i < 32; i++)
          if ((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) sv[i] = __float_as_uint(-CUDART_INF_F);
      }
      float mloc = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; i++) mloc = fmaxf(mloc, __uint_as_float(sv[i]));
      s_red[half][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");     // the two halves of every row exchange their block maxima
      const float mx = fmaxf(mloc, s_red[half ^ 1][r]);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const bool need = mx > m_ref + MH_LAZY;            // identical in both halves of the row
      const float m_new = need ? mx : m_ref;
      const float alpha = need ? mh_ex2(m_ref - m_new) : 1.f;          // m_ref = -inf (nothing seen yet) -> 0
      m_ref = m_new;
      const float base = m_new == -CUDART_INF_F ? 0.f : m_new;         // every key so far masked: p = 2^(-inf) = 0, no inf - inf
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) { float pv = mh_ex2(__uint_as_float(sv[i]) - base); psum += pv; sv[i] = __float_as_uint(pv); }
      l_run = l_run * alpha + psum;
      if (g > 0) {
        mbar_wait(pv_done, (uint32_t)(g - 1) & 1u);      // P V of the previous block retired: O is stable and the P columns are free
        if (j > 0 && __any_sync(0xffffffffu, need)) {    // rescale this warp's 16 columns of O (warp-uniform: tcgen05.ld / st are collectives)
          tc_fence_after();
          uint32_t o[16];
          tmem_ld16(lane_addr + MH_O_COL + (uint32_t)(half * 16), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(lane_addr + MH_O_COL + (uint32_t)(half * 16), o);
        }
      }
      // P as the TENSOR-MEMORY operand of P.V: this thread's 32 keys are 16 columns of fp16 pairs in the hi image and 16 in the lo image
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int c = 0; c < 16; c++) split_f16x2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1]), ph[c], pl[c]);
      tmem_st16(lane_addr + MH_P_COL + (uint32_t)(half * 16), ph);
      tmem_st16(lane_addr + MH_P_COL + 32u + (uint32_t)(half * 16), pl);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (j == nblk - 1) {
        // head finished: O / l (row sums of the two key halves are exchanged first); this warp stores 16 of the head's 32 columns
        s_red[half][r] = l_run;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float inv = 1.f / (l_run + s_red[half ^ 1][r]);        // l = 0 (all keys masked): 0 * inf = NaN, the reference's behaviour
        asm volatile("bar.sync 2, 256;" ::: "memory");
        mbar_wait(pv_done, (uint32_t)g & 1u);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(lane_addr + MH_O_COL + (uint32_t)(half * 16), o);
        tmem_ld_wait();
        tc_fence_before();
        float* ob = p.out + ((long long)b * p.L + (long long)qt * MH_BQ + r) * p.ldo + (hp * 2 + x) * 32 + half * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(ob + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, __uint_as_float(o[i + 2]) * inv,
                                                           __uint_as_float(o[i + 3]) * inv);
        m_ref = -CUDART_INF_F; l_run = 0.f;
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issue (one elected lane runs the whole loop) ===============================
    if (elect_one_sync()) {
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(MH_BKV >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);             // K-major A and B, N = 64
      const uint32_t idesc_o = (1u << 4) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // B (= V) MN-major, N = 32
      // V block, MN-major SWIZZLE_128B: one key per 128-byte row, 8-key groups 1 KB apart (SBO); a single 64-value group (LBO unused)
      const uint64_t v_desc_bits = ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t s_tmem = tmem_base + MH_S_COL, o_tmem = tmem_base + MH_O_COL;
      auto issue_scores = [&](int g) {
        const int st = g & 1, x = g >= nblk ? 1 : 0;
        mbar_wait(k_full(st), (uint32_t)(g >> 1) & 1u);
        if (g > 0) mbar_wait(s_free, (uint32_t)(g - 1) & 1u);
        tc_fence_after();
        const uint32_t k_hi = sbase + MH_OFF_K + (uint32_t)st * 16384u, k_lo = k_hi + 8192u;
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const uint64_t ko = (uint64_t)(x * 4 + ks * 2);                     // head x: bytes 64x.. of the row; 32 bytes per k-step
          const uint64_t dq = make_desc(sbase + MH_OFF_QH) + ko, dkh = make_desc(k_hi) + ko, dkl = make_desc(k_lo) + ko;
          mh_mma_ts(s_tmem, tmem_base + MH_QL_COL + (uint32_t)(x * 16 + ks * 8), dkh, idesc_s, ks != 0);   // q_lo * k_hi
          tc_mma_f16(s_tmem, dq, dkl, idesc_s, 1u);                                                      // q_hi * k_lo
          tc_mma_f16(s_tmem, dq, dkh, idesc_s, 1u);                                                      // q_hi * k_hi
        }
        tc_commit(s_full);
        tc_commit(k_empty(st));
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_scores(0);
      for (int g = 0; g < ntot; g++) {
        if (g + 1 < ntot) issue_scores(g + 1);
        const int st = g & 1, x = g >= nblk ? 1 : 0, j = g - x * nblk;
        mbar_wait(v_full(st), (uint32_t)(g >> 1) & 1u);
        mbar_wait(p_full, (uint32_t)g & 1u);
        tc_fence_after();
        const uint32_t v_hi = sbase + MH_OFF_V + (uint32_t)st * 16384u, v_lo = v_hi + 8192u;
        const uint32_t p_hi = tmem_base + MH_P_COL, p_lo = p_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {               // 16 keys per k-step: 8 TMEM columns of P, two 8-key row groups of V
          const uint64_t dvh = v_desc_bits | (uint64_t)(((v_hi + (uint32_t)ks * 2048u + (uint32_t)x * 64u) & 0x3FFFFu) >> 4);
          const uint64_t dvl = v_desc_bits | (uint64_t)(((v_lo + (uint32_t)ks * 2048u + (uint32_t)x * 64u) & 0x3FFFFu) >> 4);
          mh_mma_ts(o_tmem, p_lo + (uint32_t)(ks * 8), dvh, idesc_o, (j | ks) != 0);
          mh_mma_ts(o_tmem, p_hi + (uint32_t)(ks * 8), dvl, idesc_o, 1u);
          mh_mma_ts(o_tmem, p_hi + (uint32_t)(ks * 8), dvh, idesc_o, 1u);
        }
        tc_commit(pv_done);
        tc_commit(v_empty(st));
      }
    }
    __syncwarp();
  } else {
    // =============================== loader ===============================
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384u);
      bulk_g2s(sbase + MH_OFF_QH, qh + ((long long)b * nqt + qt) * (4LL * MH_BQ * 64) + (long long)hp * (MH_BQ * 64), 16384u, q_full);
      for (int g = 0; g < ntot; g++) {
        const int st = g & 1, j = g >= nblk ? g - nblk : g; const uint32_t ph = ((uint32_t)(g >> 1) & 1u) ^ 1u;
        const long long tile = ((long long)kvb * nblk + j) * (4LL * MH_BKV * 64) + (long long)hp * (MH_BKV * 64);
        mbar_wait(k_empty(st), ph);
        mbar_expect_tx(k_full(st), 16384u);
        bulk_g2s(sbase + MH_OFF_K + (uint32_t)st * 16384u, kh + tile, 8192u, k_full(st));
        bulk_g2s(sbase + MH_OFF_K + (uint32_t)st * 16384u + 8192u, kl + tile, 8192u, k_full(st));
        mbar_wait(v_empty(st), ph);
        mbar_expect_tx(v_full(st), 16384u);
        bulk_g2s(sbase + MH_OFF_V + (uint32_t)st * 16384u, vh + tile, 8192u, v_full(st));
        bulk_g2s(sbase + MH_OFF_V + (uint32_t)st * 16384u + 8192u, vl + tile, 8192u, v_full(st));
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// the four operand images [k hi | k lo | v hi | v lo] (fp16, tile layout) of k, v (S, 256) shared by every frame: 4 * S * 256 halfs
extern "C" int sma_attn_split_kv(const float* k, int ldk, const float* v, int ldv, int S, void* images, sma_stream_t stream) {
  if (!k || !v || !images || S <= 0) return SMA_ERR_BAD_ARG;
  if ((S % MH_BKV) || ((ldk | ldv) & 3)) return SMA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(images)) & 15) return SMA_ERR_BAD_ARG;
  // (no q: with B * L = 0 query rows the split kernel's q images are empty and the k images start at `images`)
  return sma_attn_split_launch(nullptr, 0, k, ldk, v, ldv, 0, 0, 0, 1, 0, S, 1.f, images, as_stream(stream));
}

extern "C" int64_t sma_mha_e256_workspace_bytes(int B, int kvB, int L, int S) {
  if (B <= 0 || kvB <= 0 || L <= 0 || S <= 0) return 0;
  return 2LL * MH_E * 2 * ((long long)B * L + 2LL * kvB * S);      // fp16 hi + lo images of q, k, v
}

extern "C" int sma_mha_e256_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int64_t q_bstride, int64_t kv_bstride,
                                int B, int L, int S, float scale, const uint8_t* key_mask, void* workspace, float* out, int ldo, int presplit, sma_stream_t stream) {
  // presplit: 0 = q, k, v are fp32 tensors, split here; 1 = the q images were written into `workspace` by the projection's epilogue (sma_conv_desc.split_ws),
  // k and v are split here; 2 = q, k and v images are all there already (q / k / v pointers are then unused)
  // 3 = q images in `workspace`, and `k` points at the operand images [k hi | k lo | v hi | v lo] of k, v shared by all frames (kv_bstride = 0), written once by
  // sma_attn_split_kv (frame-invariant codebook projections: nothing is split per call)
  if (presplit < 0 || presplit > 3 || !out || !workspace || B <= 0 || L <= 0 || S <= 0) return SMA_ERR_BAD_ARG;
  if ((presplit == 0 && !q) || (presplit < 2 && (!k || !v)) || (presplit == 3 && (!k || kv_bstride != 0))) return SMA_ERR_BAD_ARG;
  const uint16_t* kv_images = presplit == 3 ? reinterpret_cast<const uint16_t*>(k) : nullptr;
  if (presplit == 3) v = k;
  if (presplit) { if (!q) q = out; if (!k) { k = out; v = out; } }         // (only their alignment is looked at below)
  if ((L % MH_BQ) || (S % MH_BKV) || B > 65535) return SMA_ERR_UNSUPPORTED;
  if (((ldq | ldk | ldv | ldo) & 3) || ((q_bstride | kv_bstride) & 3)) return SMA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(workspace)) & 15)
    return SMA_ERR_BAD_ARG;
  if (key_mask && ((reinterpret_cast<uintptr_t>(key_mask) & 15) || (S & 15))) return SMA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int kvB = kv_bstride ? B : 1;
  if (presplit < 2) {
    int rs = sma_attn_split_launch(presplit ? nullptr : q, ldq, k, ldk, v, ldv, q_bstride, kv_bstride, B, kvB, L, S, scale * 1.4426950408889634f, workspace, st);
    if (rs != SMA_OK) return rs;
  }
  static SmaDevOnce once;
  if (int rc = sma_opt_in_smem(once, attn_mh_kernel, (int)MH_SMEM + 1024)) return rc;
  MhP p; p.ws = reinterpret_cast<const uint16_t*>(workspace); p.kv = kv_images; p.mask = key_mask; p.out = out; p.ldo = ldo; p.B = B; p.kvB = kvB; p.L = L; p.S = S;
  attn_mh_kernel<<<dim3(L / MH_BQ, 4, B), MH_THREADS, MH_SMEM + 1024, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
