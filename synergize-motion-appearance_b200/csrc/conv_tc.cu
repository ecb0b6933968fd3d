// tcgen05 implicit-GEMM convolution, NHWC fp32 in / fp32 out, fp32-faithful through a 3xTF32 split
// (a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation in TMEM), or single-pass TF32.
//
// GEMM view (same as conv_simt.cu): M = B*Ho*Wo output pixels, N = Cout, K = kh*kw*Cin, k = tap*Cin + c.
// One CTA = one 128 x NT output tile (NT <= 256, a multiple of 16); 6 warps:
//   warps 0-3  A producers, then the epilogue.  Per 32-channel K-chunk they gather the 128 pixel rows of the tap
//              from global memory with coalesced 128-bit loads (8 lanes cover the 128 contiguous bytes of one pixel),
//              apply the fused prologue (GroupNorm scale/shift + swish, nearest-x2 upsample, zero padding after the
//              normalisation), split every value into tf32 hi / fp32 residual lo, and store both into shared memory in
//              the canonical K-major SWIZZLE_128B UMMA layout (row = 128 B, 8-row atoms of 1 KB, 16-byte chunk index
//              XOR row%8).  After the last chunk they read the accumulator back with tcgen05.ld (warp w owns TMEM
//              lanes 32w..32w+31 = tile rows) and run the fused epilogue (bias, activation, residual, depth-to-space).
//   warp 4     TMEM allocation + the single MMA-issuing thread: per chunk 4 k-steps (K=8 each) x 3 products.
//   warp 5     weight loader: one cp.async.bulk (TMA bulk copy, no tensor map) per chunk brings the pre-swizzled
//              hi|lo weight image of this (N-tile, chunk) - packed once per weight load by sma_pack_conv_weight_tc -
//              straight into its stage slot, completing on the stage's mbarrier.
// Stages form an mbarrier ring: full[s] (128 producer arrivals + 1 expect_tx arrival + the bulk-copy bytes) and
// empty[s] (tcgen05.commit of the MMAs that read the slot).
#include "sma_common.cuh"
#include "tc_common.cuh"
#include <cuda.h>          // CUtensorMap (types only: cuTensorMapEncodeTiled is resolved at run time through cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128;          // tile rows (UMMA M, cta_group::1)
constexpr int KC = 32;           // fp32 per K-chunk = one 128-byte swizzle row
constexpr int A_BYTES = BM * KC * 4;   // 16 KB per A image (hi or lo)
constexpr int MAX_STAGES = 4;
constexpr int SMEM_DYN_MAX = 232448 - 3072;  // 227 KB opt-in limit minus room for static shared memory (barriers, bias, scales)
constexpr int SMEM_LIMIT = SMEM_DYN_MAX - 1024;  // minus 1 KB alignment slack

struct TcP {
  const float* x; const float* wtc; const float* bias; const float* pre_scale; const float* pre_shift; const float* res; float* y;
  long long in_bs, out_bs, res_bs;
  int Hi, Wi, Cin, in_ld, Cout, kh, kw, stride, pad_t, pad_l, up, pre_act;
  int Ho, Wo, out_ld, act, res_ld, d2s;
  int M, HoWo, nchunks, cpt /* chunks per tap */, NT, stages, passes, tmem_cols;
};

__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1 KB alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_bytes = p.NT * KC * 4;                       // one B image (hi or lo)
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;
  const int m0 = blockIdx.x * BM, nt = blockIdx.y;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * MAX_STAGES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 128 + 1); mbar_init(empty_bar(s), 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4) {
    // =============================== A producers ===============================
    const int cq = lane & 7;                     // 16-byte chunk of the 128-byte row
    const int rsub = lane >> 3;                  // 4 rows per warp instruction
    int iy0[8], ix0[8]; long long boff[8]; int bidx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int m = m0 + warp * 32 + j * 4 + rsub;
      bool ok = m < p.M;
      int mm = ok ? m : 0;
      int b = mm / p.HoWo; int r = mm - b * p.HoWo; int oy = r / p.Wo; int ox = r - oy * p.Wo;
      bidx[j] = b; boff[j] = (long long)b * p.in_bs;
      iy0[j] = ok ? oy * p.stride - p.pad_t : -0x40000000;   // invalid rows never pass the range test
      ix0[j] = ox * p.stride - p.pad_l;
    }
    const int Hv = p.Hi << p.up, Wv = p.Wi << p.up;
    int cc = 0, ky = 0, kx = 0;
    for (int kc = 0; kc < p.nchunks; kc++) {
      const int s = kc % p.stages; const uint32_t ph = (kc / p.stages) & 1;
      const int c = cc * KC + cq * 4;
      float4 v[8]; bool ok[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int iy = iy0[j] + ky, ix = ix0[j] + kx;
        ok[j] = (unsigned)iy < (unsigned)Hv && (unsigned)ix < (unsigned)Wv;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[j]) v[j] = __ldg(reinterpret_cast<const float4*>(p.x + boff[j] + ((long long)(iy >> p.up) * p.Wi + (ix >> p.up)) * p.in_ld + c));
      }
      if (p.pre_scale) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (ok[j]) {
            float4 sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)bidx[j] * p.Cin + c));
            float4 sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)bidx[j] * p.Cin + c));
            v[j].x = sma_act(fmaf(v[j].x, sc.x, sh.x), p.pre_act); v[j].y = sma_act(fmaf(v[j].y, sc.y, sh.y), p.pre_act);
            v[j].z = sma_act(fmaf(v[j].z, sc.z, sh.z), p.pre_act); v[j].w = sma_act(fmaf(v[j].w, sc.w, sh.w), p.pre_act);
          }
        }
      }
      mbar_wait(empty_bar(s), ph ^ 1u);          // slot free (its previous MMAs have completed)
      const uint32_t a_hi = sbase + (uint32_t)s * stage_bytes, a_lo = a_hi + A_BYTES;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int r = warp * 32 + j * 4 + rsub;
        const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((cq ^ (r & 7)) << 4);
        float hx = tf32_rna(v[j].x), hy = tf32_rna(v[j].y), hz = tf32_rna(v[j].z), hw = tf32_rna(v[j].w);
        sts128(a_hi + off, hx, hy, hz, hw);
        if (p.passes == 3) sts128(a_lo + off, v[j].x - hx, v[j].y - hy, v[j].z - hz, v[j].w - hw);
      }
      fence_async_smem();                        // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(full_bar(s));
      if (++kx == p.kw) { kx = 0; if (++ky == p.kh) { ky = 0; ++cc; } }     // chunk order: channel chunk outer, tap inner
    }
    // =============================== epilogue ===============================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const bool mok = m < p.M;
    const int mm = mok ? m : 0;
    const int b = mm / p.HoWo; const int r = mm - b * p.HoWo; const int oy = r / p.Wo; const int ox = r - oy * p.Wo;
    const int nbase = nt * p.NT;
    const int nend = min(p.Cout, nbase + p.NT);          // columns of THIS tile (a narrowed N tile need not be a multiple of the 32-column read)
    const bool vec_ok = (p.out_ld & 3) == 0 && (p.out_bs & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (!p.res || ((p.res_ld & 3) == 0 && (p.res_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    const int Cq = p.d2s > 1 ? p.Cout / (p.d2s * p.d2s) : p.Cout;     // channels of the depth-to-space output
    for (int n0 = 0; n0 < p.NT && nbase + n0 < nend; n0 += 32) {
      uint32_t a[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]),
            "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]),
            "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]),
            "=r"(a[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (mok) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const int n = nbase + n0 + q * 4;
          if (n >= nend) break;
          float o[4];
#pragma unroll
          for (int t = 0; t < 4; t++) {
            float acc = __uint_as_float(a[q * 4 + t]);
            if (p.bias && n + t < nend) acc += __ldg(p.bias + n + t);
            o[t] = sma_act(acc, p.act);
          }
          // destination of column n for this pixel
          float* dst; int cval = n;
          if (p.d2s > 1) {
            int qd = n / Cq; cval = n - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
            long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
            dst = p.y + (long long)b * p.out_bs + pix * p.out_ld + cval;
          } else {
            dst = p.y + (long long)b * p.out_bs + (long long)r * p.out_ld + n;
          }
          if (vec_ok && n + 4 <= nend && (Cq & 3) == 0) {
            if (p.res) {
              float4 rr = __ldg(reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + n));
              o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
            }
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int t = 0; t < 4; t++) {
              if (n + t >= nend) break;
              float val = o[t];
              if (p.res) val += __ldg(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + n + t);
              if (p.d2s > 1) {
                int nn = n + t; int qd = nn / Cq; int c2 = nn - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
                long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
                p.y[(long long)b * p.out_bs + pix * p.out_ld + c2] = val;
              } else {
                dst[t] = val;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =============================== MMA issuer (whole warp converged, one elected lane issues) ===============================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kc = 0; kc < p.nchunks; kc++) {
      const int s = kc % p.stages; const uint32_t ph = (kc / p.stages) & 1;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_hi = sbase + (uint32_t)s * stage_bytes;
      const uint64_t dah = make_desc(a_hi), dal = make_desc(a_hi + A_BYTES);
      const uint64_t dbh = make_desc(a_hi + 2 * A_BYTES), dbl = make_desc(a_hi + 2 * A_BYTES + b_bytes);
      if (elect_one_sync()) {
#pragma unroll
        for (int k4 = 0; k4 < KC / 8; k4++) {
          const uint64_t ko = (uint64_t)(k4 * 2);          // 8 tf32 = 32 bytes = 2 x 16-byte units along K inside the swizzle row
          if (p.passes == 3) {
            tc_mma_tf32(tmem_base, dal + ko, dbh + ko, idesc, (kc | k4) != 0);
            tc_mma_tf32(tmem_base, dah + ko, dbl + ko, idesc, 1u);
            tc_mma_tf32(tmem_base, dah + ko, dbh + ko, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_base, dah + ko, dbh + ko, idesc, (kc | k4) != 0);
          }
        }
        tc_commit(empty_bar(s));                           // frees the slot when these MMAs retire
        if (kc == p.nchunks - 1) tc_commit(accum_bar);     // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  } else {
    // =============================== weight loader ===============================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(p.passes == 3 ? 2 * b_bytes : b_bytes);
      const char* src = reinterpret_cast<const char*>(p.wtc) + (long long)nt * p.nchunks * (2 * b_bytes);
      for (int kc = 0; kc < p.nchunks; kc++) {
        const int s = kc % p.stages; const uint32_t ph = (kc / p.stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), bytes);
        bulk_g2s(sbase + (uint32_t)s * stage_bytes + 2 * A_BYTES, src + (long long)kc * (2 * b_bytes), bytes, full_bar(s));
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}



// The q | k | v projection written straight into the attention kernels' operand images (attn256.cu / attn_mh.cu): the 8 consecutive columns a thread holds per
// 256-bit store are exactly one 16-byte piece of a 128-byte tile row - the granularity of the split pass this replaces (which re-read 4.6 GB of
// fp32 q / k / v per step to write the same bytes).  q is pre-scaled by scale * log2 e; hi = rn(x), lo = rn(x - hi); SWIZZLE_128B rows; q_lo piece-major.
__device__ __forceinline__ void store_attn_split(uint16_t* ws, long long img_q, long long img_k, float qscale, int L, const float (&o)[8], int b, int r, int n) {
  const int which = n >> 8, e = n & 255, g = e >> 3, c4 = g >> 3, ch = g & 7;
  const float s = which == 0 ? qscale : 1.f;
  uint4 hi, lo;
  split_f16x2(o[0] * s, o[1] * s, hi.x, lo.x); split_f16x2(o[2] * s, o[3] * s, hi.y, lo.y);
  split_f16x2(o[4] * s, o[5] * s, hi.z, lo.z); split_f16x2(o[6] * s, o[7] * s, hi.w, lo.w);
  const int RB = which == 0 ? 128 : 64;                                // rows per tile: 128 queries / 64 keys
  const int rb = r / RB, rr = r - rb * RB;
  const long long tile = ((long long)b * (L / RB) + rb) * (4LL * RB * 64);
  const long long off = tile + (long long)c4 * RB * 64 + (((uint32_t)rr * 128u + (uint32_t)((ch ^ (rr & 7)) << 4)) >> 1);
  if (which == 0) {
    *reinterpret_cast<uint4*>(ws + off) = hi;
    *reinterpret_cast<uint4*>(ws + img_q + tile + ((long long)(c4 * 8 + ch) * 128 + rr) * 8) = lo;       // piece-major for the TMEM loaders
  } else {
    uint16_t* kh = ws + 2 * img_q + (which == 2 ? 2 * img_k : 0);
    *reinterpret_cast<uint4*>(kh + off) = hi; *reinterpret_cast<uint4*>(kh + img_k + off) = lo;
  }
}

// GroupNorm statistics of the tile a convolution has just computed (fused producer-side: the standalone statistics pass re-reads every normalised
// tensor from HBM, 15 GB per 64-frame step).  The 8 consecutive channels o[0..7] of this lane's pixel row (zero for rows outside the image) are
// reduced over the warp's 32 rows as 4 channel PAIRS x {sum, sum of squares}: a transposing butterfly - three rounds that halve the number of
// values while exchanging them across the lane bits 4, 3, 2, then two plain rounds - 9 shuffles for the 8 values; lane 4j ends up with value j
// (j < 4: sum of pair j, j >= 4: sum of squares of pair j - 4) and stores it.  `dst` = partial row of this (frame, 32-row chunk) + 2 * first pair.
__device__ __forceinline__ void gn_pairs_reduce_store(const float (&o)[8], bool mok, int lane, float* dst) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float x0 = mok ? o[2 * j] : 0.f, x1 = mok ? o[2 * j + 1] : 0.f;
    v[j] = x0 + x1; v[4 + j] = fmaf(x0, x0, x1 * x1);
  }
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float keep = h16 ? v[4 + i] : v[i], send = h16 ? v[i] : v[4 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const float keep = h8 ? v[2 + i] : v[i], send = h8 ? v[i] : v[2 + i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float keep = h4 ? v[1] : v[0], send = h4 ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  if ((lane & 3) == 0) { const int j = lane >> 2; dst[(j & 3) * 2 + (j >> 2)] = v[0]; }
}

// =================================================================================================================
// v2: persistent, halo-resident kernel for stride-1 convolutions (1x1, 3x3, 7x7; optional fused nearest-x2 upsample).
//
// One CTA per SM loops over output tiles of 8 (x) by 16 (y) pixels (1x1: 128 consecutive pixels).  For every 32-channel
// chunk the producers load the (8+kw-1) x (16+kh-1) input halo ONCE, apply the prologue and the tf32 split ONCE per
// element, and store one 128-byte row per halo pixel (SWIZZLE_128B keyed on the absolute shared-memory address).
// The im2col for tap (ky,kx) is then done by the tensor core itself: the A descriptor starts at halo row
// ky*HALO_W+kx and uses the halo row pitch HALO_W*128 B as the 8-row-atom stride (an 8-pixel output row segment is 8
// consecutive halo rows), so the kh*kw taps re-read shared memory instead of L2 and cost no producer work.
// Weights stream through their own ring, one (channel chunk, tap) image per bulk copy.  Two TMEM accumulators let the
// 4 epilogue warps drain tile t while the MMA thread works on tile t+1.
//   warps 0-3 epilogue | warp 4 MMA issue + TMEM alloc | warp 5 weight loader | warps 6-13 halo producers
// =================================================================================================================
constexpr int V2_PROD_WARPS = 8;                       // halo producers: enough loads in flight to cover HBM latency
constexpr int V2_THREADS = 192 + 32 * V2_PROD_WARPS;       // register-staged mode, 14 warps: 4 epilogue | MMA | weight loader | 8 halo producers
// staged-input (TMA) mode, 16 warps: 8 epilogue (two per tensor-memory lane quarter, alternate 32-column blocks) | MMA | copy issuer (one thread drives
// the weight ring and the TMA staging ring) | 6 converters
constexpr int V2_CONV_WARPS = 6;
constexpr int V2_THREADS_TMA = 512;
constexpr int V2_PGROUPS = 2;                          // producer groups working on alternate channel chunks (2x the latency budget each)
constexpr int V2_PROWS = 4 * V2_PROD_WARPS / V2_PGROUPS;   // halo rows per pass of one group (8 lanes per 128-byte row)
#ifndef SMA_V2_UNROLL
#define SMA_V2_UNROLL 12             // (macro: tools/build_variants.sh A/B-tests it; 6 -> 12 = -11 % on the halo-latency-bound 128->64 3x3 @256^2)
#endif
constexpr int V2_UNROLL = SMA_V2_UNROLL;                          // loads in flight per producer thread
constexpr int MAX_SA = 3, MAX_SB = 8, MAX_NS = 8;


struct Tc2P {
  const float* x; const float* wtc; const float* bias; const float* pre_scale; const float* pre_shift; const float* res; float* y;
  long long in_bs, out_bs, res_bs;
  int Hi, Wi, Cin, in_ld, Cout, kh, kw, pad_t, pad_l, up, pre_act;
  int Ho, Wo, out_ld, act, res_ld, d2s;
  const float* aux; long long aux_bs; int aux_ld; float sft_w;   // SFT epilogue (aux != nullptr): y = res + sft_w * (res * aux + v), v = act(acc + bias)
  const float* wscale;        // F16 only: per-output-channel power-of-two factor that undoes the weight pre-scaling
  int res_pipe;               // host: the launch qualifies for the RES = 1 instantiation (residual, no activation, 256-bit eligible)
  float acc_corr;             // F16 only: 1 + 1.6e-8 * (adds into the main accumulator): undoes the mean truncation bias of the tensor core's fp32 accumulation
  int HoWo, cpt, taps, NT, ntiles_n, passes, tmem_cols;
  int fuse;                   // 3-pass, NT <= 128: the hi and lo weight images (adjacent in the ring) are read as ONE B tile of 2*NT rows, so
                              // a_hi*[b_hi | b_lo] is one MMA; its lo half lands in accumulator columns [NT, 2NT) and is added in the epilogue
  int acc_cols;               // TMEM columns per accumulator (NT or 2*NT)
  uint16_t* split_ws; long long split_img_q, split_img_k; float split_qscale;   // != nullptr: the output is written as the attention kernels' fp16 hi / lo tile images
                              // (q | k | v of E = 256: columns [0,256) q, [256,512) k, [512,768) v; attn256.cu attn256_split_kernel's layout) instead of fp32 rows
  float* gnp; int gn_chunks;  // fused GroupNorm partial sums: [frame][chunk = 4 * tile-in-image + lane quarter][Cout / 2 pairs][sum, sum of squares]; nullptr: off
  int dbg;                    // timing experiments (results invalid): bit 2 = epilogue skips its residual loads and stores
  int flat, tiles_x, tiles_per_img, total_tiles;
  int halo_w, HP, a_img_bytes, a_stage_bytes, b_img_bytes, SA, SB;
  // TMA-staged input (TMA = 1): warp 14 brings the fp32 halo of every (tile, 64-channel chunk) into a ring of NS shared-memory slots with
  // cp.async.bulk.tensor (a 4-D map over (C, W, H, B): zero fill outside the image = the convolution's padding; flat 1x1 layers: a 3-D map over
  // (C, pixels, B)), `parts` boxes of `slot_rows` halo rows each; the 8 producer warps convert slot by slot.  No thread waits on global memory.
  int NS, parts, slot_rows, slot_bytes, rs /* image rows per box */, tma_b_fixed /* batch stride 0: always coordinate 0 */;
  int patch;                  // > 0: patch-embedding conv (kernel = stride = patch, no padding) as a strided gather: a 5-D map over (C, px, X, py, B*Y); chunk = (channel chunk, tap)
  alignas(64) CUtensorMap tmap;
  alignas(64) CUtensorMap tmap2;   // second input tensor (cpt1 < cpt): channel chunks [cpt1, cpt) are read from it (two convolutions over different tensors summed in one accumulator)
  int cpt1;
  int pre_ld;                 // row pitch of pre_scale / pre_shift (= channels of the first input tensor)
  int x2_center;              // the second tensor's chunks contribute through the CENTRE tap only (a 1x1 conv over x2 added to a k x k conv over x: a ResBlock's
                              // conv2 + its 1x1 skip conv); the weight image then holds one blob per such chunk; nblob = weight blobs per (tile, N tile)
  int nblob;
};

// ACT: epilogue activation; PRE: -1 no prologue, else the prologue activation applied after scale/shift (compile-time so that the
// unrolled epilogue / producer bodies stay small: the three warp roles share one instruction cache)
// F16: operands are split into two fp16 halves (hi = rn(x), lo = rn(x - hi); same 11-bit significand as tf32) and multiplied with
// kind::f16 (K = 16 per instruction: twice the MACs per tensor-core cycle and per shared-memory byte of kind::tf32); a K-chunk
// (one 128-byte swizzle row) is then 64 channels.
// RES: 1 = the 256-bit epilogue with the residual prefetched one column block ahead (own instantiation: its register budget must not touch the others)
template <int ACT, int PRE, int F16, int RES, int TMA>
__global__ void __launch_bounds__(TMA ? V2_THREADS_TMA : V2_THREADS, 1) conv_tc2_kernel(const __grid_constant__ Tc2P p) {
  constexpr int KCH = F16 ? 64 : 32;
  // warp roles: [0, EPW) epilogue, W_MMA, W_MMA + 1 weight loader, then (TMA) the tensor-load issuer and the converters, else the 8 halo producers
  constexpr int EPW = TMA ? 8 : 4, W_MMA = EPW, W_LOAD = EPW + 1, W_CONV0 = EPW + 2;                       // channels per K-chunk (128 bytes of operand)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_SA + 2 * MAX_SB + 4 + 2 * MAX_NS];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256];              // bias of the current N tile (epilogue warps only)
  __shared__ __align__(16) float s_scale[F16 ? 256 : 4];   // F16: un-scaling factors of the current N tile
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (MAX_SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * MAX_SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * MAX_SA + MAX_SB + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + 2 + s); };
  auto s_full = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + 4 + s); };
  auto s_empty = [&](int s) { return bar0 + 8u * (2 * MAX_SA + 2 * MAX_SB + 4 + MAX_NS + s); };
  const uint32_t a_ring = sbase, b_ring = sbase + (uint32_t)p.SA * p.a_stage_bytes;
  const int b_stage_bytes = 2 * p.b_img_bytes;
  const uint32_t stg_ring = b_ring + (uint32_t)p.SB * b_stage_bytes;      // (TMA) fp32 staging slots; 1 KB aligned like everything before it

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.SA; s++) { mbar_init(a_full(s), TMA ? 32 * V2_CONV_WARPS : 32 * V2_PROD_WARPS / V2_PGROUPS); mbar_init(a_empty(s), 1); }
    if (TMA) for (int s = 0; s < p.NS; s++) { mbar_init(s_full(s), 1); mbar_init(s_empty(s), V2_CONV_WARPS); }
    for (int s = 0; s < p.SB; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; s++) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 32 * EPW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // tile -> (image b, first output row / column or first flat pixel, N tile)
  auto tile_in_img = [&](int tile) { const int mt = tile / p.ntiles_n; return mt - (mt / p.tiles_per_img) * p.tiles_per_img; };
  auto decode = [&](int tile, int& b, int& ty0, int& tx0, int& nt) {
    nt = tile % p.ntiles_n; int mt = tile / p.ntiles_n;
    b = mt / p.tiles_per_img; int t = mt - b * p.tiles_per_img;
    if (p.flat) { ty0 = t * BM; tx0 = 0; } else { int tyi = t / p.tiles_x; ty0 = tyi * 16; tx0 = (t - tyi * p.tiles_x) * 8; }
  };

  if (warp < EPW) {
    // =============================== epilogue ===============================
    // warp w reads tensor-memory lanes 32 (w % 4) .. +31 (= tile rows); with 8 warps, warp w and w + 4 share a lane quarter and take alternate
    // 32-column blocks: the per-tile instruction stream of an epilogue thread halves (it was as long as the MMAs of a tile)
    const int wq = warp & 3, half = warp >> 2;
    const int m = wq * 32 + lane;
    const bool vec_ok = (p.out_ld & 3) == 0 && (p.out_bs & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (!p.res || ((p.res_ld & 3) == 0 && (p.res_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    const int Cq = p.d2s > 1 ? p.Cout / (p.d2s * p.d2s) : p.Cout;
    const bool v8_ok = (p.d2s <= 1 || (Cq & 7) == 0) && (p.Cout & 7) == 0 && (p.out_ld & 7) == 0 && (p.out_bs & 7) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 31) == 0) &&
                       (!p.res || ((p.res_ld & 7) == 0 && (p.res_bs & 7) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 31) == 0));   // (the SFT epilogue is only dispatched here when this holds)
    int tcount = 0, cur_nt = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      int b, ty0, tx0, nt; decode(tile, b, ty0, tx0, nt);
      const int ab = tcount & 1; const uint32_t aph = (tcount >> 1) & 1;
      if (nt != cur_nt) {                           // (re)stage the bias slice of this N tile; named barrier 1 = the epilogue warps
        asm volatile("bar.sync 1, %0;" ::"r"(32 * EPW) : "memory");
        for (int i = threadIdx.x; i < p.NT; i += 32 * EPW) {
          int n = nt * p.NT + i; s_bias[i] = (p.bias && n < p.Cout) ? __ldg(p.bias + n) : 0.f;
          if (F16) s_scale[i] = __ldg(p.wscale + n) * p.acc_corr;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * EPW) : "memory");      // (every epilogue warp: a count of 128 let four of the eight warps of the staged-input instantiations run ahead of the other four's writes)
        cur_nt = nt;
      }
      int oy, ox, r; bool mok;
      if (p.flat) { r = ty0 + m; mok = r < p.HoWo; oy = r / p.Wo; ox = r - oy * p.Wo; }
      else { oy = ty0 + (m >> 3); ox = tx0 + (m & 7); mok = oy < p.Ho && ox < p.Wo; r = oy * p.Wo + ox; }
      if constexpr (RES == 1) {
        const int nbase = nt * p.NT;
        // 256-bit path for layers with a residual.  Loaded next to its use, every 8 columns of the residual exposed a full DRAM latency (32 per
        // tile of a 256-column linear: the short-K transformer linears ran 4x below their MMA / HBM time).  Here (1) the residual row of the
        // NEXT tile of this CTA is prefetched into L2 one tile ahead (two 128-byte lines per thread and 64 columns), and (2) the four loads of a
        // 32-column block are issued together, before the accumulator block is read, so that one (L2) latency is exposed per block.
        float* yrow = p.y + (long long)b * p.out_bs + (long long)r * p.out_ld;
        const float* rrow = p.res + (long long)b * p.res_bs + (long long)(mok ? r : 0) * p.res_ld;       // (rows outside the image re-read row 0: unconditional loads)
        const float* arow = p.aux ? p.aux + (long long)b * p.aux_bs + (long long)(mok ? r : 0) * p.aux_ld : nullptr;
        float rc[32], rn[32];
  #define SMA_LD_V8(A, O, ptr)                                                                                                          \
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                                  \
                 : "=f"(A[(O) + 0]), "=f"(A[(O) + 1]), "=f"(A[(O) + 2]), "=f"(A[(O) + 3]), "=f"(A[(O) + 4]), "=f"(A[(O) + 5]), "=f"(A[(O) + 6]), "=f"(A[(O) + 7]) \
                 : "l"(ptr))
        // (columns past Cout re-read column block 0: unconditional loads keep the arrays in registers)
  #define SMA_LD_BLOCK(A, rowp, n0_)                                                                         \
    {                                                                                                        \
      _Pragma("unroll") for (int q8_ = 0; q8_ < 4; q8_++) {                                                  \
        const int n_ = nbase + (n0_) + q8_ * 8;                                                              \
        SMA_LD_V8(A, q8_ * 8, (rowp) + (n_ < p.Cout ? n_ : nbase));                                          \
      }                                                                                                      \
    }
        {                                                // L2 prefetch of the next tile's residual (and SFT scale) row of this thread
          const int tile2 = tile + (int)gridDim.x;
          if (tile2 < p.total_tiles && half == 0) {
            int b2, ty2, tx2, nt2; decode(tile2, b2, ty2, tx2, nt2);
            int r2; bool ok2;
            if (p.flat) { r2 = ty2 + m; ok2 = r2 < p.HoWo; }
            else { const int oy2 = ty2 + (m >> 3), ox2 = tx2 + (m & 7); ok2 = oy2 < p.Ho && ox2 < p.Wo; r2 = oy2 * p.Wo + ox2; }
            if (ok2) {
              const float* q = p.res + (long long)b2 * p.res_bs + (long long)r2 * p.res_ld + nt2 * p.NT;
              const int nb2 = min(p.NT, p.Cout - nt2 * p.NT) * 4;
              for (int o = 0; o < nb2; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(q) + o));
              if (p.aux) {
                const float* qa = p.aux + (long long)b2 * p.aux_bs + (long long)r2 * p.aux_ld + nt2 * p.NT;
                for (int o = 0; o < nb2; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(qa) + o));
              }
            }
          }
        }
        mbar_wait(acc_full(ab), aph);
        tc_fence_after();
        if (half * 32 >= p.NT || nbase + half * 32 >= p.Cout) {      // (8 warps) no column block for this half: release the accumulator right away
          tc_fence_before();
          mbar_arrive(acc_empty(ab));
        }
        for (int n0 = half * 32; n0 < p.NT && nbase + n0 < p.Cout; n0 += 8 * EPW) {
          uint32_t a[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * p.acc_cols + n0);
          tmem_ld32(taddr, a);
          if (p.fuse) {                                  // the lo products were accumulated NT columns further right (two halves: registers)
  #pragma unroll
            for (int hh = 0; hh < 2; hh++) {
              uint32_t a2[16];
              tmem_ld16(taddr + (uint32_t)(p.NT + hh * 16), a2);
              tmem_ld_wait();
  #pragma unroll
              for (int i = 0; i < 16; i++) a[hh * 16 + i] = __float_as_uint(__uint_as_float(a[hh * 16 + i]) + __uint_as_float(a2[i]));
            }
          } else {
            tmem_ld_wait();
          }
          SMA_LD_BLOCK(rc, rrow, n0)                      // (after the lo half of the accumulator is folded in: register budget)
          if (arow) SMA_LD_BLOCK(rn, arow, n0)
          const bool last = n0 + 8 * EPW >= p.NT || nbase + n0 + 8 * EPW >= p.Cout;      // this warp's last column block
          if (last) {                                    // last column block: the accumulator may be overwritten while we store
            tc_fence_before();
            mbar_arrive(acc_empty(ab));
          }
          if ((mok || p.gnp) && !(p.dbg & 4)) {        // (RES = 1 is only dispatched for 256-bit-eligible launches; with fused statistics every lane runs the shuffles)
  #pragma unroll
            for (int q8 = 0; q8 < 4; q8++) {
              const int n = nbase + n0 + q8 * 8;
              if (n >= p.Cout) break;
              float o[8];
              const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[n0 + q8 * 8]), b1 = *reinterpret_cast<const float4*>(&s_bias[n0 + q8 * 8 + 4]);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              if (F16) {
                const float4 s0 = *reinterpret_cast<const float4*>(&s_scale[n0 + q8 * 8]), s1 = *reinterpret_cast<const float4*>(&s_scale[n0 + q8 * 8 + 4]);
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] = sma_act(fmaf(__uint_as_float(a[q8 * 8 + t]), ss[t], bb[t]), ACT);
              } else {
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] = sma_act(__uint_as_float(a[q8 * 8 + t]) + bb[t], ACT);
              }
              if (arow) {         // Fuse_sft_block tail (appmotioncodebook_arch.py:50-51): dec + w * (dec * scale + shift), this conv = shift
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] = rc[q8 * 8 + t] + p.sft_w * (rc[q8 * 8 + t] * rn[q8 * 8 + t] + o[t]);
              } else {
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] += rc[q8 * 8 + t];
              }
              if (p.gnp) gn_pairs_reduce_store(o, mok, lane, p.gnp + (((long long)b * p.gn_chunks + tile_in_img(tile) * 4 + wq) * (p.Cout >> 1) + (n >> 1)) * 2);
              float* dst = yrow + n;
              if (p.d2s > 1) {      // depth-to-space (un-patchify): the 8 columns lie inside one sub-pixel's channel block (Cq % 8 == 0)
                const int qd = n / Cq, cval = n - qd * Cq, p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
                const long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
                dst = p.y + (long long)b * p.out_bs + pix * p.out_ld + cval;
              }
              if (mok) {
                if (p.split_ws) store_attn_split(p.split_ws, p.split_img_q, p.split_img_k, p.split_qscale, p.HoWo, o, b, r, n);
                else
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]),
                             "f"(o[6]), "f"(o[7])
                             : "memory");
              }
            }
          }
        }

      } else {
        const int nbase = nt * p.NT;
        mbar_wait(acc_full(ab), aph);
        tc_fence_after();
        if (half * 32 >= p.NT || nbase + half * 32 >= p.Cout) {      // (8 warps) no column block for this half: release the accumulator right away
          tc_fence_before();
          mbar_arrive(acc_empty(ab));
        }
        for (int n0 = half * 32; n0 < p.NT && nbase + n0 < p.Cout; n0 += 8 * EPW) {
          uint32_t a[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * p.acc_cols + n0);
          tmem_ld32(taddr, a);
          if (p.fuse) {                                  // a_hi * b_lo was accumulated NT columns further right
            uint32_t a2[32];
            tmem_ld32(taddr + (uint32_t)p.NT, a2);
            tmem_ld_wait();
  #pragma unroll
            for (int i = 0; i < 32; i++) a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(a2[i]));
          } else {
            tmem_ld_wait();
          }
          if (n0 + 8 * EPW >= p.NT || nbase + n0 + 8 * EPW >= p.Cout) {     // this warp's last column block: the accumulator may be overwritten while we store
            tc_fence_before();
            mbar_arrive(acc_empty(ab));
          }
          if ((mok || p.gnp) && !(p.dbg & 4) && v8_ok) {
            // 256-bit residual loads / stores: every lane moves whole 32-byte sectors (with 128-bit accesses a warp-level instruction touches
            // half of 32 different sectors, and the epilogue - not the MMAs - bounds the low-K layers)
            const int rs_ = mok ? r : 0;                      // (fused statistics: rows outside the image run along with row 0 and store nothing)
            float* yrow = p.y + (long long)b * p.out_bs + (long long)rs_ * p.out_ld;
            const float* rrow = p.res ? p.res + (long long)b * p.res_bs + (long long)rs_ * p.res_ld : nullptr;
            const float* arow = p.aux ? p.aux + (long long)b * p.aux_bs + (long long)rs_ * p.aux_ld : nullptr;
  #pragma unroll
            for (int q8 = 0; q8 < 4; q8++) {
              const int n = nbase + n0 + q8 * 8;
              if (n >= p.Cout) break;
              float o[8];
              const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[n0 + q8 * 8]), b1 = *reinterpret_cast<const float4*>(&s_bias[n0 + q8 * 8 + 4]);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              if (F16) {
                const float4 s0 = *reinterpret_cast<const float4*>(&s_scale[n0 + q8 * 8]), s1 = *reinterpret_cast<const float4*>(&s_scale[n0 + q8 * 8 + 4]);
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] = sma_act(fmaf(__uint_as_float(a[q8 * 8 + t]), ss[t], bb[t]), ACT);
              } else {
  #pragma unroll
                for (int t = 0; t < 8; t++) o[t] = sma_act(__uint_as_float(a[q8 * 8 + t]) + bb[t], ACT);
              }
              if (rrow) {
                float r0, r1, r2, r3, r4, r5, r6, r7;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3), "=f"(r4), "=f"(r5), "=f"(r6), "=f"(r7)
                             : "l"(rrow + n));
                if (arow) {       // Fuse_sft_block tail (appmotioncodebook_arch.py:50-51): dec + w * (dec * scale + shift), this conv = shift
                  float a0, a1, a2, a3, a4, a5, a6, a7;
                  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7)
                               : "l"(arow + n));
                  o[0] = r0 + p.sft_w * (r0 * a0 + o[0]); o[1] = r1 + p.sft_w * (r1 * a1 + o[1]); o[2] = r2 + p.sft_w * (r2 * a2 + o[2]);
                  o[3] = r3 + p.sft_w * (r3 * a3 + o[3]); o[4] = r4 + p.sft_w * (r4 * a4 + o[4]); o[5] = r5 + p.sft_w * (r5 * a5 + o[5]);
                  o[6] = r6 + p.sft_w * (r6 * a6 + o[6]); o[7] = r7 + p.sft_w * (r7 * a7 + o[7]);
                } else {
                  o[0] += r0; o[1] += r1; o[2] += r2; o[3] += r3; o[4] += r4; o[5] += r5; o[6] += r6; o[7] += r7;
                }
              }
              if (p.gnp) gn_pairs_reduce_store(o, mok, lane, p.gnp + (((long long)b * p.gn_chunks + tile_in_img(tile) * 4 + wq) * (p.Cout >> 1) + (n >> 1)) * 2);
              float* dst = yrow + n;
              if (p.d2s > 1) {      // depth-to-space (un-patchify): the 8 columns lie inside one sub-pixel's channel block (Cq % 8 == 0)
                const int qd = n / Cq, cval = n - qd * Cq, p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
                const long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
                dst = p.y + (long long)b * p.out_bs + pix * p.out_ld + cval;
              }
              if (mok) {
                if (p.split_ws) store_attn_split(p.split_ws, p.split_img_q, p.split_img_k, p.split_qscale, p.HoWo, o, b, r, n);
                else
                asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]),
                             "f"(o[6]), "f"(o[7])
                             : "memory");
              }
            }
          } else if (mok && !(p.dbg & 4)) {
  #pragma unroll
            for (int q = 0; q < 8; q++) {
              const int n = nbase + n0 + q * 4;
              if (n >= p.Cout) break;
              float o[4];
  #pragma unroll
              for (int t = 0; t < 4; t++) {
                if (F16) o[t] = sma_act(fmaf(__uint_as_float(a[q * 4 + t]), s_scale[n0 + q * 4 + t], s_bias[n0 + q * 4 + t]), ACT);
                else o[t] = sma_act(__uint_as_float(a[q * 4 + t]) + s_bias[n0 + q * 4 + t], ACT);
              }
              if (vec_ok && n + 4 <= p.Cout && (Cq & 3) == 0) {
                float* dst;
                if (p.d2s > 1) {
                  int qd = n / Cq; int cval = n - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
                  long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
                  dst = p.y + (long long)b * p.out_bs + pix * p.out_ld + cval;
                } else {
                  dst = p.y + (long long)b * p.out_bs + (long long)r * p.out_ld + n;
                }
                if (p.res) {
                  float4 rr = __ldg(reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + n));
                  o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
                }
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
  #pragma unroll
                for (int t = 0; t < 4; t++) {
                  if (n + t >= p.Cout) break;
                  float val = o[t];
                  if (p.res) val += __ldg(p.res + (long long)b * p.res_bs + (long long)r * p.res_ld + n + t);
                  if (p.d2s > 1) {
                    int nn = n + t; int qd = nn / Cq; int c2 = nn - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
                    long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
                    p.y[(long long)b * p.out_bs + pix * p.out_ld + c2] = val;
                  } else {
                    p.y[(long long)b * p.out_bs + (long long)r * p.out_ld + n + t] = val;
                  }
                }
              }
            }
          }
        }

      }
    }
  } else if (warp == W_MMA) {
    // =============================== MMA issuer (whole warp converged, one elected lane issues) ===============================
    // fp32 accumulate; A/B format 2 = tf32 (kind::tf32) or 0 = fp16 (kind::f16); K-major A and B
    const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * p.NT) >> 3) << 17);      // same, N = 2 NT
    // A: 8-row atoms are 8 consecutive halo rows; next atom = next output row = halo row pitch
    const uint64_t a_desc_hi_bits = (1ull << 16) | ((uint64_t)((p.halo_w * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
    // ONE elected lane runs the whole tile loop (the elect region encloses the loops) with wrap-around ring counters instead of
    // divisions: measured with the tensor-memory-operand kernel, the per-tap scalar overhead of the old form (warp-sync + elect +
    // reconvergence + `%` and `/` by run-time ring sizes) let the tensor pipe drain between taps.
    if (elect_one_sync()) {
      const uint32_t row_step = (uint32_t)(p.halo_w * 128) >> 4;          // one halo row down, in 16-byte units
      const bool three = p.passes == 3;
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait(acc_empty(ab), ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * (uint32_t)p.acc_cols;
        uint32_t acc = 0u;                                   // the first MMA of a tile overwrites the accumulator
        for (int cc = 0; cc < p.cpt; cc++) {
          mbar_wait(a_full(sa), pha);
          const uint32_t a_hi = a_ring + sa * (uint32_t)p.a_stage_bytes;
          const uint64_t dah0 = a_desc_hi_bits | (uint64_t)((a_hi & 0x3FFFFu) >> 4);
          const uint64_t dal0 = a_desc_hi_bits | (uint64_t)(((a_hi + (uint32_t)p.a_img_bytes) & 0x3FFFFu) >> 4);
          uint32_t roff = 0;
          const bool centre_only = p.x2_center && cc >= p.cpt1;
          for (int ky = 0; ky < p.kh; ky++, roff += row_step) {
            for (int kx = 0; kx < p.kw; kx++) {
              if (centre_only && (ky != p.pad_t || kx != p.pad_l)) continue;
              mbar_wait(b_full(sb), phb);
              tc_fence_after();
              const uint64_t toff = (uint64_t)(roff + (uint32_t)kx * 8u);
              const uint64_t dah = dah0 + toff, dal = dal0 + toff;
              const uint32_t b_hi = b_ring + sb * (uint32_t)b_stage_bytes;
              const uint64_t dbh = make_desc(b_hi), dbl = make_desc(b_hi + p.b_img_bytes);
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) {                   // 4 k-steps of 32 bytes (8 tf32 / 16 fp16) per 128-byte row
                const uint64_t ko = (uint64_t)(k4 * 2);
                if (p.fuse) {
                  tc_mma<F16>(d_tmem, dah + ko, dbh + ko, idesc2, acc);          // a_hi * [b_hi | b_lo]  (N = 2 NT)
                  // a_lo * b_hi goes to the LO half of the accumulator (columns [NT, 2 NT), added in the epilogue with round-to-nearest):
                  // the tensor core adds into its fp32 accumulator with truncation (measured: a bias of -1.6e-8 per add, relative), so
                  // the main accumulator should see as few adds as possible - one per k-step instead of two
                  if (three) tc_mma<F16>(d_tmem + (uint32_t)p.NT, dal + ko, dbh + ko, idesc, 1u);     // (2-product mode: the activations' lo halves are dropped)
                } else if (three) {
                  tc_mma<F16>(d_tmem, dal + ko, dbh + ko, idesc, acc);
                  tc_mma<F16>(d_tmem, dah + ko, dbl + ko, idesc, 1u);
                  tc_mma<F16>(d_tmem, dah + ko, dbh + ko, idesc, 1u);
                } else if (p.passes == 2) {
                  tc_mma<F16>(d_tmem, dah + ko, dbl + ko, idesc, acc);
                  tc_mma<F16>(d_tmem, dah + ko, dbh + ko, idesc, 1u);
                } else {
                  tc_mma<F16>(d_tmem, dah + ko, dbh + ko, idesc, acc);
                }
                acc = 1u;
              }
              tc_commit(b_empty(sb));
              if (++sb == (uint32_t)p.SB) { sb = 0; phb ^= 1u; }
            }
          }
          tc_commit(a_empty(sa));
          if (++sa == (uint32_t)p.SA) { sa = 0; pha ^= 1u; }
        }
        tc_commit(acc_full(ab));
      }
    }
    __syncwarp();
  } else if (warp == W_LOAD) {
    // =============================== weight loader ===============================
    if (TMA && lane == 0) {
      // staged-input mode: ONE thread feeds both rings.  It polls the two "slot free" barriers without blocking on either (a blocked weight slot
      // must not hold back a free staging slot and vice versa), so that a 16th warp is not needed for the tensor loads (16 warps = 128 registers).
      const uint32_t bytes = (uint32_t)(p.passes >= 2 ? b_stage_bytes : p.b_img_bytes);
      const uint64_t tm1 = reinterpret_cast<uint64_t>(&p.tmap), tm2 = reinterpret_cast<uint64_t>(&p.tmap2);
      const int nblob = p.nblob;
      auto test = [&](uint32_t bar, uint32_t parity) -> bool {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        return ok != 0;
      };
      // weight cursor
      int wt = blockIdx.x, wj = 0; uint32_t wsb = 0, wph = 0;
      // staging cursor
      int st_ = blockIdx.x, scc = 0, spart = 0; uint32_t ss = 0, phs = 0;
      int sb_, sty0, stx0, snt; decode(st_ < p.total_tiles ? st_ : 0, sb_, sty0, stx0, snt);
      bool wdone = wt >= p.total_tiles, sdone = st_ >= p.total_tiles || (p.dbg & 2);
      while (!wdone || !sdone) {
        bool progressed = false;
        if (!sdone && test(s_empty(ss), phs ^ 1u)) {
          mbar_expect_tx(s_full(ss), (uint32_t)p.slot_bytes);
          const uint32_t dst = stg_ring + ss * (uint32_t)p.slot_bytes;
          const int bc = p.tma_b_fixed ? 0 : sb_;
          const bool second = scc >= p.cpt1;                 // chunk of the second input tensor
          const uint64_t tm = second ? tm2 : tm1;
          const int ch0 = (second ? scc - p.cpt1 : scc) * 64;
          if (p.patch) {            // chunk scc = (channel chunk, tap (py, px)); box = 32 consecutive tokens (one token row or a part of it)
            const int pp = p.patch * p.patch, cch = scc / pp, tap = scc - cch * pp, py = tap / p.patch, px = tap - py * p.patch;
            const int tok = sty0 + spart * 32, tyy = tok / p.Wo, txx = tok - tyy * p.Wo;
            asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(dst), "l"(tm), "r"(cch * 64), "r"(px), "r"(txx), "r"(py), "r"(sb_ * p.Ho + tyy), "r"(s_full(ss)) : "memory");
          } else if (p.flat) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(tm), "r"(ch0), "r"(sty0 + spart * p.slot_rows), "r"(bc), "r"(s_full(ss)) : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst), "l"(tm), "r"(ch0), "r"(stx0 - p.pad_l), "r"(sty0 - p.pad_t + spart * p.rs), "r"(bc), "r"(s_full(ss)) : "memory");
          }
          if (++ss == (uint32_t)p.NS) { ss = 0; phs ^= 1u; }
          if (++spart == p.parts) {
            spart = 0;
            if (++scc == p.cpt) {
              scc = 0; st_ += gridDim.x;
              if (st_ >= p.total_tiles) sdone = true; else decode(st_, sb_, sty0, stx0, snt);
            }
          }
          progressed = true;
        }
        if (!wdone && test(b_empty(wsb), wph ^ 1u)) {
          if (p.dbg & 1) mbar_arrive(b_full(wsb));                        // timing experiment: no weight traffic
          else {
            const char* src = reinterpret_cast<const char*>(p.wtc) + ((long long)(wt % p.ntiles_n) * nblob + wj) * b_stage_bytes;
            mbar_expect_tx(b_full(wsb), bytes);
            bulk_g2s(b_ring + wsb * (uint32_t)b_stage_bytes, src, bytes, b_full(wsb));
          }
          if (++wsb == (uint32_t)p.SB) { wsb = 0; wph ^= 1u; }
          if (++wj == nblob) { wj = 0; wt += gridDim.x; if (wt >= p.total_tiles) wdone = true; }
          progressed = true;
        }
        if (!progressed) __nanosleep(32);
      }
    } else if (!TMA && lane == 0) {
      const uint32_t bytes = (uint32_t)(p.passes >= 2 ? b_stage_bytes : p.b_img_bytes);
      int jt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.ntiles_n;
        const char* src = reinterpret_cast<const char*>(p.wtc) + (long long)nt * p.nblob * b_stage_bytes;
        const int nblob = p.nblob;
        for (int j = 0; j < nblob; j++, jt++) {
          const int sb = jt % p.SB; const uint32_t phb = (jt / p.SB) & 1;
          mbar_wait(b_empty(sb), phb ^ 1u);
          if (p.dbg & 1) { mbar_arrive(b_full(sb)); continue; }          // timing experiment: no weight traffic
          mbar_expect_tx(b_full(sb), bytes);
          bulk_g2s(b_ring + (uint32_t)sb * b_stage_bytes, src + (long long)j * b_stage_bytes, bytes, b_full(sb));
        }
      }
    }
    __syncwarp();
  } else if (TMA) {
    // =============================== halo converters (staged-input mode): shared fp32 slot -> prologue -> fp16 hi / lo operand rows ===============================
    // (A/B'd: moving the two sleeping roles - weight loader, TMA issuer - onto the MMA warp's scheduler so that no converter competes with the MMA
    //  thread for issue slots is 1-2 % SLOWER: spreading the converters evenly over the four schedulers matters more)
    const int ptid = threadIdx.x - 32 * W_CONV0; const int cq = ptid & 7; const int prow = ptid >> 3;      // 24 halo rows per pass of the 6 warps
    const int swap = cq >> 2;                                   // lanes 4-7 read their two 16-byte halves in the opposite order: conflict-free LDS.128
    const int Hv = p.Hi, Wv = p.Wi;
    uint32_t ss = 0, phs = 0; int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int b, ty0, tx0, nt; decode(tile, b, ty0, tx0, nt);
      for (int cc = 0; cc < p.cpt; cc++, it++) {
        const int sa = it % p.SA; const uint32_t pha = (it / p.SA) & 1;
        const uint32_t a_hi = a_ring + (uint32_t)sa * p.a_stage_bytes, a_lo = a_hi + p.a_img_bytes;
        const int c = cc * KCH + cq * 8;
        const bool pro = PRE >= 0 && cc < p.cpt1;                // the prologue belongs to the first input tensor (pre_scale / pre_shift are (B, Cin1))
        float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sc1 = sc0, sh1 = sh0;
        if (pro) {
          sc0 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.pre_ld + c));
          sc1 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.pre_ld + c + 4));
          sh0 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.pre_ld + c));
          sh1 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.pre_ld + c + 4));
        }
        mbar_wait(a_empty(sa), pha ^ 1u);
        if (p.dbg & 2) { mbar_arrive(a_full(sa)); continue; }          // timing experiment: no halo traffic, no conversion
        int hp = prow;                                           // this thread's rows run on across the boxes of the chunk: prow, prow + 24, ...
        for (int part = 0; part < p.parts; part++) {
          mbar_wait(s_full(ss), phs);
          const uint32_t slot = stg_ring + ss * (uint32_t)p.slot_bytes;
          const int hp_end = min((part + 1) * p.slot_rows, p.HP);
          for (; hp < hp_end; hp += 4 * V2_CONV_WARPS) {
            const int lr = hp - part * p.slot_rows;
            const uint32_t src = slot + (uint32_t)lr * 256u + (uint32_t)cq * 32u;
            float4 u0, u1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(u0.x), "=f"(u0.y), "=f"(u0.z), "=f"(u0.w) : "r"(src + (uint32_t)swap * 16u));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(u1.x), "=f"(u1.y), "=f"(u1.z), "=f"(u1.w) : "r"(src + 16u - (uint32_t)swap * 16u));
            float4 t0 = swap ? u1 : u0, t1 = swap ? u0 : u1;
            if (pro) {
              bool ok;                                         // zero padding applies AFTER the normalisation: pixels outside the image stay zero
              if (p.flat) ok = ty0 + hp < p.HoWo;
              else { const int hy = hp / p.halo_w, hx = hp - hy * p.halo_w; ok = (unsigned)(ty0 + hy - p.pad_t) < (unsigned)Hv && (unsigned)(tx0 + hx - p.pad_l) < (unsigned)Wv; }
              if (ok) {
                t0.x = pre_act_fast(fmaf(t0.x, sc0.x, sh0.x), PRE); t0.y = pre_act_fast(fmaf(t0.y, sc0.y, sh0.y), PRE);
                t0.z = pre_act_fast(fmaf(t0.z, sc0.z, sh0.z), PRE); t0.w = pre_act_fast(fmaf(t0.w, sc0.w, sh0.w), PRE);
                t1.x = pre_act_fast(fmaf(t1.x, sc1.x, sh1.x), PRE); t1.y = pre_act_fast(fmaf(t1.y, sc1.y, sh1.y), PRE);
                t1.z = pre_act_fast(fmaf(t1.z, sc1.z, sh1.z), PRE); t1.w = pre_act_fast(fmaf(t1.w, sc1.w, sh1.w), PRE);
              }
            }
            const uint32_t off = (uint32_t)hp * 128u + (uint32_t)((cq ^ (hp & 7)) << 4);
            uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
            split_f16x2(t0.x, t0.y, h0, l0); split_f16x2(t0.z, t0.w, h1, l1);
            split_f16x2(t1.x, t1.y, h2, l2); split_f16x2(t1.z, t1.w, h3, l3);
            sts128u(a_hi + off, h0, h1, h2, h3);
            if (p.passes == 3) sts128u(a_lo + off, l0, l1, l2, l3);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(ss));              // this warp is done reading the slot
          if (++ss == (uint32_t)p.NS) { ss = 0; phs ^= 1u; }
        }
        fence_async_smem();
        mbar_arrive(a_full(sa));
      }
    }
  } else {
    // =============================== halo producers ===============================
    const int pgroup = (threadIdx.x - 192) / (32 * V2_PROD_WARPS / V2_PGROUPS);
    const int ptid = (threadIdx.x - 192) % (32 * V2_PROD_WARPS / V2_PGROUPS); const int cq = ptid & 7; const int prow = ptid >> 3;   // V2_PROWS halo rows per pass
    const int Hv = p.Hi << p.up, Wv = p.Wi << p.up;
    const int npass = (p.HP + V2_PROWS - 1) / V2_PROWS;
    const bool in32 = (p.in_ld & 7) == 0 && (p.in_bs & 7) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 31) == 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int b, ty0, tx0, nt; decode(tile, b, ty0, tx0, nt);
      const float* xb = p.x + (long long)b * p.in_bs;
      // halo row hp of this tile -> source pixel index (ok = inside the image; padding rows stay zero)
      auto halo_pixel = [&](int hp, bool& ok) -> long long {
        if (p.flat) { int r = ty0 + hp; ok = r < p.HoWo; return r; }
        int hy = hp / p.halo_w; int hx = hp - hy * p.halo_w;
        int iy = ty0 + hy - p.pad_t, ix = tx0 + hx - p.pad_l;
        ok = (unsigned)iy < (unsigned)Hv && (unsigned)ix < (unsigned)Wv;
        return (long long)(iy >> p.up) * p.Wi + (ix >> p.up);
      };
      for (int cc = 0; cc < p.cpt; cc++, it++) {
        if ((it % V2_PGROUPS) != pgroup) continue;      // the other group's chunk
        const int sa = it % p.SA; const uint32_t pha = (it / p.SA) & 1;
        const uint32_t a_hi = a_ring + (uint32_t)sa * p.a_stage_bytes, a_lo = a_hi + p.a_img_bytes;
        bool waited = false;
        if (p.dbg & 2) {                                   // timing experiment: no halo loads / stores
          mbar_wait(a_empty(sa), pha ^ 1u);
          mbar_arrive(a_full(sa));
          continue;
        }
        if (!F16) {
          const int c = cc * KCH + cq * 4;
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (PRE >= 0) {
            sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.Cin + c));
            sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.Cin + c));
          }
          for (int pass0 = 0; pass0 < npass; pass0 += V2_UNROLL) {
            float4 v[V2_UNROLL]; bool ok[V2_UNROLL];
#pragma unroll
            for (int u = 0; u < V2_UNROLL; u++) {
              const int hp = (pass0 + u) * V2_PROWS + prow;
              ok[u] = false; v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (hp < p.HP) {
                const long long pix = halo_pixel(hp, ok[u]);
                if (ok[u]) v[u] = __ldg(reinterpret_cast<const float4*>(xb + pix * p.in_ld + c));
              }
            }
            if (!waited) { mbar_wait(a_empty(sa), pha ^ 1u); waited = true; }     // first loads are in flight while we wait
#pragma unroll
            for (int u = 0; u < V2_UNROLL; u++) {
              const int hp = (pass0 + u) * V2_PROWS + prow;
              if (hp >= p.HP) continue;
              float4 t = v[u];
              if (PRE >= 0 && ok[u]) {
                t.x = pre_act_fast(fmaf(t.x, sc.x, sh.x), PRE); t.y = pre_act_fast(fmaf(t.y, sc.y, sh.y), PRE);
                t.z = pre_act_fast(fmaf(t.z, sc.z, sh.z), PRE); t.w = pre_act_fast(fmaf(t.w, sc.w, sh.w), PRE);
              }
              const uint32_t off = (uint32_t)hp * 128u + (uint32_t)((cq ^ (hp & 7)) << 4);
              float hx_ = tf32_rna(t.x), hy_ = tf32_rna(t.y), hz_ = tf32_rna(t.z), hw_ = tf32_rna(t.w);
              sts128(a_hi + off, hx_, hy_, hz_, hw_);
              if (p.passes == 3) sts128(a_lo + off, t.x - hx_, t.y - hy_, t.z - hz_, t.w - hw_);
            }
          }
        } else {
          constexpr int UN = V2_UNROLL / 2;                 // two float4 (8 channels = one 16-byte fp16 unit) per halo row and lane
          const int c = cc * KCH + cq * 8;
          float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sc1 = sc0, sh1 = sh0;
          if (PRE >= 0) {
            sc0 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.Cin + c));
            sc1 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.Cin + c + 4));
            sh0 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.Cin + c));
            sh1 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.Cin + c + 4));
          }
          for (int pass0 = 0; pass0 < npass; pass0 += UN) {
            float4 v0[UN], v1[UN]; bool ok[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) {
              const int hp = (pass0 + u) * V2_PROWS + prow;
              ok[u] = false; v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
              if (hp < p.HP) {
                const long long pix = halo_pixel(hp, ok[u]);
                if (ok[u]) {
                  const float* src = xb + pix * p.in_ld + c;
                  if (in32) {                     // one 256-bit load = one whole 32-byte sector per lane
                    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v0[u].x), "=f"(v0[u].y), "=f"(v0[u].z), "=f"(v0[u].w),
                                 "=f"(v1[u].x), "=f"(v1[u].y), "=f"(v1[u].z), "=f"(v1[u].w) : "l"(src));
                  } else {
                    v0[u] = __ldg(reinterpret_cast<const float4*>(src)); v1[u] = __ldg(reinterpret_cast<const float4*>(src) + 1);
                  }
                }
              }
            }
            if (!waited) { mbar_wait(a_empty(sa), pha ^ 1u); waited = true; }
#pragma unroll
            for (int u = 0; u < UN; u++) {
              const int hp = (pass0 + u) * V2_PROWS + prow;
              if (hp >= p.HP) continue;
              float4 t0 = v0[u], t1 = v1[u];
              if (PRE >= 0 && ok[u]) {
                t0.x = pre_act_fast(fmaf(t0.x, sc0.x, sh0.x), PRE); t0.y = pre_act_fast(fmaf(t0.y, sc0.y, sh0.y), PRE);
                t0.z = pre_act_fast(fmaf(t0.z, sc0.z, sh0.z), PRE); t0.w = pre_act_fast(fmaf(t0.w, sc0.w, sh0.w), PRE);
                t1.x = pre_act_fast(fmaf(t1.x, sc1.x, sh1.x), PRE); t1.y = pre_act_fast(fmaf(t1.y, sc1.y, sh1.y), PRE);
                t1.z = pre_act_fast(fmaf(t1.z, sc1.z, sh1.z), PRE); t1.w = pre_act_fast(fmaf(t1.w, sc1.w, sh1.w), PRE);
              }
              const uint32_t off = (uint32_t)hp * 128u + (uint32_t)((cq ^ (hp & 7)) << 4);
              uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
              split_f16x2(t0.x, t0.y, h0, l0); split_f16x2(t0.z, t0.w, h1, l1);
              split_f16x2(t1.x, t1.y, h2, l2); split_f16x2(t1.z, t1.w, h3, l3);
              sts128u(a_hi + off, h0, h1, h2, h3);
              if (p.passes == 3) sts128u(a_lo + off, l0, l1, l2, l3);
            }
          }
        }
        fence_async_smem();
        mbar_arrive(a_full(sa));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// packed [K][ldw] fp32 weight -> per (N-tile, chunk) image: [hi: NT rows x 128 B, SWIZZLE_128B][lo: same]
__global__ void pack_tc_kernel(const float* __restrict__ wp, int ldw, int Cin, int taps, int Cout, int NT, int ntiles, float* __restrict__ out) {
  const int nchunks = taps * Cin / KC;
  const long long total = (long long)ntiles * nchunks * NT * KC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(i % KC); long long t = i / KC; int n = (int)(t % NT); t /= NT; int kc = (int)(t % nchunks); int nt = (int)(t / nchunks);
    const int cc = kc / taps, tap = kc - cc * taps;              // image chunk order: channel chunk outer, tap inner
    int col = nt * NT + n, k = tap * Cin + cc * KC + kk;         // row of the [K][ldw] packed weight (k = tap*Cin + c)
    float w = col < Cout ? wp[(long long)k * ldw + col] : 0.f;
    uint32_t hb; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(w));
    float hi = __uint_as_float(hb), lo = w - hi;
    long long blob = ((long long)nt * nchunks + kc) * (2LL * NT * KC);
    int phys = (n >> 3) * 256 + (n & 7) * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));   // float index inside the image
    out[blob + phys] = hi;
    out[blob + (long long)NT * KC + phys] = lo;
  }
}

// per output column n: factor 2^e with max_k |w[k][n]| * 2^-e in [0.5, 1)  (exact power-of-two pre-scaling keeps the fp16 lo halves
// of small weights out of the subnormal range); columns >= Cout get 1
__global__ void f16_colscale_kernel(const float* __restrict__ wp, int ldw, int K, int Cout, int ncols, float* __restrict__ inv_scale) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= ncols) return;
  float mx = 0.f;
  if (n < Cout) for (int k = 0; k < K; k++) mx = fmaxf(mx, fabsf(wp[(long long)k * ldw + n]));
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &e); e = max(-100, min(100, e)); }
  inv_scale[n] = ldexpf(1.f, e);
}

// packed [K][ldw] fp32 weight -> per (N-tile, 64-channel chunk, tap) fp16 images [hi: NT rows x 128 B, SWIZZLE_128B][lo: same] of w * 2^-e
__global__ void pack_tc16_kernel(const float* __restrict__ wp, int ldw, int Cin, int taps, int Cout, int NT, int ntiles,
                                 const float* __restrict__ inv_scale, uint16_t* __restrict__ out) {
  constexpr int KH = 64;
  const int nchunks = taps * Cin / KH;
  const long long total = (long long)ntiles * nchunks * NT * KH;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(i % KH); long long t = i / KH; int n = (int)(t % NT); t /= NT; int kc = (int)(t % nchunks); int nt = (int)(t / nchunks);
    const int cc = kc / taps, tap = kc - cc * taps;
    int col = nt * NT + n, k = tap * Cin + cc * KH + kk;
    float w = col < Cout ? wp[(long long)k * ldw + col] / inv_scale[col] : 0.f;      // exact: the divisor is a power of two
    __half hi = __float2half_rn(w); __half lo = __float2half_rn(w - __half2float(hi));
    long long blob = ((long long)nt * nchunks + kc) * (2LL * NT * KH);
    int phys = (n >> 3) * 512 + (n & 7) * 64 + ((((kk >> 3) ^ (n & 7)) << 3) | (kk & 7));    // fp16 index inside the image
    out[blob + phys] = __half_as_ushort(hi);
    out[blob + (long long)NT * KH + phys] = __half_as_ushort(lo);
  }
}

int tc_ntile(int Cout) { return Cout <= 256 ? ((Cout + 15) & ~15) : 256; }

}  // namespace

extern "C" int64_t sma_conv_weight_tc_floats(int Cout, int Cin, int kh, int kw, int nt) {
  if (Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || (Cin % KC) || nt < 0 || nt > 256 || (nt & 15)) return 0;
  int NT = nt ? nt : tc_ntile(Cout); int ntiles = (Cout + NT - 1) / NT;
  return (int64_t)ntiles * (kh * kw * Cin / KC) * 2 * NT * KC;
}

extern "C" int sma_pack_conv_weight_tc(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, int nt, float* w_tc, sma_stream_t stream) {
  if (!w_packed || !w_tc || Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || ldw < Cout || nt < 0 || nt > 256 || (nt & 15)) return SMA_ERR_BAD_ARG;
  if (Cin % KC) return SMA_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(w_tc) & 15) return SMA_ERR_BAD_ARG;
  int NT = nt ? nt : tc_ntile(Cout); int ntiles = (Cout + NT - 1) / NT; int K = kh * kw * Cin;
  long long total = (long long)ntiles * (K / KC) * NT * KC;
  int blocks = (int)((total + 255) / 256); if (blocks > 8192) blocks = 8192;
  pack_tc_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_packed, ldw, Cin, kh * kw, Cout, NT, ntiles, w_tc);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

// fp16 image: [ntiles*NT floats of un-scaling factors][per (N tile, 64-channel chunk, tap): hi image | lo image]; size in floats
extern "C" int64_t sma_conv_weight_tc16_floats(int Cout, int Cin, int kh, int kw) {
  if (Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || (Cin % 64)) return 0;
  int NT = tc_ntile(Cout); int ntiles = (Cout + NT - 1) / NT;
  return (int64_t)ntiles * NT + (int64_t)ntiles * (kh * kw * Cin / 64) * NT * 64;      // 2 images x NT x 64 halfs = NT*64 floats per chunk
}

extern "C" int sma_pack_conv_weight_tc16(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, float* w_tc16, sma_stream_t stream) {
  if (!w_packed || !w_tc16 || Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || ldw < Cout) return SMA_ERR_BAD_ARG;
  if (Cin % 64) return SMA_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(w_tc16) & 15) return SMA_ERR_BAD_ARG;
  int NT = tc_ntile(Cout); int ntiles = (Cout + NT - 1) / NT; int K = kh * kw * Cin; int ncols = ntiles * NT;
  f16_colscale_kernel<<<(ncols + 127) / 128, 128, 0, as_stream(stream)>>>(w_packed, ldw, K, Cout, ncols, w_tc16);
  SMA_LAUNCH_CHECK();
  long long total = (long long)ntiles * (K / 64) * NT * 64;
  int blocks = (int)((total + 255) / 256); if (blocks > 8192) blocks = 8192;
  pack_tc16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_packed, ldw, Cin, kh * kw, Cout, NT, ntiles, w_tc16,
                                                          reinterpret_cast<uint16_t*>(w_tc16 + ncols));
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}



// cuTensorMapEncodeTiled through the runtime's driver entry point query (the library does not link libcuda: it must load on a GPU-less box)
typedef CUresult (*SmaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static SmaEncodeTiledFn sma_tmap_encoder() {
  static std::atomic<void*> cached{nullptr};
  void* f = cached.load(std::memory_order_acquire);
  if (f) return reinterpret_cast<SmaEncodeTiledFn>(f);
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) return nullptr;
  cached.store(f, std::memory_order_release);
  return reinterpret_cast<SmaEncodeTiledFn>(f);
}

template <int ACT, int PRE, int F16, int RES, int TMA>
static int launch_tc2_inst2(const Tc2P& p, int grid, int smem, cudaStream_t st) {
  static SmaDevOnce once;             // per instantiation and per device
  if (int rc = sma_opt_in_smem(once, conv_tc2_kernel<ACT, PRE, F16, RES, TMA>, SMEM_DYN_MAX)) return rc;
  conv_tc2_kernel<ACT, PRE, F16, RES, TMA><<<grid, TMA ? V2_THREADS_TMA : V2_THREADS, smem, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
template <int ACT, int PRE, int F16, int RES = 0>
static int launch_tc2_inst(const Tc2P& p, int grid, int smem, cudaStream_t st) {
  if (F16 && p.NS > 0) return launch_tc2_inst2<ACT, PRE, F16, RES, F16 ? 1 : 0>(p, grid, smem, st);
  return launch_tc2_inst2<ACT, PRE, F16, RES, 0>(p, grid, smem, st);
}
template <int ACT, int PRE, int F16>
static int launch_tc2_res(const Tc2P& p, int grid, int smem, cudaStream_t st) {
  if constexpr (F16 != 0) {
    if constexpr (ACT == SMA_ACT_NONE && (PRE == -1 || PRE == SMA_ACT_SWISH)) {
      if (p.res_pipe) return launch_tc2_inst<ACT, PRE, 1, 1>(p, grid, smem, st);       // no room for it (256-column 3x3 tiles): register epilogue with the residual pipeline
    }
  }
  return launch_tc2_inst<ACT, PRE, F16, 0>(p, grid, smem, st);
}
template <int ACT, int F16>
static int launch_tc2_pre(int pre, const Tc2P& p, int grid, int smem, cudaStream_t st) {
  switch (pre) {
    case -1: return launch_tc2_res<ACT, -1, F16>(p, grid, smem, st);
    case SMA_ACT_NONE: return launch_tc2_res<ACT, SMA_ACT_NONE, F16>(p, grid, smem, st);
    case SMA_ACT_SWISH: return launch_tc2_res<ACT, SMA_ACT_SWISH, F16>(p, grid, smem, st);
    default: return SMA_ERR_UNSUPPORTED;     // other prologue activations: gather / CUDA-core kernels
  }
}
template <int F16>
static int launch_tc2(int act, int pre, const Tc2P& p, int grid, int smem, cudaStream_t st) {
  switch (act) {
    case SMA_ACT_NONE: return launch_tc2_pre<SMA_ACT_NONE, F16>(pre, p, grid, smem, st);
    case SMA_ACT_RELU: return launch_tc2_pre<SMA_ACT_RELU, F16>(pre, p, grid, smem, st);
    case SMA_ACT_LEAKY02: return launch_tc2_pre<SMA_ACT_LEAKY02, F16>(pre, p, grid, smem, st);
    case SMA_ACT_GELU: return launch_tc2_pre<SMA_ACT_GELU, F16>(pre, p, grid, smem, st);
    case SMA_ACT_SIGMOID: return launch_tc2_pre<SMA_ACT_SIGMOID, F16>(pre, p, grid, smem, st);
    case SMA_ACT_SWISH: return launch_tc2_pre<SMA_ACT_SWISH, F16>(pre, p, grid, smem, st);
    default: return SMA_ERR_BAD_ARG;
  }
}

// persistent halo kernel (stride 1); returns SMA_ERR_UNSUPPORTED when not eligible
static int conv_tc2_try(sma_conv_desc* d, cudaStream_t st, bool f16) {
  const int kch = f16 ? 64 : KC;
  if (d->Cin % kch) return SMA_ERR_UNSUPPORTED;
  if (d->aux) {      // SFT epilogue: only on the 256-bit epilogue path, with a residual (the decoder feature) of the same geometry
    const bool ok = d->res && d->d2s <= 1 && (d->Cout & 7) == 0 && (d->out_ld & 7) == 0 && (d->out_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 31) == 0 &&
                    (d->res_ld & 7) == 0 && (d->res_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->res) & 31) == 0 &&
                    (d->aux_ld & 7) == 0 && (d->aux_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 31) == 0;
    if (!ok) return SMA_ERR_UNSUPPORTED;
  }
  // patch embedding (appmotioncodebook_arch.py:222,229,236: Rearrange + Linear = a p x p conv of stride p): every tap's operand tile is a strided gather of
  // the input, which a 5-D tensor map expresses directly; runs as a flat GEMM over the tokens with (channel chunk, tap) as the K chunks (staged-input mode only)
  const bool patch = f16 && d->stride >= 2 && d->stride == d->kh && d->kh == d->kw && !d->upsample2 && d->pad_t == 0 && d->pad_l == 0 && !d->pre_scale &&
                     d->Hi == d->Ho * d->kh && d->Wi == d->Wo * d->kw && (d->Wo % 32) == 0 && ((d->Ho * d->Wo) % BM) == 0 &&
                     (d->B == 1 || d->in_bstride == (int64_t)d->Hi * d->Wi * d->in_ld) && !(d->tc_variant & 1024);
  if (!patch && (d->stride != 1 || (d->kh != d->kw && !(d->kh == 1 || d->kw == 1)))) return SMA_ERR_UNSUPPORTED;
  const bool flat = patch || (d->kh == 1 && d->kw == 1 && !d->upsample2 && d->pad_t == 0 && d->pad_l == 0 && d->Ho == d->Hi && d->Wo == d->Wi);
  if (!flat && (d->Ho < 8 || d->Wo < 4)) return SMA_ERR_UNSUPPORTED;      // tiny feature maps: the gather kernel packs images into one tile
  Tc2P p;
  p.x = d->x; p.bias = d->bias; p.pre_scale = d->pre_scale; p.pre_shift = d->pre_shift; p.res = d->res; p.y = d->y;
  p.aux = d->aux; p.aux_bs = d->aux_bstride; p.aux_ld = d->aux_ld; p.sft_w = d->sft_w;
  p.NT = tc_ntile(d->Cout); p.ntiles_n = (d->Cout + p.NT - 1) / p.NT;
  p.wscale = f16 ? d->w_tc16 : nullptr;
  p.wtc = f16 ? d->w_tc16 + p.ntiles_n * p.NT : d->w_tc;
  p.in_bs = d->in_bstride; p.out_bs = d->out_bstride; p.res_bs = d->res_bstride;
  p.Hi = d->Hi; p.Wi = d->Wi; p.Cin = d->Cin; p.in_ld = d->in_ld; p.Cout = d->Cout; p.kh = d->kh; p.kw = d->kw;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.up = d->upsample2 ? 1 : 0; p.pre_act = d->pre_act; p.Ho = d->Ho; p.Wo = d->Wo; p.out_ld = d->out_ld;
  p.act = d->act; p.res_ld = d->res_ld; p.d2s = d->d2s;
  p.HoWo = d->Ho * d->Wo; p.cpt = d->Cin / kch; p.taps = d->kh * d->kw;
  p.patch = patch ? d->kh : 0;
  if (patch) { p.cpt *= p.taps; p.taps = 1; p.kh = p.kw = 1; }       // the kernel sees a 1x1 conv over Cin * p * p channels (weight image order: channel chunk outer, tap inner)
  p.cpt1 = p.cpt;
  if (d->x2) {        // two input tensors: staged-input mode only, plain inputs (no prologue), both channel counts multiples of 64, same geometry
    if (!f16 || patch || flat || (d->pre_scale && !d->x2_k1) || d->upsample2 || d->Cin1 <= 0 || d->Cin1 >= d->Cin || (d->Cin1 % 64) || ((d->Cin - d->Cin1) % 64) ||
        (d->in2_ld & 3) || (d->in2_bstride & 3) || d->in2_bstride == 0 || d->in_bstride == 0 || (reinterpret_cast<uintptr_t>(d->x2) & 15))
      return SMA_ERR_UNSUPPORTED;
    p.cpt1 = d->Cin1 / 64;
  }
  p.pre_ld = d->x2 ? d->Cin1 : d->Cin;
  p.x2_center = (d->x2 && d->x2_k1) ? 1 : 0;
  p.nblob = p.x2_center ? p.cpt1 * p.taps + (p.cpt - p.cpt1) : p.cpt * p.taps;
  p.passes = (d->precision == SMA_PREC_TF32 || d->precision == SMA_PREC_F16) ? 1 : (f16 && d->precision == SMA_PREC_F16X2) ? 2 : 3;
  p.flat = flat ? 1 : 0;
  if (flat) { p.tiles_x = 1; p.tiles_per_img = (p.HoWo + BM - 1) / BM; p.halo_w = 8; p.HP = BM; }
  else {
    p.tiles_x = (d->Wo + 7) / 8; p.tiles_per_img = p.tiles_x * ((d->Ho + 15) / 16);
    p.halo_w = 8 + d->kw - 1; p.HP = p.halo_w * (16 + d->kh - 1);
  }
  const long long total = (long long)d->B * p.tiles_per_img * p.ntiles_n;
  if (total > 0x7fffffffLL) return SMA_ERR_UNSUPPORTED;
  p.total_tiles = (int)total;
  p.a_img_bytes = (p.HP * 128 + 1023) & ~1023;
  p.a_stage_bytes = 2 * p.a_img_bytes;
  p.b_img_bytes = p.NT * 128;                    // NT rows of one 128-byte K-chunk (32 tf32 or 64 fp16)
  const int b_stage = 2 * p.b_img_bytes;
  int SA = 2;
  // TMA-staged input: stride 1, no fused upsample, 1x1 / 3x3, 16-byte aligned strides; as many staging slots as leave the weight ring >= 3 stages
  // (2 for the 256-column tiles); layers where that is not possible (3x3 with 128 / 256-column tiles) keep the register-staged producers
  p.NS = 0; p.parts = p.slot_rows = p.slot_bytes = p.rs = 0; p.tma_b_fixed = d->in_bstride == 0 ? 1 : 0;
  int budget = SMEM_LIMIT - SA * p.a_stage_bytes;
  if (f16 && !(d->tc_variant & 1024) && !d->upsample2 && (patch || (d->kh == d->kw && (d->kh == 1 || d->kh == 3))) && (d->in_bstride & 3) == 0) {
    if (patch) { p.parts = 4; p.rs = 0; p.slot_rows = 32; }                        // 32 consecutive tokens of one token row per box
    else if (flat) { p.parts = 4; p.rs = 0; p.slot_rows = 32; }
    else { p.parts = d->kh == 3 ? 3 : 2; p.rs = (16 + d->kh - 1) / p.parts; p.slot_rows = p.halo_w * p.rs; }
    p.slot_bytes = p.slot_rows * 256;
    const int sb_min = p.NT <= 128 ? 3 : 2;
    int NS = 2 * p.parts < MAX_NS ? 2 * p.parts : MAX_NS;
    while (NS >= 2 && (budget - NS * p.slot_bytes) / b_stage < sb_min) NS--;
    // single pass: the MMAs of a chunk are 3x shorter, so less than a whole chunk (+1 box) in flight exposes the load latency: register path instead
    p.NS = (NS >= 2 && NS > p.parts / 2 && (p.passes == 3 || NS > p.parts)) ? NS : 0;
  }
  if ((patch || d->x2) && p.NS == 0) return SMA_ERR_UNSUPPORTED;          // (the register-staged producers have no strided gather: gather kernel instead)
  int SB = (budget - p.NS * p.slot_bytes) / b_stage;
  // (a single A stage is not an option: the two producer groups could then be two barrier phases apart - parity aliasing)
  if (SB < 2) return SMA_ERR_UNSUPPORTED;
  if (SB > MAX_SB) SB = MAX_SB;
  p.SA = SA; p.SB = SB;
  p.fuse = (p.passes >= 2 && p.NT <= 128 && !(d->tc_variant & 128)) ? 1 : 0;       // tc_variant bit 7: keep the three separate MMAs (tests)
  p.acc_cols = p.fuse ? 2 * p.NT : p.NT;
  // The tensor core adds every MMA result into its fp32 accumulator with truncation toward zero (tools/acc_bias.py: the relative error of a
  // conv is a BIAS of -1.6e-8 per add for mixed-sign terms, the same on every shape from 36 to 864 adds, 10x the rounding noise of an fp32
  // FFMA chain).  The epilogue multiplies the mean back in; what remains is the data-dependent part (partial sums that stay far below or
  // above the random-walk average), at most the size of the correction itself (<= 1.4e-5 for the longest chain of this network).
  p.res_pipe = (d->res && d->act == SMA_ACT_NONE && d->d2s <= 1 && (d->Cout & 7) == 0 && (d->out_ld & 7) == 0 && (d->out_bstride & 7) == 0 &&
                (reinterpret_cast<uintptr_t>(d->y) & 31) == 0 && (d->res_ld & 7) == 0 && (d->res_bstride & 7) == 0 &&
                (reinterpret_cast<uintptr_t>(d->res) & 31) == 0 && !(d->tc_variant & 512)) ? 1 : 0;
  const int main_adds = p.nblob * 4 * (p.fuse ? 1 : p.passes);
  p.acc_corr = (d->tc_variant & 256) ? 1.f : 1.f + 1.6e-8f * (float)main_adds;
  p.dbg = (d->tc_variant >> 1) & 7;
  p.split_ws = nullptr; p.split_img_q = p.split_img_k = 0; p.split_qscale = 1.f;
  if (d->split_ws) {      // attention operand images instead of fp32 rows: flat 1x1 layers producing q (256 columns) or q | k | v (768) of E = 256, whole 128-row tiles
    const bool v8 = (d->out_ld & 7) == 0 && (d->out_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 31) == 0 && (!d->res || ((d->res_ld & 7) == 0 && (d->res_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->res) & 31) == 0));
    if (!flat || patch || d->d2s > 1 || d->out_nchw || (d->Cout != 256 && d->Cout != 768) || (p.HoWo % 128) || !v8 || d->act != SMA_ACT_NONE ||
        (reinterpret_cast<uintptr_t>(d->split_ws) & 15) || d->gn_want)
      return SMA_ERR_UNSUPPORTED;
    p.split_ws = reinterpret_cast<uint16_t*>(d->split_ws);
    p.split_img_q = (long long)d->B * p.HoWo * 256; p.split_img_k = d->Cout == 768 ? p.split_img_q : 0;
    p.split_qscale = d->split_qscale;
  }
  // fused GroupNorm partial sums: only where the 256-bit epilogue runs (every lane then walks the same column blocks) and the output is the whole
  // normalised tensor (no depth-to-space); the caller learns through gn_chunks (0 = not produced: it runs sma_groupnorm_stats instead)
  {
    const int Cq_ = d->d2s > 1 ? d->Cout / (d->d2s * d->d2s) : d->Cout;
    const bool v8 = (d->d2s <= 1 || (Cq_ & 7) == 0) && (d->Cout & 7) == 0 && (d->out_ld & 7) == 0 && (d->out_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 31) == 0 &&
                    (!d->res || ((d->res_ld & 7) == 0 && (d->res_bstride & 7) == 0 && (reinterpret_cast<uintptr_t>(d->res) & 31) == 0));
    const bool want = d->gn_want != 0 && v8 && !d->out_nchw;
    p.gn_chunks = want ? p.tiles_per_img * 4 : 0;
    p.gnp = (want && !d->plan_only) ? d->gn_partial : nullptr;
    if (want && !d->plan_only && !d->gn_partial) return SMA_ERR_BAD_ARG;
  }
  int cols = 32; while (cols < 2 * p.acc_cols + ((p.NT & 31) ? 32 : 0)) cols <<= 1;      // the epilogue reads 32 columns at a time
  if (cols > 512) return SMA_ERR_UNSUPPORTED;
  p.tmem_cols = cols;
  const int smem = SA * p.a_stage_bytes + SB * b_stage + p.NS * p.slot_bytes + 1024;
  d->gn_chunks = p.gn_chunks;                 // (past the last point where this kernel can decline the launch)
  if (d->plan_only) return SMA_OK;
  if (p.NS > 0) {
    SmaEncodeTiledFn enc = sma_tmap_encoder();
    if (!enc) return SMA_ERR_CUDA;
    const cuuint64_t nb = d->in_bstride == 0 ? 1 : (cuuint64_t)d->B;
    const cuuint64_t bstride_bytes = d->in_bstride == 0 ? (cuuint64_t)d->Hi * d->Wi * d->in_ld * 4ull : (cuuint64_t)d->in_bstride * 4ull;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult cr;
    if (patch) {
      const cuuint64_t P = (cuuint64_t)d->kh, ld4 = (cuuint64_t)d->in_ld * 4ull;
      const cuuint64_t gdim[5] = {(cuuint64_t)d->Cin, P, (cuuint64_t)d->Wo, P, (cuuint64_t)d->B * d->Ho};       // (channel, px, token x, py, frame * token rows + token y)
      const cuuint64_t gstr[4] = {ld4, P * ld4, (cuuint64_t)d->Wi * ld4, P * (cuuint64_t)d->Wi * ld4};
      const cuuint32_t box[5] = {64, 1, 32, 1, 1}, ones5[5] = {1, 1, 1, 1, 1};
      cr = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(d->x), gdim, gstr, box, ones5, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (flat) {
      const cuuint64_t gdim[3] = {(cuuint64_t)d->Cin, (cuuint64_t)p.HoWo, nb};
      const cuuint64_t gstr[2] = {(cuuint64_t)d->in_ld * 4ull, bstride_bytes};
      const cuuint32_t box[3] = {64, (cuuint32_t)p.slot_rows, 1};
      cr = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d->x), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t c1 = d->x2 ? (cuuint64_t)d->Cin1 : (cuuint64_t)d->Cin;
      const cuuint64_t gdim[4] = {c1, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, nb};
      const cuuint64_t gstr[3] = {(cuuint64_t)d->in_ld * 4ull, (cuuint64_t)d->Wi * d->in_ld * 4ull, bstride_bytes};
      const cuuint32_t box[4] = {64, (cuuint32_t)p.halo_w, (cuuint32_t)p.rs, 1};
      cr = enc(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->x), gdim, gstr, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr == CUDA_SUCCESS && d->x2) {
        const cuuint64_t gdim2[4] = {(cuuint64_t)(d->Cin - d->Cin1), (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
        const cuuint64_t gstr2[3] = {(cuuint64_t)d->in2_ld * 4ull, (cuuint64_t)d->Wi * d->in2_ld * 4ull, (cuuint64_t)d->in2_bstride * 4ull};
        cr = enc(&p.tmap2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->x2), gdim2, gstr2, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
    }
    if (cr != CUDA_SUCCESS) return SMA_ERR_CUDA;
  }
  const int g_num_sms = sma_num_sms();
  if (g_num_sms <= 0) return SMA_ERR_CUDA;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  const int pre = d->pre_scale ? d->pre_act : -1;
  return f16 ? launch_tc2<1>(d->act, pre, p, grid, smem, st) : launch_tc2<0>(d->act, pre, p, grid, smem, st);
}

// returns SMA_ERR_UNSUPPORTED when the shape / layout is not eligible (the caller then uses the CUDA-core kernel)
int sma_conv2d_tc_try(sma_conv_desc* d, cudaStream_t st) {
  if ((!d->w_tc && !d->w_tc16 && !d->plan_only) || d->out_nchw) return SMA_ERR_UNSUPPORTED;
  if ((d->Cin % KC) || (d->in_ld & 3) || (d->in_bstride & 3) || (reinterpret_cast<uintptr_t>(d->x) & 15)) return SMA_ERR_UNSUPPORTED;
  if (d->pre_scale && ((reinterpret_cast<uintptr_t>(d->pre_scale) | reinterpret_cast<uintptr_t>(d->pre_shift)) & 15)) return SMA_ERR_UNSUPPORTED;
  const long long M = (long long)d->B * d->Ho * d->Wo;
  if (M < 64) return SMA_ERR_UNSUPPORTED;
  // plan_only: report the kernel (and the tf32 image tile) this launch would use if every weight image were available; nothing is launched
  const bool want16 = (d->precision == SMA_PREC_F16X3 || d->precision == SMA_PREC_F16 || d->precision == SMA_PREC_F16X2) && (d->plan_only || (d->w_tc16 && !(reinterpret_cast<uintptr_t>(d->w_tc16) & 15)));
  if (want16 && !(d->tc_variant & 1)) {
    int r2 = conv_tc2_try(d, st, true);
    if (r2 != SMA_ERR_UNSUPPORTED) { d->kernel_used = 3; return r2; }
  }
  if (!d->plan_only && (!d->w_tc || (reinterpret_cast<uintptr_t>(d->w_tc) & 15))) return SMA_ERR_UNSUPPORTED;
  const int nt_default = tc_ntile(d->Cout);
  if (!(d->tc_variant & 1) && (d->plan_only || d->w_tc_nt == 0 || d->w_tc_nt == nt_default)) {
    int r2 = conv_tc2_try(d, st, false);
    if (r2 != SMA_ERR_UNSUPPORTED) { d->kernel_used = 2; if (d->plan_only) d->w_tc_nt = nt_default; return r2; }
  }
  if (d->aux || d->x2 || d->split_ws) return SMA_ERR_UNSUPPORTED;
  // gather kernel: one CTA per (128 rows, NT columns).  Few rows (tiny feature maps: the hourglass bottlenecks) would leave most SMs idle
  // with the widest tile, so the image may be packed with a narrower NT (more CTAs, each streaming a quarter of the weights).
  int NTg = d->w_tc_nt ? d->w_tc_nt : nt_default;
  if (d->plan_only) {
    NTg = nt_default;
    const long long mt = (M + BM - 1) / BM;
    while (NTg > 64 && (NTg % 64) == 0 && mt * ((d->Cout + NTg - 1) / NTg) < 128) NTg >>= 1;
    d->w_tc_nt = NTg; d->kernel_used = 1;
    return SMA_OK;
  }
  if (NTg < 16 || NTg > 256 || (NTg & 15)) return SMA_ERR_BAD_ARG;
  TcP p;
  p.x = d->x; p.wtc = d->w_tc; p.bias = d->bias; p.pre_scale = d->pre_scale; p.pre_shift = d->pre_shift; p.res = d->res; p.y = d->y;
  p.in_bs = d->in_bstride; p.out_bs = d->out_bstride; p.res_bs = d->res_bstride;
  p.Hi = d->Hi; p.Wi = d->Wi; p.Cin = d->Cin; p.in_ld = d->in_ld; p.Cout = d->Cout; p.kh = d->kh; p.kw = d->kw; p.stride = d->stride;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.up = d->upsample2 ? 1 : 0; p.pre_act = d->pre_act; p.Ho = d->Ho; p.Wo = d->Wo; p.out_ld = d->out_ld;
  p.act = d->act; p.res_ld = d->res_ld; p.d2s = d->d2s;
  p.M = (int)M; p.HoWo = d->Ho * d->Wo; p.cpt = d->Cin / KC; p.nchunks = d->kh * d->kw * p.cpt;
  p.NT = NTg; p.passes = (d->precision == SMA_PREC_TF32 || d->precision == SMA_PREC_F16) ? 1 : 3;
  const int ntiles = (d->Cout + p.NT - 1) / p.NT;
  const int stage_bytes = 2 * A_BYTES + 2 * p.NT * KC * 4;
  int stages = SMEM_LIMIT / stage_bytes; if (stages > MAX_STAGES) stages = MAX_STAGES; if (stages > p.nchunks) stages = p.nchunks;
  if (stages < 1) return SMA_ERR_UNSUPPORTED;
  p.stages = stages;
  int cols = 32; while (cols < p.NT) cols <<= 1;
  p.tmem_cols = cols;
  const int smem = stages * stage_bytes + 1024;
  static SmaDevOnce once;
  if (int rc = sma_opt_in_smem(once, conv_tc_kernel, SMEM_DYN_MAX)) return rc;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)ntiles);
  d->kernel_used = 1;
  conv_tc_kernel<<<grid, 192, smem, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
