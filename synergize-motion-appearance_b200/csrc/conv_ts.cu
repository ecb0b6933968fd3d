// Weight-stationary-operand ("TS") tcgen05 convolution: fp16-split implicit GEMM with the WEIGHTS as the A operand in tensor
// memory and the PIXELS as the B operand in shared memory.
//
// Why.  With both operands in shared memory (conv_tc.cu) every kind::f16 MMA of a 128 x N tile reads (128 + N) * 32 bytes of
// shared memory, and the tensor core's shared-memory read path moves ~64 B/cycle: 192 cycles for N = 256 (128 of math), 128 for
// N = 128 (64 of math), 96 for N = 64 (32 of math) - measured on every layer of this network, and unchanged when all global
// memory traffic is switched off.  Here the GEMM is transposed,  D[cout][pixel] = sum_k W[cout][k] * X[pixel][k] :
//   * A = 128 output channels x 16 k of weights, read from TENSOR MEMORY (no shared-memory traffic at all);
//   * B = 128 pixels x 16 k: the same halo-resident, tap-shifted shared-memory descriptors as conv_tc.cu (the im2col is still done
//     by the descriptor), 4 KB per MMA = 64 cycles = exactly the math time of an M128 x N128 x K16 MMA;
//   * Cout <= 64: the hi and lo fp16 halves of the weights are STACKED along M (rows 0-63 = hi, rows 64-127 = lo), so the
//     otherwise half-empty tensor core computes w_hi*x and w_lo*x in one instruction: 2 MMAs per k-step instead of 3
//     (and the w_lo*x_lo term comes for free).
// Weights stream L2 -> registers -> tensor memory (tcgen05.st) through 4 writer warps; a ring of 8 x 32 columns holds the units.
//
// One CTA per SM, persistent over tiles of 8 x 16 output pixels (1x1 convs: 128 consecutive pixels) x 128 output channels.
//   warps 0-7   epilogue, two groups on alternate 32-pixel blocks (warp w owns TMEM lanes 32(w%4)..): accumulator rows are OUTPUT
//               CHANNELS, so a thread holds one channel for 32 pixels; scale/bias/activation are per-thread scalars; the 32 x 32 block is
//               transposed through a warp-private staging tile so that residual loads and stores are 128 contiguous bytes (32 channels)
//               per pixel
//   warps 8-11  weight writers (warp w owns TMEM lanes 32(w%4)..): ld.global 128 B per lane -> tcgen05.st.32x32b.x32
//   warp 12     TMEM allocation + MMA issue (one elected lane)
//   warps 13-19 halo producers (two groups on alternate channel chunks): load, prologue, fp16 hi/lo split, swizzled store
#include "sma_common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int TS_PIX = 128;                              // pixels per tile = UMMA N
constexpr int TS_PROD_WARPS = 7, TS_PGROUPS = 2;     // 20 warps in all = 5 per scheduler: 96 registers per thread
constexpr int TS_THREADS = 32 * (8 + 4 + 1 + TS_PROD_WARPS);   // 640
constexpr int TS_PROD_T0 = 32 * 13;                     // first producer thread
constexpr int TS_PROWS = 4 * TS_PROD_WARPS / TS_PGROUPS;      // halo rows per pass of one group (8 lanes per 128-byte row)
constexpr int TS_UNROLL = 3;                            // halo rows in flight per producer thread (two float4 each)
constexpr int TS_MAX_SA = 4;
constexpr int TS_WCOL0 = 256;                           // TMEM: [0,128) accumulator 0 | [128,256) accumulator 1 | [256,512) weight ring
constexpr int TS_WCOLS = 256;
constexpr int TS_STG = 8 * 4096;                        // epilogue staging: one 32 x 32 fp32 tile per epilogue warp
constexpr int TS_SMEM_DYN_MAX = 232448 - 2048;

struct TsP {
  const float* x; const float* wts; const float* wscale; const float* bias; const float* pre_scale; const float* pre_shift; const float* res; float* y;
  long long in_bs, out_bs, res_bs;
  int Hi, Wi, Cin, in_ld, Cout, kh, kw, pad_t, pad_l, up;
  int Ho, Wo, out_ld, res_ld, d2s;
  int HoWo, cpt, taps, MB, passes, upt, ustride, nslots;
  int flat, tiles_x, tiles_per_img, total_tiles;
  int halo_w, HP, a_img_bytes, a_stage_bytes, SA, stg_off, dbg;
  int resident;               // 1: all weight units of a 128-channel block stay in tensor memory (loaded once per block change)
  int nacc;                   // accumulators in tensor memory: 2 (epilogue of tile t overlaps the MMAs of t+1) or 1 (256-pixel tiles)
  int npx, tile_h, wcol0, tiles_m;   // pixels per tile (= UMMA N), tile height, first weight column, M tiles (images x tiles per image)
};

__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}

// profiling aid (tools/tc_dbg.py): SM cycles and nanoseconds the MMA warp of CTA 0 spent in its tile loop during the last launch
__device__ long long g_ts_prof[8];

template <int ACT, int PRE, int STACKED>
__global__ void __launch_bounds__(TS_THREADS, 1) conv_ts_kernel(const TsP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TS_MAX_SA + 16 + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (TS_MAX_SA + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * TS_MAX_SA + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (2 * TS_MAX_SA + 8 + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * TS_MAX_SA + 16 + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * TS_MAX_SA + 18 + s); };
  const uint32_t a_ring = sbase;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.SA; s++) { mbar_init(a_full(s), 32 * TS_PROD_WARPS / TS_PGROUPS); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.nslots; s++) { mbar_init(w_full(s), p.resident ? 4 * p.cpt * p.taps * p.upt : 4 * p.upt); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < 2; s++) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // tile -> (image b, first output row / column or first flat pixel, block of 128 output channels)
  auto decode = [&](int tile, int& b, int& ty0, int& tx0, int& mb) {
    int mt;
    if (p.resident) { mb = tile / p.tiles_m; mt = tile - mb * p.tiles_m; }     // channel-block major: weights change rarely
    else { mb = tile % p.MB; mt = tile / p.MB; }
    b = mt / p.tiles_per_img; int t = mt - b * p.tiles_per_img;
    if (p.flat) { ty0 = t * p.npx; tx0 = 0; } else { int tyi = t / p.tiles_x; ty0 = tyi * p.tile_h; tx0 = (t - tyi * p.tiles_x) * 8; }
  };
  const int my_tiles = blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < 8) {
    // =============================== epilogue ===============================
    const int eg = warp >> 2, wq = warp & 3;               // group (alternate 32-pixel blocks), TMEM lane quarter
    const bool vec_ok = (p.out_ld & 3) == 0 && (p.out_bs & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (!p.res || ((p.res_ld & 3) == 0 && (p.res_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    const int Cq = p.d2s > 1 ? p.Cout / (p.d2s * p.d2s) : p.Cout;
    const uint32_t stg0 = sbase + (uint32_t)p.stg_off + (uint32_t)eg * 16384u;      // this group's four staging tiles
    const uint32_t stg = stg0 + (uint32_t)wq * 4096u;      // this warp's 32 pixels x 32 rows staging tile (pixel-major)
    const bool simple = vec_ok && p.d2s <= 1 && (p.Cout & 3) == 0;                  // plain NHWC float4 stores
    int tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      int b, ty0, tx0, mb; decode(tile, b, ty0, tx0, mb);
      const int ab = p.nacc == 2 ? (tcount & 1) : 0; const uint32_t aph = (uint32_t)(p.nacc == 2 ? (tcount >> 1) : tcount) & 1u;
      const float* resb = p.res ? p.res + (long long)b * p.res_bs : nullptr;
      float* yb = p.y + (long long)b * p.out_bs;
      // four consecutive output channels n.. of tile pixel mm (values already scaled / biased / activated): residual add + store
      auto emit = [&](int mm, int n, float (&o)[4]) {
        int oy = 0, ox = 0, r; bool mok;
        if (p.flat) { r = ty0 + mm; mok = r < p.HoWo; if (p.d2s > 1) { oy = r / p.Wo; ox = r - oy * p.Wo; } }
        else { oy = ty0 + (mm >> 3); ox = tx0 + (mm & 7); mok = oy < p.Ho && ox < p.Wo; r = oy * p.Wo + ox; }
        if (!mok || n >= p.Cout || mm >= p.npx) return;
        if (simple) {
          if (resb) {
            float4 rv = __ldg(reinterpret_cast<const float4*>(resb + (long long)r * p.res_ld + n));
            o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
          }
          *reinterpret_cast<float4*>(yb + (long long)r * p.out_ld + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else if (vec_ok && n + 4 <= p.Cout && (Cq & 3) == 0) {
          float* dst;
          if (p.d2s > 1) {
            int qd = n / Cq; int cval = n - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
            long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
            dst = yb + pix * p.out_ld + cval;
          } else {
            dst = yb + (long long)r * p.out_ld + n;
          }
          if (resb) {
            float4 rv = __ldg(reinterpret_cast<const float4*>(resb + (long long)r * p.res_ld + n));
            o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
          }
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (n + e >= p.Cout) break;
            float val = o[e];
            if (resb) val += __ldg(resb + (long long)r * p.res_ld + n + e);
            if (p.d2s > 1) {
              int nn = n + e; int qd = nn / Cq; int c2 = nn - qd * Cq; int p1 = qd / p.d2s, p2 = qd - p1 * p.d2s;
              long long pix = (long long)(oy * p.d2s + p1) * (p.Wo * p.d2s) + (ox * p.d2s + p2);
              yb[pix * p.out_ld + c2] = val;
            } else {
              yb[(long long)r * p.out_ld + n + e] = val;
            }
          }
        }
      };
      // per-thread constants
      float sc = 0.f, bi = 0.f;                            // plain: this lane's accumulator row = one output channel
      float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), bi4 = sc4;    // stacked: the four channels this lane stores
      const int cbase = mb * 128 + wq * 32;
      if (!STACKED) {
        const int c_lane = cbase + lane;
        sc = c_lane < p.Cout ? __ldg(p.wscale + c_lane) : 0.f;
        bi = (p.bias && c_lane < p.Cout) ? __ldg(p.bias + c_lane) : 0.f;
      } else {
        const int n = (lane & 15) * 4;
        const float* ws = p.wscale + n;
        sc4 = make_float4(n < p.Cout ? __ldg(ws) : 0.f, n + 1 < p.Cout ? __ldg(ws + 1) : 0.f, n + 2 < p.Cout ? __ldg(ws + 2) : 0.f, n + 3 < p.Cout ? __ldg(ws + 3) : 0.f);
        if (p.bias) bi4 = make_float4(n < p.Cout ? __ldg(p.bias + n) : 0.f, n + 1 < p.Cout ? __ldg(p.bias + n + 1) : 0.f,
                                      n + 2 < p.Cout ? __ldg(p.bias + n + 2) : 0.f, n + 3 < p.Cout ? __ldg(p.bias + n + 3) : 0.f);
      }
      mbar_wait(acc_full(ab), aph);
      tc_fence_after();
      if (p.dbg & 4) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(acc_empty(ab)); continue; }
      bool released = false;
#pragma unroll 1
      for (int px0 = eg * 32; px0 < p.npx; px0 += 64) {
        uint32_t a[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * p.npx + px0), a);
        tmem_ld_wait();
        if (px0 + 64 >= p.npx) {                           // this warp's last block: its share of the accumulator has been read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(ab));
          released = true;
        }
        if (!STACKED) {
          // rows = 32 channels of this warp: scale / bias / activation per thread, transpose through the warp's own tile
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float v = sma_act(fmaf(__uint_as_float(a[j]), sc, bi), ACT);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + (uint32_t)(j * 32 + lane) * 4u), "f"(v) : "memory");
          }
          __syncwarp();
          const int q = lane & 7, rsub = lane >> 3;
#pragma unroll
          for (int t = 0; t < 8; t++) {
            const int pl = t * 4 + rsub;                    // pixel within this block of 32
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(stg + (uint32_t)(pl * 32 + q * 4) * 4u));
            float o[4] = {v.x, v.y, v.z, v.w};
            emit(px0 + pl, cbase + q * 4, o);
          }
          __syncwarp();
        } else {
          // rows 0-63 (quarters 0,1) = w_hi * x, rows 64-127 (quarters 2,3) = w_lo * x of channels 0..63: every warp of the group parks its
          // raw block, then each warp finishes 8 of the 32 pixels for all 64 channels (hi + lo, scale, bias, activation, residual, store)
#pragma unroll
          for (int j = 0; j < 32; j++) asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + (uint32_t)(j * 32 + lane) * 4u), "r"(a[j]) : "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
          const int q = lane & 15, half = q >> 3;          // channels 4q..4q+3 live in tile `half` (hi) and `half + 2` (lo)
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const int pl = wq * 8 + t * 2 + (lane >> 4);
            const uint32_t off = (uint32_t)(pl * 32 + (q & 7) * 4) * 4u;
            float4 vh, vl;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vh.x), "=f"(vh.y), "=f"(vh.z), "=f"(vh.w) : "r"(stg0 + (uint32_t)half * 4096u + off));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vl.x), "=f"(vl.y), "=f"(vl.z), "=f"(vl.w) : "r"(stg0 + (uint32_t)(half + 2) * 4096u + off));
            float o[4] = {sma_act(fmaf(vh.x + vl.x, sc4.x, bi4.x), ACT), sma_act(fmaf(vh.y + vl.y, sc4.y, bi4.y), ACT),
                          sma_act(fmaf(vh.z + vl.z, sc4.z, bi4.z), ACT), sma_act(fmaf(vh.w + vl.w, sc4.w, bi4.w), ACT)};
            emit(px0 + pl, q * 4, o);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");     // staging tiles are rewritten by the next pixel block
        }
      }
      if (!released) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(acc_empty(ab)); }   // group without a block in this tile
    }
  } else if (warp < 12) {
    // =============================== weight writers: L2 -> registers -> tensor memory ===============================
    // two units in registers: the next unit's loads are in flight while the current one is written
    const int qd = warp & 3;                               // TMEM lane quarter of this warp (= warp % 4)
    constexpr int wgroup = 0, WG = 1;
    const int m = qd * 32 + lane;                          // accumulator row / A row fed by this thread
    const int upt_tile = p.cpt * p.taps * p.upt;           // units per tile
    const long long total = (long long)my_tiles * upt_tile;
    // flattened unit g of this CTA -> its 128 bytes of row m
    auto unit_ptr = [&](long long g) -> const uint4* {
      const int ti = (int)(g / upt_tile); const int j = (int)(g - (long long)ti * upt_tile);
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int mb = tile % p.MB;
      const int blob = j / p.upt, within = j - blob * p.upt;
      const long long unit = ((long long)mb * p.cpt * p.taps + blob) * p.ustride + within;
      return reinterpret_cast<const uint4*>(p.wts + unit * 4096) + m;      // piece i of row m lives at uint4 index i * 128 + m
    };
    auto load = [&](long long g, uint32_t (&r)[32]) {
      if (p.dbg & 1) return;
      const uint4* src = unit_ptr(g);
#pragma unroll
      for (int i = 0; i < 8; i++) { uint4 v = __ldg(src + i * 128); r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w; }
    };
    auto put = [&](long long g, const uint32_t (&r)[32]) {
      const long long blob = g / p.upt; const int within = (int)(g - blob * p.upt);
      const int slot = (int)(blob % p.nslots); const uint32_t ph = (uint32_t)(blob / p.nslots) & 1u;
      mbar_wait(w_empty(slot), ph ^ 1u);                   // the MMAs that read this slot have retired
      tc_fence_after();
      tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(p.wcol0 + slot * 32 * p.upt + within * 32), r);
      tmem_st_wait();                                      // warp-collective: every lane's rows have landed
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_full(slot));            // one arrival per warp: 256 per-thread arrivals per tap on one shared-memory word
    };                                                     // serialise in the shared-memory pipe that also feeds the MMA's pixel operand
    uint32_t r0[32], r1[32];
#pragma unroll
    for (int i = 0; i < 32; i++) { r0[i] = 0u; r1[i] = 0u; }
    if (p.resident) {
      // one epoch per run of tiles with the same channel block: (re)load all its units; the two groups take alternate units
      int cur_mb = -1, epoch = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int mb = tile / p.tiles_m;
        if (mb == cur_mb) continue;
        cur_mb = mb;
        if (epoch > 0) { mbar_wait(w_empty(0), (uint32_t)(epoch - 1) & 1u); tc_fence_after(); }    // every MMA on the old weights has retired
        for (int u = wgroup; u < upt_tile; u += WG) {
          if (!(p.dbg & 1)) {
            const int blob = u / p.upt, within = u - blob * p.upt;
            const long long unit = ((long long)mb * p.cpt * p.taps + blob) * p.ustride + within;
            const uint4* src = reinterpret_cast<const uint4*>(p.wts + unit * 4096) + m;
#pragma unroll
            for (int i = 0; i < 8; i++) { uint4 v = __ldg(src + i * 128); r0[4 * i] = v.x; r0[4 * i + 1] = v.y; r0[4 * i + 2] = v.z; r0[4 * i + 3] = v.w; }
          }
          tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(p.wcol0 + u * 32), r0);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(w_full(0));
        }
        ++epoch;
      }
    }
    long long g = p.resident ? total : wgroup;
    if (g < total) load(g, r0);
    for (; g < total; g += 2 * WG) {
      if (g + WG < total) load(g + WG, r1);                // the next unit's loads are in flight while this one is written
      put(g, r0);
      if (g + WG < total) {
        if (g + 2 * WG < total) load(g + 2 * WG, r0);
        put(g + WG, r1);
      }
    }
  } else if (warp == 12) {
    // =============================== MMA issuer ===============================
    // ONE elected lane runs the whole tile loop (the elect region encloses the loops: no per-tap warp-sync / elect / reconvergence),
    // with wrap-around counters instead of divisions: this scalar instruction stream has to stay well below the 768 tensor cycles of
    // a tap, or the tensor pipe drains while the thread computes descriptors (measured: 103 instead of 64 cycles per MMA).
    if (elect_one_sync()) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.npx >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // f16 x f16 -> f32, M = 128, N = npx
      // pixel operand descriptor (K-major, SWIZZLE_128B): 8-row atoms are 8 consecutive halo rows; next atom = next output row
      const uint64_t x_desc_hi = ((uint64_t)((uint32_t)((p.halo_w * 128) >> 4) | (1u << 14) | (2u << 29))) << 32;
      const uint32_t row_step = (uint32_t)(p.halo_w * 128) >> 4;          // one halo row down, in 16-byte units
      const uint32_t slot_cols = (uint32_t)(32 * p.upt);
      const bool three = p.passes == 3;
      uint32_t slot = 0, wph = 0, sa = 0, apha = 0, tcount = 0;
      int cur_mb = -1; uint32_t epoch = 0;
      long long prof_c0 = 0, prof_g0 = 0;
      if (blockIdx.x == 0) { prof_c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_g0)); }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t ab = p.nacc == 2 ? (tcount & 1u) : 0u;
        mbar_wait(acc_empty(ab), ((p.nacc == 2 ? (tcount >> 1) : tcount) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * (uint32_t)p.npx;
        uint32_t acc = 0u;                                   // first MMA of the tile overwrites the accumulator
        if (p.resident) {
          const int mb = tile / p.tiles_m;
          if (mb != cur_mb) {                                // new channel block: hand the weight columns back, wait for the new set
            if (cur_mb >= 0) tc_commit(w_empty(0));
            cur_mb = mb;
            mbar_wait(w_full(0), epoch & 1u);
            tc_fence_after();
            ++epoch;
          }
          slot = 0;                                          // unit index within the resident set
        }
        for (int cc = 0; cc < p.cpt; cc++) {
          mbar_wait(a_full(sa), apha);
          const uint32_t x_hi = a_ring + sa * (uint32_t)p.a_stage_bytes;
          const uint32_t lo_h = ((x_hi & 0x3FFFFu) >> 4) | (1u << 16), lo_l = (((x_hi + (uint32_t)p.a_img_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          uint32_t roff = 0;
          for (int ky = 0; ky < p.kh; ky++, roff += row_step) {
            for (int kx = 0; kx < p.kw; kx++) {
              if (!p.resident) { mbar_wait(w_full(slot), wph); tc_fence_after(); }
              const uint64_t dxh = x_desc_hi | (uint64_t)(lo_h + roff + (uint32_t)kx * 8u);
              const uint64_t dxl = x_desc_hi | (uint64_t)(lo_l + roff + (uint32_t)kx * 8u);
              const uint32_t wcol = tmem_base + (uint32_t)p.wcol0 + slot * slot_cols;
#pragma unroll
              for (int k4 = 0; k4 < 4; k4++) {               // 4 k-steps of 16 fp16 = 8 TMEM columns = 32 bytes of a halo row
                const uint64_t ko = (uint64_t)(k4 * 2);
                const uint32_t wk = wcol + (uint32_t)(k4 * 8);
                if (STACKED) {
                  tc_mma_ts_f16(d_tmem, wk, dxh + ko, idesc, acc);                         // [w_hi ; w_lo] * x_hi
                  if (three) tc_mma_ts_f16(d_tmem, wk, dxl + ko, idesc, 1u);               // [w_hi ; w_lo] * x_lo
                } else if (three) {
                  tc_mma_ts_f16(d_tmem, wk, dxl + ko, idesc, acc);                         // w_hi * x_lo
                  tc_mma_ts_f16(d_tmem, wk + 32u, dxh + ko, idesc, 1u);                    // w_lo * x_hi
                  tc_mma_ts_f16(d_tmem, wk, dxh + ko, idesc, 1u);                          // w_hi * x_hi
                } else {
                  tc_mma_ts_f16(d_tmem, wk, dxh + ko, idesc, acc);
                }
                acc = 1u;
              }
              if (p.resident) ++slot;
              else { tc_commit(w_empty(slot)); if (++slot == (uint32_t)p.nslots) { slot = 0; wph ^= 1u; } }
            }
          }
          tc_commit(a_empty(sa));
          if (++sa == (uint32_t)p.SA) { sa = 0; apha ^= 1u; }
        }
        tc_commit(acc_full(ab));
      }
      if (blockIdx.x == 0) {
        long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        g_ts_prof[0] = clock64() - prof_c0; g_ts_prof[1] = g1 - prof_g0;
      }
    }
    __syncwarp();
  } else {
    // =============================== halo producers ===============================
    constexpr int GT = 32 * TS_PROD_WARPS / TS_PGROUPS;    // threads per producer group
    const int pgroup = (threadIdx.x - TS_PROD_T0) / GT;
    const int ptid = (threadIdx.x - TS_PROD_T0) % GT; const int cq = ptid & 7; const int prow = ptid >> 3;
    const int Hv = p.Hi << p.up, Wv = p.Wi << p.up;
    const int npass = (p.HP + TS_PROWS - 1) / TS_PROWS;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int b, ty0, tx0, mb; decode(tile, b, ty0, tx0, mb);
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
        if ((it % TS_PGROUPS) != pgroup) continue;          // the other group's chunk
        const int sa = it % p.SA; const uint32_t pha = (it / p.SA) & 1;
        const uint32_t a_hi = a_ring + (uint32_t)sa * p.a_stage_bytes, a_lo = a_hi + p.a_img_bytes;
        bool waited = false;
        if (p.dbg & 2) { mbar_wait(a_empty(sa), pha ^ 1u); fence_async_smem(); mbar_arrive(a_full(sa)); continue; }
        const int c = cc * 64 + cq * 8;                     // two float4 (8 channels = one 16-byte fp16 unit) per halo row and lane
        float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sh0 = make_float4(0.f, 0.f, 0.f, 0.f), sc1 = sc0, sh1 = sh0;
        if (PRE >= 0) {
          sc0 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.Cin + c));
          sc1 = __ldg(reinterpret_cast<const float4*>(p.pre_scale + (long long)b * p.Cin + c + 4));
          sh0 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.Cin + c));
          sh1 = __ldg(reinterpret_cast<const float4*>(p.pre_shift + (long long)b * p.Cin + c + 4));
        }
        for (int pass0 = 0; pass0 < npass; pass0 += TS_UNROLL) {
          float4 v0[TS_UNROLL], v1[TS_UNROLL]; bool ok[TS_UNROLL];
#pragma unroll
          for (int u = 0; u < TS_UNROLL; u++) {
            const int hp = (pass0 + u) * TS_PROWS + prow;
            ok[u] = false; v0[u] = make_float4(0.f, 0.f, 0.f, 0.f); v1[u] = v0[u];
            if (hp < p.HP) {
              const long long pix = halo_pixel(hp, ok[u]);
              if (ok[u]) {
                const float4* src = reinterpret_cast<const float4*>(xb + pix * p.in_ld + c);
                v0[u] = __ldg(src); v1[u] = __ldg(src + 1);
              }
            }
          }
          if (!waited) { mbar_wait(a_empty(sa), pha ^ 1u); waited = true; }     // first loads are in flight while we wait
#pragma unroll
          for (int u = 0; u < TS_UNROLL; u++) {
            const int hp = (pass0 + u) * TS_PROWS + prow;
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
        fence_async_smem();
        mbar_arrive(a_full(sa));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// per output column n: factor 2^e with max_k |w[k][n]| * 2^-e in [0.5, 1); columns >= Cout get 1
__global__ void ts_colscale_kernel(const float* __restrict__ wp, int ldw, int K, int Cout, int ncols, float* __restrict__ inv_scale) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= ncols) return;
  float mx = 0.f;
  if (n < Cout) for (int k = 0; k < K; k++) mx = fmaxf(mx, fabsf(wp[(long long)k * ldw + n]));
  int e = 0;
  if (mx > 0.f && mx < 3.0e38f) { frexpf(mx, &e); e = max(-100, min(100, e)); }
  inv_scale[n] = ldexpf(1.f, e);
}

// packed [K][ldw] fp32 weight -> units of 128 rows x 64 fp16, stored as 8 pieces of [128 rows][16 bytes]:
//   plain   (Cout > 64): [block of 128 channels][64-channel chunk][tap][hi | lo], row m = channel 128*mb + m
//   stacked (Cout <= 64): [chunk][tap][one unit], rows 0-63 = hi of channel m, rows 64-127 = lo of channel m - 64
__global__ void pack_ts_kernel(const float* __restrict__ wp, int ldw, int Cin, int taps, int Cout, int MB, int stacked,
                               const float* __restrict__ inv_scale, uint16_t* __restrict__ out) {
  const int cpt = Cin / 64, UP = stacked ? 1 : 2;
  const long long total = (long long)MB * cpt * taps * UP * 128 * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(i & 63); long long t = i >> 6; int m = (int)(t & 127); t >>= 7; int u = (int)(t % UP); t /= UP;
    int tap = (int)(t % taps); t /= taps; int cc = (int)(t % cpt); int mb = (int)(t / cpt);
    const int cout = stacked ? (m & 63) : mb * 128 + m;
    const int part = stacked ? (m >> 6) : u;
    const int k = tap * Cin + cc * 64 + kk;
    float w = cout < Cout ? wp[(long long)k * ldw + cout] / inv_scale[cout] : 0.f;      // exact: the divisor is a power of two
    __half hi = __float2half_rn(w); __half lo = __float2half_rn(w - __half2float(hi));
    // unit = 8 pieces x 128 rows x 16 bytes: a warp of writer threads (32 consecutive rows) loads 512 contiguous bytes per instruction
    const long long unit = i >> 13;
    out[(unit << 13) + (long long)((kk >> 3) * 128 + m) * 8 + (kk & 7)] = __half_as_ushort(part ? lo : hi);
  }
}

template <int ACT, int PRE, int STACKED>
int launch_ts_inst(const TsP& p, int grid, int smem, cudaStream_t st) {
  static SmaDevOnce once;             // per instantiation and per device
  if (int rc = sma_opt_in_smem(once, conv_ts_kernel<ACT, PRE, STACKED>, TS_SMEM_DYN_MAX)) return rc;
  conv_ts_kernel<ACT, PRE, STACKED><<<grid, TS_THREADS, smem, st>>>(p);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
template <int ACT, int STACKED>
int launch_ts_pre(int pre, const TsP& p, int grid, int smem, cudaStream_t st) {
  switch (pre) {
    case -1: return launch_ts_inst<ACT, -1, STACKED>(p, grid, smem, st);
    case SMA_ACT_NONE: return launch_ts_inst<ACT, SMA_ACT_NONE, STACKED>(p, grid, smem, st);
    case SMA_ACT_SWISH: return launch_ts_inst<ACT, SMA_ACT_SWISH, STACKED>(p, grid, smem, st);
    default: return SMA_ERR_UNSUPPORTED;
  }
}
template <int STACKED>
int launch_ts(int act, int pre, const TsP& p, int grid, int smem, cudaStream_t st) {
  switch (act) {
    case SMA_ACT_NONE: return launch_ts_pre<SMA_ACT_NONE, STACKED>(pre, p, grid, smem, st);
    case SMA_ACT_RELU: return launch_ts_pre<SMA_ACT_RELU, STACKED>(pre, p, grid, smem, st);
    case SMA_ACT_LEAKY02: return launch_ts_pre<SMA_ACT_LEAKY02, STACKED>(pre, p, grid, smem, st);
    case SMA_ACT_GELU: return launch_ts_pre<SMA_ACT_GELU, STACKED>(pre, p, grid, smem, st);
    case SMA_ACT_SIGMOID: return launch_ts_pre<SMA_ACT_SIGMOID, STACKED>(pre, p, grid, smem, st);
    case SMA_ACT_SWISH: return launch_ts_pre<SMA_ACT_SWISH, STACKED>(pre, p, grid, smem, st);
    default: return SMA_ERR_BAD_ARG;
  }
}


}  // namespace

extern "C" int sma_debug_conv_ts_prof(long long* cycles_ns) {
  return cudaMemcpyFromSymbol(cycles_ns, g_ts_prof, sizeof(long long) * 8) == cudaSuccess ? SMA_OK : SMA_ERR_CUDA;
}

extern "C" int64_t sma_conv_weight_ts_floats(int Cout, int Cin, int kh, int kw) {
  if (Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || (Cin % 64)) return 0;
  const int stacked = Cout <= 64; const int MB = stacked ? 1 : (Cout + 127) / 128;
  return (int64_t)MB * 128 + (int64_t)MB * (Cin / 64) * kh * kw * (stacked ? 1 : 2) * 4096;
}

extern "C" int sma_pack_conv_weight_ts(const float* w_packed, int ldw, int Cout, int Cin, int kh, int kw, float* w_ts, sma_stream_t stream) {
  if (!w_packed || !w_ts || Cout <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || ldw < Cout) return SMA_ERR_BAD_ARG;
  if (Cin % 64) return SMA_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(w_ts) & 15) return SMA_ERR_BAD_ARG;
  const int stacked = Cout <= 64; const int MB = stacked ? 1 : (Cout + 127) / 128; const int ncols = MB * 128; const int K = kh * kw * Cin;
  ts_colscale_kernel<<<(ncols + 127) / 128, 128, 0, as_stream(stream)>>>(w_packed, ldw, K, Cout, ncols, w_ts);
  SMA_LAUNCH_CHECK();
  const long long total = (long long)MB * (Cin / 64) * kh * kw * (stacked ? 1 : 2) * 128 * 64;
  int blocks = (int)((total + 255) / 256); if (blocks > 8192) blocks = 8192;
  pack_ts_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_packed, ldw, Cin, kh * kw, Cout, MB, stacked, w_ts, reinterpret_cast<uint16_t*>(w_ts + ncols));
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

// returns SMA_ERR_UNSUPPORTED when the shape / layout is not eligible (the caller then tries the shared-memory-operand kernels)
int sma_conv2d_ts_try(sma_conv_desc* d, cudaStream_t st) {
  if (d->aux || d->plan_only) return SMA_ERR_UNSUPPORTED;      // (the tensor-memory-operand kernel is an opt-in experiment: never planned)
  if (!d->w_ts || d->x2 || d->split_ws || d->out_nchw || (d->precision != SMA_PREC_F16X3 && d->precision != SMA_PREC_F16) || (d->tc_variant & 1)) return SMA_ERR_UNSUPPORTED;
  if ((d->Cin % 64) || (d->in_ld & 3) || (d->in_bstride & 3) || (reinterpret_cast<uintptr_t>(d->x) & 15) || (reinterpret_cast<uintptr_t>(d->w_ts) & 15))
    return SMA_ERR_UNSUPPORTED;
  if (d->pre_scale && ((reinterpret_cast<uintptr_t>(d->pre_scale) | reinterpret_cast<uintptr_t>(d->pre_shift)) & 15)) return SMA_ERR_UNSUPPORTED;
  if (d->stride != 1 || (d->kh != d->kw && !(d->kh == 1 || d->kw == 1))) return SMA_ERR_UNSUPPORTED;
  const bool flat = d->kh == 1 && d->kw == 1 && !d->upsample2 && d->pad_t == 0 && d->pad_l == 0 && d->Ho == d->Hi && d->Wo == d->Wi;
  if (!flat && (d->Ho < 8 || d->Wo < 4)) return SMA_ERR_UNSUPPORTED;
  if ((long long)d->B * d->Ho * d->Wo < 64) return SMA_ERR_UNSUPPORTED;
  TsP p;
  const int stacked = d->Cout <= 64;
  p.MB = stacked ? 1 : (d->Cout + 127) / 128;
  p.x = d->x; p.wscale = d->w_ts; p.wts = d->w_ts + p.MB * 128; p.bias = d->bias; p.pre_scale = d->pre_scale; p.pre_shift = d->pre_shift;
  p.res = d->res; p.y = d->y;
  p.in_bs = d->in_bstride; p.out_bs = d->out_bstride; p.res_bs = d->res_bstride;
  p.Hi = d->Hi; p.Wi = d->Wi; p.Cin = d->Cin; p.in_ld = d->in_ld; p.Cout = d->Cout; p.kh = d->kh; p.kw = d->kw;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.up = d->upsample2 ? 1 : 0; p.Ho = d->Ho; p.Wo = d->Wo; p.out_ld = d->out_ld;
  p.res_ld = d->res_ld; p.d2s = d->d2s;
  p.HoWo = d->Ho * d->Wo; p.cpt = d->Cin / 64; p.taps = d->kh * d->kw;
  p.passes = d->precision == SMA_PREC_F16 ? 1 : 3;
  p.ustride = stacked ? 1 : 2;                          // units stored per (chunk, tap)
  p.upt = (!stacked && p.passes == 3) ? 2 : 1;          // units used per (chunk, tap)
  // resident weights: all units of a 128-channel block fit beside two accumulators of 128 (or, for 3x3, 112) pixels
  const int units_mb = p.cpt * p.taps * p.upt;
  p.resident = 0; p.npx = TS_PIX; p.tile_h = 16;
  if (units_mb * 32 <= 512 - 2 * TS_PIX) p.resident = 1;
  else if (!flat && units_mb * 32 <= 512 - 2 * 112) { p.resident = 1; p.npx = 112; p.tile_h = 14; }
  if (d->tc_variant & 16) { p.resident = 0; p.npx = TS_PIX; p.tile_h = 16; }      // experiments: force the streaming ring
  p.nacc = 2;
  if (!p.resident && !flat && !(d->tc_variant & 64)) {
    // streamed weights: 8 x 32 = 256-pixel tiles (one N = 256 MMA per weight k-step) halve the L2 -> SM weight traffic per MAC, which is what
    // bounds the 128-pixel variant (32 KB per tap and SM); the single accumulator exposes the epilogue drain (~10 % of a tile)
    p.npx = 256; p.tile_h = 32; p.nacc = 1;
  }
  if (!p.resident && p.nacc == 2 && !(d->tc_variant & 32)) return SMA_ERR_UNSUPPORTED;   // 128-pixel streaming ring (flat 1x1 with many channels): on request
  p.wcol0 = p.nacc * p.npx;
  p.nslots = p.resident ? 1 : TS_WCOLS / (32 * p.upt);
  p.flat = flat ? 1 : 0;
  if (flat) { p.tiles_x = 1; p.tiles_per_img = (p.HoWo + p.npx - 1) / p.npx; p.halo_w = 8; p.HP = p.npx; }
  else {
    p.tiles_x = (d->Wo + 7) / 8; p.tiles_per_img = p.tiles_x * ((d->Ho + p.tile_h - 1) / p.tile_h);
    p.halo_w = 8 + d->kw - 1; p.HP = p.halo_w * (p.tile_h + d->kh - 1);
  }
  p.tiles_m = d->B * p.tiles_per_img;
  const long long total = (long long)d->B * p.tiles_per_img * p.MB;
  if (total > 0x7fffffffLL) return SMA_ERR_UNSUPPORTED;
  p.total_tiles = (int)total;
  p.a_img_bytes = (p.HP * 128 + 1023) & ~1023;
  p.a_stage_bytes = 2 * p.a_img_bytes;
  const int avail = TS_SMEM_DYN_MAX - 1024 - TS_STG;
  int SA = avail / p.a_stage_bytes;
  // two producer groups work on alternate chunks: with a single stage a group could be two barrier phases ahead (parity aliasing)
  if (SA < 2) return SMA_ERR_UNSUPPORTED;
  if (SA > TS_MAX_SA) SA = TS_MAX_SA;
  p.SA = SA; p.stg_off = SA * p.a_stage_bytes; p.dbg = (d->tc_variant >> 1) & 7;
  const int smem = p.stg_off + TS_STG + 1024;
  const int ts_num_sms = sma_num_sms();
  if (ts_num_sms <= 0) return SMA_ERR_CUDA;
  const int grid = p.total_tiles < ts_num_sms ? p.total_tiles : ts_num_sms;
  const int pre = d->pre_scale ? d->pre_act : -1;
  return stacked ? launch_ts<1>(d->act, pre, p, grid, smem, st) : launch_ts<0>(d->act, pre, p, grid, smem, st);
}
