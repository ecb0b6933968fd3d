// VectorQuantizer lookup: pairwise L2 -> argmin -> gather (archs/vqgan_arch.py:39-73).
//
// Bit-exactness of the indices needs the reference's evaluation order: the distance is
// fl( fl(|z|^2 + |e_j|^2) - fl(2 * dot(z, e_j)) ), NOT |e_j|^2 - 2 z.e_j (the large common |z|^2 term
// quantises the distances and creates ties that torch.argmin breaks towards the lowest index;
// SURVEY.md section 7 "Bit-exact VQ indices").  All sums are plain fp32 (no TF32).  One warp per
// row of z: lanes stride over the codes, each lane keeps its best (distance, index) with strict <
// so the lowest index wins inside a lane, then a warp-shuffle argmin with (distance, index)
// lexicographic order resolves ties across lanes towards the lowest index.
#include "sma_common.cuh"
#include <math_constants.h>

namespace {


// block = 8 warps = 8 rows of z held in shared memory; codebook streamed through L2 (256 KB..1 MB, resident)
template <int E>
__global__ void __launch_bounds__(256) vq_lookup_kernel(const float* __restrict__ z, int N, const float* __restrict__ cb, int n_codes,
                                                         long long* __restrict__ idx, float* __restrict__ zq, float* __restrict__ min_dist) {
  __shared__ __align__(16) float zs[8][E];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const bool ok = row < N;
  float z2 = 0.f;
  if (ok) {
    for (int c = lane; c < E; c += 32) { float v = __ldg(z + (long long)row * E + c); zs[warp][c] = v; z2 = fmaf(v, v, z2); }
  }
  z2 = warp_sum(z2);
  __syncwarp();
  float best = CUDART_INF_F; int bi = 0x7fffffff;
  if (ok) {
    for (int j = lane; j < n_codes; j += 32) {
      const float4* e = reinterpret_cast<const float4*>(cb + (long long)j * E);
      float dot = 0.f, e2 = 0.f;
#pragma unroll 8
      for (int c = 0; c < E / 4; c++) {
        float4 ev = __ldg(e + c); float4 zv = *reinterpret_cast<const float4*>(&zs[warp][c * 4]);
        dot = fmaf(zv.x, ev.x, dot); dot = fmaf(zv.y, ev.y, dot); dot = fmaf(zv.z, ev.z, dot); dot = fmaf(zv.w, ev.w, dot);
        e2 = fmaf(ev.x, ev.x, e2); e2 = fmaf(ev.y, ev.y, e2); e2 = fmaf(ev.z, ev.z, e2); e2 = fmaf(ev.w, ev.w, e2);
      }
      float d = __fsub_rn(__fadd_rn(z2, e2), __fmul_rn(2.f, dot));
      if (d < best) { best = d; bi = j; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (ok) {
    if (lane == 0) { idx[row] = bi; if (min_dist) min_dist[row] = best; }
    if (zq) for (int c = lane; c < E; c += 32) zq[(long long)row * E + c] = __ldg(cb + (long long)bi * E + c);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------------------
// Tiled form of the same lookup for large N (the training forward quantises 65 536 tokens per call): the warp-per-row kernel above reads
// the codebook with 32 lanes on 32 different rows (16 bytes each at a 1 KB stride) and runs at 8 % of the FP32 rate (ncu: 16 ms for
// 65 536 x 1024 x 256).  Here a block keeps 128 rows of z in shared memory, streams the codebook through k-major chunks and every thread
// accumulates 8 rows x 8 codes in registers - an exact-fp32 SGEMM tile.  BIT-IDENTICAL distances: each (row, code) dot product is the same
// fmaf chain over k = 0 .. E-1 as above, |e|^2 the same ascending fmaf chain, |z|^2 the same lane-strided partial sums + xor-shuffle tree,
// d = fl(fl(|z|^2 + |e|^2) - fl(2 dot)); per thread the codes are visited in ascending order with strict <, and the 16 threads that share
// a row resolve ties towards the lowest index: the argmin is the one of the kernel above.
// ---------------------------------------------------------------------------------------------------------------------------------------
constexpr int VT_BM = 128, VT_BN = 128, VT_BK = 16, VT_THREADS = 256;

__global__ void vq_e2_kernel(const float* __restrict__ cb, int n_codes, int E, float* __restrict__ e2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_codes) return;
  const float4* e = reinterpret_cast<const float4*>(cb + (long long)j * E);
  float a = 0.f;
  for (int c = 0; c < E / 4; c++) { const float4 v = __ldg(e + c); a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a); }
  e2[j] = a;
}

template <int E>
__global__ void __launch_bounds__(VT_THREADS, 1) vq_lookup_tiled_kernel(const float* __restrict__ z, int N, const float* __restrict__ cb, int n_codes,
                                                                         const float* __restrict__ e2, long long* __restrict__ idx, float* __restrict__ zq,
                                                                         float* __restrict__ min_dist) {
  constexpr int ZP = E + 1;                                 // row pitch of the z tile (floats): rows 16 apart fall into different banks
  extern __shared__ float sm[];
  float* Zs = sm;                                           // [128][E + 1]
  float* Bs = sm + VT_BM * ZP;                              // [2][16][128]  codebook chunk, k-major
  float* z2s = Bs + 2 * VT_BK * VT_BN;                      // [128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & 15, ty = tid >> 4;                   // thread (ty, tx): rows i * 16 + ty, codes j * 16 + tx
  const int row0 = blockIdx.x * VT_BM;
  // z tile -> shared memory (rows past N: zeros), and |z|^2 exactly as the warp-per-row kernel computes it (warp w: rows w * 16 .. + 15)
  for (int rr = 0; rr < 16; rr++) {
    const int r = warp * 16 + rr; const int row = row0 + r;
    float p2 = 0.f;
    for (int c = lane; c < E; c += 32) { const float v = row < N ? __ldg(z + (long long)row * E + c) : 0.f; Zs[r * ZP + c] = v; p2 = fmaf(v, v, p2); }
    p2 = warp_sum(p2);
    if (lane == 0) z2s[r] = p2;
  }
  float best[8]; int bi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { best[i] = CUDART_INF_F; bi[i] = 0x7fffffff; }
  const int ntile = (n_codes + VT_BN - 1) / VT_BN, nk = E / VT_BK;
  // loader mapping of a chunk (128 codes x 16 k): thread -> code = tid & 127, k quad pair = tid >> 7 (2 float4 each)
  const int lcode = tid & 127, lq = tid >> 7;
  auto load_chunk = [&](int tile, int kc, float4 (&v)[2]) {
    const int code = tile * VT_BN + lcode;
#pragma unroll
    for (int h = 0; h < 2; h++)
      v[h] = code < n_codes ? __ldg(reinterpret_cast<const float4*>(cb + (long long)code * E + kc * VT_BK + (lq * 2 + h) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_chunk = [&](int buf, const float4 (&v)[2]) {
    float* b = Bs + buf * VT_BK * VT_BN;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = (lq * 2 + h) * 4;
      b[(k + 0) * VT_BN + lcode] = v[h].x; b[(k + 1) * VT_BN + lcode] = v[h].y; b[(k + 2) * VT_BN + lcode] = v[h].z; b[(k + 3) * VT_BN + lcode] = v[h].w;
    }
  };
  __syncthreads();
  for (int tile = 0; tile < ntile; tile++) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    float4 pre[2];
    load_chunk(tile, 0, pre);
    store_chunk(0, pre);
    __syncthreads();
    for (int kc = 0; kc < nk; kc++) {
      const int buf = kc & 1;
      if (kc + 1 < nk) load_chunk(tile, kc + 1, pre);          // next chunk in flight during the FMAs
      const float* b = Bs + buf * VT_BK * VT_BN;
#pragma unroll
      for (int kk = 0; kk < VT_BK; kk++) {
        float a[8], w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = Zs[(i * 16 + ty) * ZP + kc * VT_BK + kk];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = b[kk * VT_BN + j * 16 + tx];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      if (kc + 1 < nk) store_chunk(buf ^ 1, pre);
      __syncthreads();
    }
    // distances of this code tile; codes ascending per thread, strict < keeps the lowest index
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int code = tile * VT_BN + j * 16 + tx;
      if (code < n_codes) {
        const float ee = __ldg(e2 + code);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float d = __fsub_rn(__fadd_rn(z2s[i * 16 + ty], ee), __fmul_rn(2.f, acc[i][j]));
          if (d < best[i]) { best[i] = d; bi[i] = code; }
        }
      }
    }
  }
  // the 16 threads (tx) of a row: lexicographic (distance, index) minimum inside each half-warp
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best[i], o); const int oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
      if (ob < best[i] || (ob == best[i] && oi < bi[i])) { best[i] = ob; bi[i] = oi; }
    }
    const int row = row0 + i * 16 + ty;
    if (tx == 0 && row < N) { idx[row] = bi[i]; if (min_dist) min_dist[row] = best[i]; }
    if (zq && row < N) for (int c = tx; c < E; c += 16) zq[(long long)row * E + c] = __ldg(cb + (long long)bi[i] * E + c);
  }
}

// Forward values of what VectorQuantizer.forward returns besides the indices (archs/vqgan_arch.py:76-80): the straight-through tensor z + (z_q - z)
// and the loss beta * mean((z_q - z)^2) + mean((z_q - z)^2).  Deterministic two-stage sum: per-block partials in a fixed order, then one block.
constexpr int VQL_BLOCKS = 592, VQL_THREADS = 256;
__global__ void __launch_bounds__(VQL_THREADS) vq_st_partial_kernel(const float* __restrict__ z, const float* __restrict__ zq, long long n,
                                                                     float* __restrict__ st, float* __restrict__ partial) {
  __shared__ float red[VQL_THREADS / 32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * VQL_THREADS + threadIdx.x; i < n; i += (long long)VQL_BLOCKS * VQL_THREADS) {
    const float a = __ldg(z + i), d = __fsub_rn(__ldg(zq + i), a);
    if (st) st[i] = __fadd_rn(a, d);
    acc = fmaf(d, d, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < VQL_THREADS / 32; w++) t += red[w]; partial[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(32) vq_loss_final_kernel(const float* __restrict__ partial, long long n, float beta, float* __restrict__ loss) {
  float t = 0.f;
  for (int i = threadIdx.x; i < VQL_BLOCKS; i += 32) t += partial[i];
  t = warp_sum(t);
  if (threadIdx.x == 0) { const float m = t / (float)n; loss[0] = __fadd_rn(__fmul_rn(beta, m), m); }
}

}  // namespace

extern "C" int sma_vq_workspace_floats(void) { return VQL_BLOCKS; }

extern "C" int sma_vq_commit_fwd(const float* z, const float* zq, int64_t n, float beta, float* zq_st, float* workspace, float* loss, sma_stream_t stream) {
  if (!z || !zq || !workspace || !loss || n <= 0) return SMA_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  vq_st_partial_kernel<<<VQL_BLOCKS, VQL_THREADS, 0, st>>>(z, zq, (long long)n, zq_st, workspace);
  SMA_LAUNCH_CHECK();
  vq_loss_final_kernel<<<1, 32, 0, st>>>(workspace, (long long)n, beta, loss);
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}

extern "C" int sma_vq_lookup_fwd(const float* z, int N, int E, const float* codebook, int n_codes, int64_t* idx, float* zq, float* min_dist,
                                 float* workspace, sma_stream_t stream) {
  if (!z || !codebook || !idx || N <= 0 || n_codes <= 0) return SMA_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(codebook) & 15)) return SMA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  if (workspace && N >= 4 * VT_BM && (E == 256 || E == 32) && !(reinterpret_cast<uintptr_t>(workspace) & 15)) {
    // tiled exact-fp32 form (bit-identical distances and argmin); workspace: n_codes floats for |e|^2
    vq_e2_kernel<<<cdiv(n_codes, 128), 128, 0, st>>>(codebook, n_codes, E, workspace);
    SMA_LAUNCH_CHECK();
    const int smem = (VT_BM * (E + 1) + 2 * VT_BK * VT_BN + VT_BM) * (int)sizeof(float);
    if (E == 256) {
      static SmaDevOnce once;
      if (int rc = sma_opt_in_smem(once, vq_lookup_tiled_kernel<256>, smem)) return rc;
      vq_lookup_tiled_kernel<256><<<cdiv(N, VT_BM), VT_THREADS, smem, st>>>(z, N, codebook, n_codes, workspace, reinterpret_cast<long long*>(idx), zq, min_dist);
    } else {
      vq_lookup_tiled_kernel<32><<<cdiv(N, VT_BM), VT_THREADS, smem, st>>>(z, N, codebook, n_codes, workspace, reinterpret_cast<long long*>(idx), zq, min_dist);
    }
    SMA_LAUNCH_CHECK();
    return SMA_OK;
  }
  dim3 grid(cdiv(N, 8));
  if (E == 256) vq_lookup_kernel<256><<<grid, 256, 0, st>>>(z, N, codebook, n_codes, reinterpret_cast<long long*>(idx), zq, min_dist);
  else if (E == 32) vq_lookup_kernel<32><<<grid, 256, 0, st>>>(z, N, codebook, n_codes, reinterpret_cast<long long*>(idx), zq, min_dist);
  else return SMA_ERR_UNSUPPORTED;
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
