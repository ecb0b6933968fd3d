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
                                 sma_stream_t stream) {
  if (!z || !codebook || !idx || N <= 0 || n_codes <= 0) return SMA_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(codebook) & 15)) return SMA_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  dim3 grid(cdiv(N, 8));
  if (E == 256) vq_lookup_kernel<256><<<grid, 256, 0, st>>>(z, N, codebook, n_codes, reinterpret_cast<long long*>(idx), zq, min_dist);
  else if (E == 32) vq_lookup_kernel<32><<<grid, 256, 0, st>>>(z, N, codebook, n_codes, reinterpret_cast<long long*>(idx), zq, min_dist);
  else return SMA_ERR_UNSUPPORTED;
  SMA_LAUNCH_CHECK();
  return SMA_OK;
}
