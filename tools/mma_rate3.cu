// Scratch micro-benchmark 3: cycles per TS-mode MMA (M128 x N128 x K16, kind::f16) for different operand ORDERS within a tap of
// 12 MMAs: A units hi (cols 256..) / lo (cols 288..), B images x_hi (sbase) / x_lo (sbase + 24 KB).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../synergize-motion-appearance_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int PAT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  for (int i = threadIdx.x; i < 60 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hb = (1ull << 16) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      const uint32_t aoff = (uint32_t)((it % 9) / 3 * 10 + (it % 3)) * 128u;
      const uint64_t dxh = hb | (uint64_t)(((sbase + aoff) & 0x3FFFF) >> 4), dxl = hb | (uint64_t)(((sbase + 24576 + aoff) & 0x3FFFF) >> 4);
      const uint32_t d = tm + (uint32_t)(((it / 18) & 1) * 128);
      const uint32_t wcol = tm + 256u + (uint32_t)((it & 3) * 64);
      if (elect_one_sync()) {
        if (PAT == 0) {          // the conv kernel's order
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t ko = k4 * 2; const uint32_t wk = wcol + k4 * 8;
            mma_ts(d, wk, dxl + ko, idesc, 1u); mma_ts(d, wk + 32u, dxh + ko, idesc, 1u); mma_ts(d, wk, dxh + ko, idesc, 1u);
          }
        } else if (PAT == 1) {   // B-stationary pairs: (hi, lo) x x_hi, then hi x x_lo
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t ko = k4 * 2; const uint32_t wk = wcol + k4 * 8;
            mma_ts(d, wk, dxh + ko, idesc, 1u); mma_ts(d, wk + 32u, dxh + ko, idesc, 1u); mma_ts(d, wk, dxl + ko, idesc, 1u);
          }
        } else if (PAT == 2) {   // grouped by product
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) mma_ts(d, wcol + k4 * 8, dxh + (uint64_t)(k4 * 2), idesc, 1u);
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) mma_ts(d, wcol + 32u + k4 * 8, dxh + (uint64_t)(k4 * 2), idesc, 1u);
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) mma_ts(d, wcol + k4 * 8, dxl + (uint64_t)(k4 * 2), idesc, 1u);
        } else if (PAT == 3) {   // same B and same A every time (best case)
#pragma unroll
          for (int k = 0; k < 12; k++) mma_ts(d, wcol, dxh, idesc, 1u);
        } else if (PAT == 4) {   // alternate accumulators between consecutive MMAs (two half tiles)
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t ko = k4 * 2; const uint32_t wk = wcol + k4 * 8;
            mma_ts(tm, wk, dxl + ko, idesc, 1u); mma_ts(tm + 128u, wk + 32u, dxh + ko, idesc, 1u); mma_ts(tm, wk, dxh + ko, idesc, 1u);
          }
        } else if (PAT == 5) {   // kernel order with the unswizzled-pitch descriptor (SBO 1024, aligned)
          const uint64_t eh = make_desc(sbase), el = make_desc(sbase + 24576);
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {
            const uint64_t ko = k4 * 2; const uint32_t wk = wcol + k4 * 8;
            mma_ts(d, wk, el + ko, idesc, 1u); mma_ts(d, wk + 32u, eh + ko, idesc, 1u); mma_ts(d, wk, eh + ko, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int PAT> void run(long long* d, const char* name) {
  cudaFuncSetAttribute(rate_kernel<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 3000;
  rate_kernel<PAT><<<148, 128, 64 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("pattern %d %-50s: %.1f cycles/MMA (%s)\n", PAT, name, (double)c / (iters * 12.0), cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  run<0>(d, "kernel order hi*xl, lo*xh, hi*xh per k-step");
  run<1>(d, "hi*xh, lo*xh, hi*xl per k-step");
  run<2>(d, "4x hi*xh, 4x lo*xh, 4x hi*xl");
  run<3>(d, "same A, same B");
  run<4>(d, "kernel order, alternating accumulators");
  run<5>(d, "kernel order, aligned descriptors pitch 1024");
  return 0;
}
