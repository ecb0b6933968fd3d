// Scratch micro-benchmark: issue rate of tcgen05.mma kind::f16 (cta_group::1, M = 128) for different N, operand sources and
// accumulator dependence.  One CTA per SM; one elected thread issues `iters` x `per` MMAs back to back on garbage operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../synergize-motion-appearance_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS same accumulator; 1: SS alternating 2 accumulators; 2: TS same accumulator; 3: TS alternating accumulators
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0x3c003c00u;   // fp16 1.0
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = make_desc(sbase), db = make_desc(sbase + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 12; k++) {
          const uint64_t ko = (uint64_t)((k & 3) * 2);
          const uint32_t d = tm + ((mode & 1) ? (uint32_t)((k & 1) * 256) : 0u);
          if (mode < 2) tc_mma_f16(d, da + ko, db + ko, idesc, 1u);
          else mma_ts(d, tm + 480u + (uint32_t)((k & 3) * 8), db + ko, idesc, 1u);
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

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; mode++)
      for (int N : {32, 64, 128, 256}) {
        if ((mode & 1) && N > 224) { }
        rate_kernel<<<grid, 128, 64 * 1024>>>(N, mode, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d mode %d (%s, %s) N %3d: %.1f cycles/MMA  (%s)\n", grid, mode, mode < 2 ? "SS" : "TS", (mode & 1) ? "2 acc" : "1 acc", N,
               (double)c / (iters * 12.0), cudaGetErrorString(e));
      }
  return 0;
}
