// Scratch micro-benchmark 2: what slows a stream of TS-mode tcgen05.mma (M128 x N128 x K16, kind::f16) down?
//   flag 1: per 12 MMAs, a commit to a barrier + a wait on an (already completed) other barrier + fences (the per-tap protocol)
//   flag 2: warps 4-7 run tcgen05.st.32x32b.x32 + wait::st in a loop into the A columns
//   flag 4: warps 8-11 run tcgen05.ld.32x32b.x32 + wait::ld in a loop on an accumulator
//   flag 8: warps 12-15 hammer shared memory with st.shared / ld.shared
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../synergize-motion-appearance_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(512, 1) rate_kernel(int flags, int iters, long long* out, int sbo, int start_off) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 512) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  fence_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t db = (uint64_t)(((sbase + 4096 + start_off) & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
    if (elect_one_sync()) tc_commit(smem_u32(&bar[1]));      // bar[1]: completed once, waited on with parity 0 forever
    __syncwarp();
    mbar_wait(smem_u32(&bar[1]), 0);
    long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (flags & 1) { mbar_wait(smem_u32(&bar[1]), 0); tc_fence_after(); }
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 12; k++) {
          const uint64_t ko = (uint64_t)((k & 3) * 2);
          mma_ts(tm + (uint32_t)((it & 1) * 128), tm + 256u + (uint32_t)((it & 3) * 64 + (k & 3) * 8), db + ko, idesc, 1u);
        }
        if (flags & 1) { tc_commit(smem_u32(&bar[2])); tc_commit(smem_u32(&bar[3])); }
      }
      __syncwarp();
    }
    if (elect_one_sync()) tc_commit(smem_u32(&bar[0]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[0]), 0);
    long long t1 = clock64();
    long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    stop = 1;
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[13] = g1 - g0; }
  } else if (warp >= 4 && warp < 8 && (flags & 2)) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = 0x3c003c00u;
    int n = 0;
    while (!stop) {
      tmem_st32(tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)((n & 7) * 32), r);
      tmem_st_wait();
      n++;
    }
    if (lane == 0 && blockIdx.x == 0) out[1 + (warp & 3)] = n;
  } else if (warp >= 8 && warp < 12 && (flags & 4)) {
    uint32_t r[32]; uint32_t acc = 0; int n = 0;
    while (!stop) {
      tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((n & 3) * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) acc += r[i];
      n++;
    }
    if (lane == 0 && blockIdx.x == 0) out[5 + (warp & 3)] = n + (acc == 12345u);
  } else if (warp >= 12 && (flags & 8)) {
    const uint32_t base = sbase + 32768u + (uint32_t)(warp - 12) * 4096u;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f); int n = 0;
    while (!stop) {
#pragma unroll
      for (int i = 0; i < 8; i++) sts128(base + (uint32_t)((i * 32 + lane) * 16) % 4096u, v.x, v.y, v.z, v.w);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(base + (uint32_t)((i * 32 + lane) * 16) % 4096u));
        v.x += t.y;
      }
      n++;
    }
    if (lane == 0 && blockIdx.x == 0) out[9 + (warp - 12)] = n + (v.x == 1.2345f);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int iters = 2000;
  const int cfg[][3] = {{0, 1024, 0}, {0, 1024, 1}, {0, 1024, 2}, {15, 1024, 2}, {0, 1024, 0}};
  for (auto& c3 : cfg) {
    int flags = c3[0]; iters = c3[2] == 0 ? 2000 : (c3[2] == 1 ? 100000 : 400000);
    cudaMemset(d, 0, 16 * 8);
    printf("sbo %d start_off %d : ", c3[1], c3[2]);
    rate_kernel<<<148, 512, 64 * 1024>>>(flags, iters, d, c3[1], 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[16]; cudaMemcpy(c, d, 16 * 8, cudaMemcpyDeviceToHost);
    printf("flags %2d [%s%s%s%s]: %.1f cycles/MMA | st/warp %lld (%.0f cyc each) ld/warp %lld (%.0f cyc each) smem iters %lld (%s) | %.1f ms, SM clock %.0f MHz, %.0f TFLOP/s chip\n", flags,
           flags & 1 ? "tap-protocol " : "", flags & 2 ? "tmem-st " : "", flags & 4 ? "tmem-ld " : "", flags & 8 ? "smem " : "",
           (double)c[0] / (iters * 12.0), c[1], c[1] ? (double)c[0] / c[1] : 0.0, c[5], c[5] ? (double)c[0] / c[5] : 0.0, c[9], cudaGetErrorString(e), c[13] * 1e-6, (double)c[0] / c[13] * 1e3, 148.0 * iters * 12 * 2.0 * 128 * 128 * 16 / (c[13] * 1e-9) / 1e12);
  }
  return 0;
}
