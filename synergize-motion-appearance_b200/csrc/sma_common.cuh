// Shared device/host helpers for the sm_100a kernels behind include/sma_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/sma_b200.h"

extern std::atomic<int> g_sma_launches;   // defined in abi.cu; a counter only (never read by kernels)

#define SMA_LAUNCH_CHECK()                                  \
  do {                                                      \
    g_sma_launches.fetch_add(1, std::memory_order_relaxed); \
    if (cudaPeekAtLastError() != cudaSuccess) return SMA_ERR_CUDA; \
  } while (0)

static inline cudaStream_t as_stream(sma_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static constexpr int kNumSMs = 148;

__device__ __forceinline__ float sma_act(float v, int act) {
  switch (act) {
    case SMA_ACT_RELU: return fmaxf(v, 0.f);
    case SMA_ACT_LEAKY02: return v > 0.f ? v : 0.2f * v;
    case SMA_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case SMA_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case SMA_ACT_SWISH: return v / (1.f + expf(-v));
    default: return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
