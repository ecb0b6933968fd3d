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

// Per-device one-time state.  The >48 KB dynamic shared-memory opt-in (cudaFuncSetAttribute) and the SM count are properties of a
// (kernel, device) pair, so a process that drives several GPUs must set / query them once per device, not once per process.
// One bit (or slot) per device ordinal; idempotent and benign if raced (the attribute set is itself idempotent).
struct SmaDevOnce {
  std::atomic<unsigned long long> mask[4];     // 256 device ordinals
  SmaDevOnce() { for (auto& m : mask) m.store(0ull, std::memory_order_relaxed); }
};
template <class K>
static inline int sma_opt_in_smem(SmaDevOnce& once, K kernel, int bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SMA_ERR_CUDA;
  std::atomic<unsigned long long>& m = once.mask[(dev >> 6) & 3];
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(m.load(std::memory_order_acquire) & bit)) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return SMA_ERR_CUDA;
    m.fetch_or(bit, std::memory_order_release);
  }
  return SMA_OK;
}
// SM count of the current device (cached per device ordinal); <= 0 on error
static inline int sma_num_sms() {
  static std::atomic<int> cache[256];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  int v = cache[dev & 255].load(std::memory_order_relaxed);
  if (v > 0) return v;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return -1;
  cache[dev & 255].store(sms, std::memory_order_relaxed);
  return sms;
}

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
