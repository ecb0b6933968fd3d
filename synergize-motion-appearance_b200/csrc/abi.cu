// Library-level entry points of include/sma_b200.h.
#include "sma_common.cuh"

std::atomic<int> g_sma_launches{0};

extern "C" int sma_abi_version(void) { return 11; }

extern "C" const char* sma_status_string(int s) {
  switch (s) {
    case SMA_OK: return "ok";
    case SMA_ERR_BAD_ARG: return "bad argument";
    case SMA_ERR_UNSUPPORTED: return "unsupported shape";
    case SMA_ERR_CUDA: return "CUDA launch error";
    case SMA_ERR_NO_DEVICE: return "device is not sm_100 (B200)";
    default: return "unknown status";
  }
}

extern "C" int sma_device_check(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return SMA_ERR_CUDA;
  return p.major == 10 ? SMA_OK : SMA_ERR_NO_DEVICE;
}

extern "C" int sma_sizeof_conv_desc(void) { return (int)sizeof(sma_conv_desc); }

extern "C" int sma_kernel_launch_count(void) { return g_sma_launches.load(std::memory_order_relaxed); }
