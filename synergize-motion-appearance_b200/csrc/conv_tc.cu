// tcgen05 3xTF32 implicit-GEMM convolution (placeholder until the tensor-core kernel lands).
#include "sma_common.cuh"
int sma_conv2d_tc_try(const sma_conv_desc* d, cudaStream_t st) { (void)d; (void)st; return SMA_ERR_UNSUPPORTED; }
