// TF32 tensor-core dense layer for the fast mode (linear_tf32.cu).
#pragma once
#include "common.cuh"

namespace iisan {

// shapes / alignments the kernels accept (otherwise the caller stays on the fp32 FMA GEMM)
bool linear_tf32_supported(int rows, int N, int K, const float* x, int64_t ldx, const float* w, int64_t ld_other, const float* other);
int linear_tf32_forward(int rows, int N, int K, const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                        cudaStream_t st);
// dw / db are ACCUMULATED (callers pass zero-initialised buffers, like the fp32 path)
int linear_tf32_backward(int rows, int N, int K, const float* x, int64_t ldx, const float* w, const float* dy, int64_t lddy, float* dx,
                         int64_t lddx, float* dw, float* db, cudaStream_t st);

}  // namespace iisan
