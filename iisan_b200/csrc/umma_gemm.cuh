// bf16 tensor-core GEMM building block (tcgen05 + TMEM + TMA) for the fast (IISAN_COMPUTE_BF16) mode.
//   D[M,N] = epilogue( A x B ),  fp32 accumulation in TMEM.
// Operands are bf16 matrices in global memory described by (pointer, rows, cols, row pitch):
//   A "K-major"  : stored [M, K] row-major   (reduction contiguous)  -- activations in forward / dgrad
//   A "MN-major" : stored [K, M] row-major   (M contiguous)          -- dy in the weight gradient dy^T x
//   B "K-major"  : stored [N, K] row-major   -- nn.Linear weight [out, in] in forward, transposed copy in dgrad
//   B "MN-major" : stored [K, N] row-major   -- x in the weight gradient
// Up to kUmmaMaxProbs problems (the three SAN towers) share one launch through blockIdx.z.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace iisan {

constexpr int kUmmaMaxProbs = 3;

struct UmmaOperand {
  const __nv_bfloat16* ptr;
  int64_t rows, cols;     // as stored (row-major)
  int64_t pitch;          // elements between rows
};

struct UmmaEpilogue {
  float* out_f32; int64_t ld_f32;                 // optional fp32 output
  __nv_bfloat16* out_bf16; int64_t ld_bf16;       // optional bf16 output
  const float* bias;                              // [N] or null
  const __nv_bfloat16* mask; int64_t ld_mask;     // multiply by (mask > 0) or null   (ReLU backward)
  const __nv_bfloat16* resid_bf16; int64_t ld_resid_bf16;   // added last, or null
  const float* resid_f32; int64_t ld_resid_f32;
  int relu;
  int gelu;                                       // 1: out_bf16 = gelu_erf(acc + bias), out_f32 = the pre-activation acc + bias (AdapterBlock with GELU)
  const float* gelu_pre; int64_t ld_gelu_pre;     // multiply by gelu_erf'(gelu_pre) or null   (GELU backward)
  int atomic;                                     // 1: red.add into out_f32 (split-K partial sums)
  int transpose_out;                              // 1: out_f32 is addressed [n * ld_f32 + m] (fp32 output only)
};

struct UmmaProblem {
  UmmaOperand A, B;
  int a_mn_major, b_mn_major;
  int M, N, K;
  int splitk;
  UmmaEpilogue epi;
};

struct UmmaBatch {
  UmmaProblem p[kUmmaMaxProbs];
  int n;
};

// Enqueue the batch on `st`.  Shapes: N % 16 == 0, K % 8 == 0 (16-byte global pitch), pointers 16-byte aligned.
int launch_umma_gemm(const UmmaBatch& batch, cudaStream_t st);

// Many independent weight-gradient products (MN-major operands, N <= 64) in ONE launch: the 42 dW GEMMs of the chain backward.
// (The argument block is ~18 KB; kernel parameters up to 32 KB are supported by CUDA 12.1+ on sm_70+.)
constexpr int kUmmaBigProbs = 48;
struct UmmaBatchBig {
  UmmaProblem p[kUmmaBigProbs];
  int n;
};
int launch_umma_gemm_big(const UmmaBatchBig& batch, cudaStream_t st);
// Many problems of ONE operand-major combination (K/K, K/MN or MN/MN) in one launch; 64-wide column tiles when every N <= 64,
// else 256-wide.  The low-rank adjoint path (san_lr.cu): weight products, G = dz^T h, Gram blocks, combine GEMMs.
int launch_umma_gemm_many(const UmmaBatchBig& batch, cudaStream_t st);

}  // namespace iisan
