// fp32 FMA GEMM building block for the exact (IISAN_COMPUTE_FP32) mode.
//   C[m,n] = epilogue( sum_k A(m,k) * B(k,n) ),  A/B addressed through (row, col) strides so the
//   same kernel serves forward (x W^T), data-gradient (dy W) and weight-gradient (dy^T x, split-K
//   with atomics) products.  Up to kMaxProbs independent problems (the three SAN towers) share one
//   launch through blockIdx.z.
#pragma once
#include "common.cuh"

namespace iisan {

constexpr int kMaxProbs = 4;

struct GemmProb {
  const float* A; int64_t a_rs, a_cs;   // A(m,k) = A[m*a_rs + k*a_cs]
  const float* B; int64_t b_rs, b_cs;   // B(k,n) = B[k*b_rs + n*b_cs]
  float* C; int64_t ldc;                // C[m*ldc + n]
  const float* bias;                    // [N] or null
  const float* resid; int64_t ldr;      // added after bias/relu/mask, or null
  const float* mask; int64_t ldm;       // multiply by (mask[m,n] > 0) -- or, mask_gelu, by gelu'(mask[m,n]) -- or null
  float* pre; int64_t ldp;              // relu == 2: the pre-activation (bias added) is stored here as well (GELU backward), or null
  int M, N, K;
  int relu;                             // 0 none, 1 ReLU, 2 exact (erf) GELU
  int mask_gelu;
  int splitk;                           // >1: K is split over blockIdx.y and C is atomically accumulated
  int accumulate;                       // 1 (with splitk==1): C += result
};

struct GemmBatch {
  GemmProb p[kMaxProbs];
  int n;
};

constexpr int GBM = 64, GBN = 64, GBK = 16;
int launch_gemm(const GemmBatch& b, cudaStream_t st);

// Convenience builders -----------------------------------------------------------------------------
// y[M,N] = act(x[M,K] W[N,K]^T + b) (+ resid)
inline GemmProb prob_linear(const float* x, int64_t ldx, const float* w, const float* b, float* y,
                            int64_t ldy, int M, int N, int K, int relu = 0,
                            const float* resid = nullptr, int64_t ldr = 0) {
  GemmProb p{};
  p.A = x; p.a_rs = ldx; p.a_cs = 1;
  p.B = w; p.b_rs = 1; p.b_cs = K;
  p.C = y; p.ldc = ldy; p.bias = b; p.resid = resid; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K; p.relu = relu; p.splitk = 1;
  return p;
}
// dx[M,K] = dy[M,N] W[N,K]  (* (mask>0)) (+ resid)
inline GemmProb prob_dgrad(const float* dy, int64_t lddy, const float* w, float* dx, int64_t lddx,
                           int M, int N, int K, const float* mask = nullptr, int64_t ldm = 0,
                           const float* resid = nullptr, int64_t ldr = 0, int mask_gelu = 0) {
  GemmProb p{};
  p.A = dy; p.a_rs = lddy; p.a_cs = 1;
  p.B = w; p.b_rs = K; p.b_cs = 1;
  p.C = dx; p.ldc = lddx; p.mask = mask; p.ldm = ldm; p.resid = resid; p.ldr = ldr; p.mask_gelu = mask_gelu;
  p.M = M; p.N = K; p.K = N; p.splitk = 1;
  return p;
}
// dw[N,K] += dy[M,N]^T x[M,K]   (reduction over the M rows, split-K + atomics)
inline GemmProb prob_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw,
                           int M, int N, int K) {
  GemmProb p{};
  p.A = dy; p.a_rs = 1; p.a_cs = lddy;      // A(n, m) = dy[m, n]
  p.B = x; p.b_rs = ldx; p.b_cs = 1;        // B(m, k) = x[m, k]
  p.C = dw; p.ldc = K;
  p.M = N; p.N = K; p.K = M;
  int tiles = ((N + GBM - 1) / GBM) * ((K + GBN - 1) / GBN);
  int want = (592 + tiles - 1) / tiles;                 // ~4 CTAs per SM in flight
  int kt = (M + GBK - 1) / GBK;
  p.splitk = want < 1 ? 1 : (want > kt ? kt : want);
  if (p.splitk > 64) p.splitk = 64;
  if (p.splitk == 1) p.accumulate = 1;
  return p;
}

// ---- column sums: out[n] += sum_m Y[m,n]  (bias gradients) --------------------------------------------
struct ColsumProb { const float* Y; int64_t ld; int M, N; float* out; const __nv_bfloat16* Yb; };   // Y (fp32) or Yb (bf16)
struct ColsumBatch { ColsumProb p[kMaxProbs * 2]; int n; };

int launch_colsum(const ColsumBatch& b, cudaStream_t st);

}  // namespace iisan
