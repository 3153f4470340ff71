// Dense layer (com_dense: Linear 3E -> E, CC/model/model.py:37-38,72) on the tensor cores for the fast mode: warp-level
// mma.sync.m16n8k8 TF32 tiles straight from the fp32 activations / nn.Linear weight in global memory (no operand copies:
// 5632 x 192 x 64 is far too small to amortise a cast + TMA pipeline).  Fragment coordinates as in user_encoder_fused.cu:
// a thread loads four consecutive k with one 128-bit load and feeds slots (t, t+4) of two MMAs with (4t, 4t+1), (4t+2, 4t+3).
//   forward   y[M,N]  = x[M,K] W[N,K]^T + b
//   dgrad     dx[M,K] = dy[M,N] W[N,K]
//   wgrad     dW[N,K] += dy^T x ; db[N] += colsum(dy)      (row blocks of 64 staged in shared memory, red.add)
#include "common.cuh"
#include "launch.cuh"
#include "linear_tf32.cuh"

namespace iisan {

__device__ __forceinline__ uint32_t lt_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void lt_mma(float (&c)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(lt_tf32(a0)), "r"(lt_tf32(a1)), "r"(lt_tf32(a2)), "r"(lt_tf32(a3)), "r"(lt_tf32(b0)), "r"(lt_tf32(b1)));
}

constexpr int LT_WARPS = 4;        // 16 rows each
constexpr int LT_NT = 2;           // 8-column tiles per warp: small tiles, many warps (the products are latency bound, not FLOP bound)

// grid (ceil(M / 64), ceil(N / 16)), 128 threads
__global__ void __launch_bounds__(LT_WARPS * 32) lin_tf32_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, float* __restrict__ y, int64_t ldy,
                                                                      int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = blockIdx.x * (LT_WARPS * 16) + warp * 16, n0 = blockIdx.y * (LT_NT * 8);
  if (r0 >= M) return;
  float acc[LT_NT][4];
#pragma unroll
  for (int j = 0; j < LT_NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
  const bool ra = r0 + g < M, rb = r0 + g + 8 < M;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 16) {
    const float4 xa = ra ? *reinterpret_cast<const float4*>(x + (int64_t)(r0 + g) * ldx + k0 + 4 * t) : z4;
    const float4 xb = rb ? *reinterpret_cast<const float4*>(x + (int64_t)(r0 + g + 8) * ldx + k0 + 4 * t) : z4;
#pragma unroll
    for (int j = 0; j < LT_NT; ++j) {
      const int col = n0 + 8 * j + g;
      const float4 wv = col < N ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)col * K + k0 + 4 * t)) : z4;
      lt_mma(acc[j], xa.x, xb.x, xa.y, xb.y, wv.x, wv.y);
      lt_mma(acc[j], xa.z, xb.z, xa.w, xb.w, wv.z, wv.w);
    }
  }
#pragma unroll
  for (int j = 0; j < LT_NT; ++j) {
    const int col = n0 + 8 * j + 2 * t;
    if (col >= N) continue;
    const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
    if (ra) *reinterpret_cast<float2*>(y + (int64_t)(r0 + g) * ldy + col) = make_float2(acc[j][0] + b0, acc[j][1] + b1);
    if (rb) *reinterpret_cast<float2*>(y + (int64_t)(r0 + g + 8) * ldy + col) = make_float2(acc[j][2] + b0, acc[j][3] + b1);
  }
}

// dx[M,K] = dy[M,N] W[N,K]: grid (ceil(M / 64), ceil(K / 16)), contraction over N in steps of 16
__global__ void __launch_bounds__(LT_WARPS * 32) lin_tf32_dgrad_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ w,
                                                                        float* __restrict__ dx, int64_t lddx, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = blockIdx.x * (LT_WARPS * 16) + warp * 16, c0 = blockIdx.y * (LT_NT * 8);
  if (r0 >= M) return;
  float acc[LT_NT][4];
#pragma unroll
  for (int j = 0; j < LT_NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
  const bool ra = r0 + g < M, rb = r0 + g + 8 < M;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int n0 = 0; n0 < N; n0 += 16) {
    const float4 da = ra ? *reinterpret_cast<const float4*>(dy + (int64_t)(r0 + g) * lddy + n0 + 4 * t) : z4;
    const float4 db = rb ? *reinterpret_cast<const float4*>(dy + (int64_t)(r0 + g + 8) * lddy + n0 + 4 * t) : z4;
    const float* wp = w + (int64_t)(n0 + 4 * t) * K;
#pragma unroll
    for (int j = 0; j < LT_NT; ++j) {
      const int col = c0 + 8 * j + g;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
      if (col < K) { w0 = __ldg(wp + col); w1 = __ldg(wp + K + col); w2 = __ldg(wp + 2 * (int64_t)K + col); w3 = __ldg(wp + 3 * (int64_t)K + col); }
      lt_mma(acc[j], da.x, db.x, da.y, db.y, w0, w1);
      lt_mma(acc[j], da.z, db.z, da.w, db.w, w2, w3);
    }
  }
#pragma unroll
  for (int j = 0; j < LT_NT; ++j) {
    const int col = c0 + 8 * j + 2 * t;
    if (col >= K) continue;
    if (ra) *reinterpret_cast<float2*>(dx + (int64_t)(r0 + g) * lddx + col) = make_float2(acc[j][0], acc[j][1]);
    if (rb) *reinterpret_cast<float2*>(dx + (int64_t)(r0 + g + 8) * lddx + col) = make_float2(acc[j][2], acc[j][3]);
  }
}

// dW[N,K] += dy^T x and db[N] += colsum(dy) over a block of LT_RB rows staged in shared memory (row strides = 8 mod 32 banks:
// the transposed fragment reads of both operands are conflict-free); 8 warps walk the (N/16) x (K/8) output tiles
constexpr int LT_RB = 64;
constexpr int LT_WG_THREADS = 256;
__global__ void __launch_bounds__(LT_WG_THREADS) lin_tf32_wgrad_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x,
                                                                        int64_t ldx, float* __restrict__ dw, float* __restrict__ dbias, int M,
                                                                        int N, int K) {
  extern __shared__ __align__(16) float lsm[];
  const int SY = N + 8, SX = K + 8;
  float* sy = lsm;                    // [LT_RB][SY]
  float* sx = lsm + LT_RB * SY;       // [LT_RB][SX]
  const int r0 = blockIdx.x * LT_RB;
  for (int idx = threadIdx.x; idx < LT_RB * (N / 4); idx += LT_WG_THREADS) {
    const int r = idx / (N / 4), c4 = idx % (N / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < M) v = *reinterpret_cast<const float4*>(dy + (int64_t)(r0 + r) * lddy + 4 * c4);
    *reinterpret_cast<float4*>(sy + r * SY + 4 * c4) = v;
  }
  for (int idx = threadIdx.x; idx < LT_RB * (K / 4); idx += LT_WG_THREADS) {
    const int r = idx / (K / 4), c4 = idx % (K / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < M) v = *reinterpret_cast<const float4*>(x + (int64_t)(r0 + r) * ldx + 4 * c4);
    *reinterpret_cast<float4*>(sx + r * SX + 4 * c4) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int MT = N / 16, NT = K / 8;
  for (int tile = warp; tile < MT * NT; tile += LT_WG_THREADS / 32) {
    const int mt = tile / NT, nt = tile % NT;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < LT_RB / 8; ++ks) {
      const float* dp = sy + (8 * ks + t) * SY + mt * 16 + g;
      const float* xp = sx + (8 * ks + t) * SX + nt * 8 + g;
      lt_mma(acc, dp[0], dp[8], dp[4 * SY], dp[4 * SY + 8], xp[0], xp[4 * SX]);
    }
    float* o = dw + (int64_t)(mt * 16 + g) * K + nt * 8 + 2 * t;
    atomicAdd(reinterpret_cast<float2*>(o), make_float2(acc[0], acc[1]));
    atomicAdd(reinterpret_cast<float2*>(o + (int64_t)8 * K), make_float2(acc[2], acc[3]));
  }
  if (dbias) {
    for (int n = threadIdx.x; n < N; n += LT_WG_THREADS) {
      float s = 0.f;
      for (int r = 0; r < LT_RB; ++r) s += sy[r * SY + n];
      atomicAdd(dbias + n, s);
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool linear_tf32_supported(int rows, int N, int K, const float* x, int64_t ldx, const float* w, int64_t ld_other, const float* other) {
  return rows > 0 && N % 16 == 0 && K % 16 == 0 && N <= 256 && K <= 512 && ldx % 4 == 0 && ld_other % 4 == 0 && aligned16(x) && aligned16(w) &&
         aligned16(other);
}

int linear_tf32_forward(int rows, int N, int K, const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                        cudaStream_t st) {
  dim3 grid((rows + LT_WARPS * 16 - 1) / (LT_WARPS * 16), (N + LT_NT * 8 - 1) / (LT_NT * 8));
  { LaunchScope ls_(IISAN_K_GEMM, st); lin_tf32_fwd_kernel<<<grid, LT_WARPS * 32, 0, st>>>(x, ldx, w, b, y, ldy, rows, N, K); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int linear_tf32_backward(int rows, int N, int K, const float* x, int64_t ldx, const float* w, const float* dy, int64_t lddy, float* dx,
                         int64_t lddx, float* dw, float* db, cudaStream_t st) {
  if (dx) {
    dim3 grid((rows + LT_WARPS * 16 - 1) / (LT_WARPS * 16), (K + LT_NT * 8 - 1) / (LT_NT * 8));
    { LaunchScope ls_(IISAN_K_GEMM, st); lin_tf32_dgrad_kernel<<<grid, LT_WARPS * 32, 0, st>>>(dy, lddy, w, dx, lddx, rows, N, K); }
    IISAN_LAUNCH_OK();
  }
  if (dw || db) {
    if (!dw) return IISAN_EINVAL;
    const size_t smem = (size_t)LT_RB * (N + 8 + K + 8) * sizeof(float);
    static size_t attr = 0;
    if (smem > attr) {
      IISAN_CUDA_OK(cudaFuncSetAttribute(lin_tf32_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    { LaunchScope ls_(IISAN_K_GEMM, st); lin_tf32_wgrad_kernel<<<(rows + LT_RB - 1) / LT_RB, LT_WG_THREADS, smem, st>>>(dy, lddy, x, ldx, dw, db, rows, N, K); }
    IISAN_LAUNCH_OK();
  }
  return IISAN_OK;
}

}  // namespace iisan
