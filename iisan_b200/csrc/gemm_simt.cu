// fp32 FMA GEMM + column-sum kernels (exact mode building blocks); see gemm_simt.cuh.
#include "gemm_simt.cuh"
#include "launch.cuh"

namespace iisan {


__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmBatch batch) {
  const GemmProb& P = batch.p[blockIdx.z];
  const int tiles_n = (P.N + GBN - 1) / GBN;
  const int tiles_m = (P.M + GBM - 1) / GBM;
  if ((int)blockIdx.x >= tiles_n * tiles_m) return;
  if ((int)blockIdx.y >= P.splitk) return;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int m0 = tm * GBM, n0 = tn * GBN;
  // K range of this split
  const int kt_total = (P.K + GBK - 1) / GBK;
  const int kt_per = (kt_total + P.splitk - 1) / P.splitk;
  const int kt_beg = blockIdx.y * kt_per;
  const int kt_end = min(kt_total, kt_beg + kt_per);
  if (kt_beg >= kt_end) return;

  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // thread computes rows ty*4.., cols tx*4..
  const bool a_kc = (P.a_cs == 1);          // k contiguous in memory
  const bool b_nc = (P.b_cs == 1);          // n contiguous in memory

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kt = kt_beg; kt < kt_end; ++kt) {
    const int k0 = kt * GBK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (a_kc) { k = e % GBK; m = e / GBK; } else { m = e % GBM; k = e / GBM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < P.M && gk < P.K) v = __ldg(P.A + (int64_t)gm * P.a_rs + (int64_t)gk * P.a_cs);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (b_nc) { n = e % GBN; k = e / GBN; } else { k = e % GBK; n = e / GBK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < P.N && gk < P.K) v = __ldg(P.B + (int64_t)gk * P.b_rs + (int64_t)gn * P.b_cs);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool first_split = (blockIdx.y == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= P.N) continue;
      float v = acc[i][j];
      if (P.splitk > 1) {
        if (first_split && P.bias) v += __ldg(P.bias + gn);
        atomicAdd(P.C + (int64_t)gm * P.ldc + gn, v);
      } else {
        if (P.bias) v += __ldg(P.bias + gn);
        if (P.relu) v = fmaxf(v, 0.f);
        if (P.mask) v = (__ldg(P.mask + (int64_t)gm * P.ldm + gn) > 0.f) ? v : 0.f;
        if (P.resid) v += __ldg(P.resid + (int64_t)gm * P.ldr + gn);
        float* c = P.C + (int64_t)gm * P.ldc + gn;
        if (P.accumulate) v += *c;
        *c = v;
      }
    }
  }
}

int launch_gemm(const GemmBatch& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  int max_tiles = 0, max_split = 1;
  for (int i = 0; i < b.n; ++i) {
    const GemmProb& P = b.p[i];
    if (P.M <= 0 || P.N <= 0 || P.K <= 0) return IISAN_EINVAL;
    int t = ((P.M + GBM - 1) / GBM) * ((P.N + GBN - 1) / GBN);
    if (t > max_tiles) max_tiles = t;
    if (P.splitk > max_split) max_split = P.splitk;
  }
  dim3 grid(max_tiles, max_split, b.n);
  { LaunchScope ls_(IISAN_K_GEMM, st); gemm_simt_kernel<<<grid, 256, 0, st>>>(b); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

__global__ void __launch_bounds__(256) colsum_kernel(const ColsumBatch batch, int rows_per_cta) {
  const ColsumProb& P = batch.p[blockIdx.z];
  const int col = blockIdx.x * 32 + (threadIdx.x % 32);
  const int r0 = blockIdx.y * rows_per_cta;
  if (r0 >= P.M) return;
  const int r1 = min(P.M, r0 + rows_per_cta);
  float s = 0.f;
  if (col < P.N)
  {
    if (P.Y) { for (int r = r0 + threadIdx.x / 32; r < r1; r += 8) s += __ldg(P.Y + (int64_t)r * P.ld + col); }
    else { for (int r = r0 + threadIdx.x / 32; r < r1; r += 8) s += __bfloat162float(P.Yb[(int64_t)r * P.ld + col]); }
  }
  __shared__ float red[8][33];
  red[threadIdx.x / 32][threadIdx.x % 32] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (col < P.N) atomicAdd(P.out + col, t);
  }
}

int launch_colsum(const ColsumBatch& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  int maxN = 0, maxM = 0;
  for (int i = 0; i < b.n; ++i) { maxN = max(maxN, b.p[i].N); maxM = max(maxM, b.p[i].M); }
  const int rows_per_cta = 256;
  dim3 grid((maxN + 31) / 32, (maxM + rows_per_cta - 1) / rows_per_cta, b.n);
  { LaunchScope ls_(IISAN_K_MISC, st); colsum_kernel<<<grid, 256, 0, st>>>(b, rows_per_cta); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan
