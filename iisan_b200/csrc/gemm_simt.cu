// fp32 FMA GEMM + column-sum kernels (exact mode building blocks); see gemm_simt.cuh.
#include "gemm_simt.cuh"
#include "launch.cuh"

namespace iisan {


// BM x 64 output tile per CTA (BM = 64 or 32; the narrow tile doubles the CTA count of the skinny SASRec products), 256 threads,
// each thread (BM/16) x 4 outputs; the next k-tile's global loads are issued into registers before the current tile's FMAs.
template <int BM>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmBatch batch) {
  constexpr int TM = BM / 16;                 // rows per thread
  constexpr int AL = BM * GBK / 256;          // A elements loaded per thread per k-tile
  const GemmProb& P = batch.p[blockIdx.z];
  const int tiles_n = (P.N + GBN - 1) / GBN;
  const int tiles_m = (P.M + BM - 1) / BM;
  if ((int)blockIdx.x >= tiles_n * tiles_m) return;
  if ((int)blockIdx.y >= P.splitk) return;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int m0 = tm * BM, n0 = tn * GBN;
  const int kt_total = (P.K + GBK - 1) / GBK;
  const int kt_per = (kt_total + P.splitk - 1) / P.splitk;
  const int kt_beg = blockIdx.y * kt_per;
  const int kt_end = min(kt_total, kt_beg + kt_per);
  if (kt_beg >= kt_end) return;

  __shared__ float As[GBK][BM + 4];
  __shared__ float Bs[GBK][GBN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // thread computes rows ty*TM.., cols tx*4..
  const bool a_kc = (P.a_cs == 1);          // k contiguous in memory
  const bool b_nc = (P.b_cs == 1);          // n contiguous in memory

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[AL], rb[4];
  auto fetch = [&](int kt) {
    const int k0 = kt * GBK;
#pragma unroll
    for (int i = 0; i < AL; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (a_kc) { k = e % GBK; m = e / GBK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < P.M && gk < P.K) ? __ldg(P.A + (int64_t)gm * P.a_rs + (int64_t)gk * P.a_cs) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (b_nc) { n = e % GBN; k = e / GBN; } else { k = e % GBK; n = e / GBK; }
      const int gn = n0 + n, gk = k0 + k;
      rb[i] = (gn < P.N && gk < P.K) ? __ldg(P.B + (int64_t)gk * P.b_rs + (int64_t)gn * P.b_cs) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < AL; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (a_kc) { k = e % GBK; m = e / GBK; } else { m = e % BM; k = e / BM; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (b_nc) { n = e % GBN; k = e / GBN; } else { k = e % GBK; n = e / GBK; }
      Bs[k][n] = rb[i];
    }
  };

  fetch(kt_beg);
  for (int kt = kt_beg; kt < kt_end; ++kt) {
    stage();
    __syncthreads();
    if (kt + 1 < kt_end) fetch(kt + 1);      // in flight while this tile is multiplied
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float av[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[k][ty * TM + i];
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool first_split = (blockIdx.y == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
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
        if (P.relu == 1) v = fmaxf(v, 0.f);
        else if (P.relu == 2) { if (P.pre) P.pre[(int64_t)gm * P.ldp + gn] = v; v = gelu_erf(v); }
        if (P.mask) {
          const float mk = __ldg(P.mask + (int64_t)gm * P.ldm + gn);
          v = P.mask_gelu ? v * gelu_erf_grad(mk) : (mk > 0.f ? v : 0.f);
        }
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
  int max_tiles = 0, max_tiles32 = 0, max_split = 1;
  int64_t ctas = 0;
  for (int i = 0; i < b.n; ++i) {
    const GemmProb& P = b.p[i];
    if (P.M <= 0 || P.N <= 0 || P.K <= 0) return IISAN_EINVAL;
    const int t = ((P.M + GBM - 1) / GBM) * ((P.N + GBN - 1) / GBN);
    const int t32 = ((P.M + 31) / 32) * ((P.N + GBN - 1) / GBN);
    if (t > max_tiles) max_tiles = t;
    if (t32 > max_tiles32) max_tiles32 = t32;
    if (P.splitk > max_split) max_split = P.splitk;
    ctas += (int64_t)t * (P.splitk < 1 ? 1 : P.splitk);
  }
  if (ctas < 2 * 148) {      // short grid: 32-row tiles
    dim3 grid(max_tiles32, max_split, b.n);
    { LaunchScope ls_(IISAN_K_GEMM, st); gemm_simt_kernel<32><<<grid, 256, 0, st>>>(b); }
  } else {
    dim3 grid(max_tiles, max_split, b.n);
    { LaunchScope ls_(IISAN_K_GEMM, st); gemm_simt_kernel<64><<<grid, 256, 0, st>>>(b); }
  }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

// out[n] += sum_m Y[m, n]: a thread owns 8 consecutive columns (one 128-bit load per row for bf16, two for fp32), a warp 256
// columns, the 8 warps of a CTA take alternate rows of a 256-row block; partial sums meet in shared memory, one red.add per
// column and CTA.
__global__ void __launch_bounds__(256) colsum_kernel(const ColsumBatch batch, int rows_per_cta) {
  const ColsumProb& P = batch.p[blockIdx.z];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  if (r0 >= P.M || blockIdx.x * 256 >= P.N) return;
  const int r1 = min(P.M, r0 + rows_per_cta);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vec = (col + 8 <= P.N) && (P.ld % 8 == 0) &&
                   ((reinterpret_cast<uintptr_t>(P.Y ? (const void*)P.Y : (const void*)P.Yb) & 15) == 0);
  if (col < P.N) {
    if (vec && !P.Y) {
      for (int r = r0 + warp; r < r1; r += 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(P.Yb + (int64_t)r * P.ld + col));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); s[2 * i] += f.x; s[2 * i + 1] += f.y; }
      }
    } else if (vec) {
      for (int r = r0 + warp; r < r1; r += 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.Y + (int64_t)r * P.ld + col));
        const float4 c = __ldg(reinterpret_cast<const float4*>(P.Y + (int64_t)r * P.ld + col) + 1);
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += c.x; s[5] += c.y; s[6] += c.z; s[7] += c.w;
      }
    } else {
      for (int r = r0 + warp; r < r1; r += 8)
        for (int i = 0; i < 8; ++i)
          if (col + i < P.N) s[i] += P.Y ? __ldg(P.Y + (int64_t)r * P.ld + col + i) : __bfloat162float(P.Yb[(int64_t)r * P.ld + col + i]);
    }
  }
  __shared__ float red[8][257];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = s[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < P.N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(P.out + c, t);
  }
}

int launch_colsum(const ColsumBatch& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  int maxN = 0, maxM = 0;
  for (int i = 0; i < b.n; ++i) { maxN = max(maxN, b.p[i].N); maxM = max(maxM, b.p[i].M); }
  const int rows_per_cta = 128;
  dim3 grid((maxN + 255) / 256, (maxM + rows_per_cta - 1) / rows_per_cta, b.n);
  { LaunchScope ls_(IISAN_K_MISC, st); colsum_kernel<<<grid, 256, 0, st>>>(b, rows_per_cta); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan
