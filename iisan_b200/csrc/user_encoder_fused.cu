// SASRec user encoder, whole encoder per CTA (E = 64, L = 10; exact mode: fp32 FMA, fast mode: TF32 tensor-core linears).
//
// Algorithm restated from the reference's PyTorch modules (nothing ported; same operator order as user_encoder.cu):
//   entry                 CC/model/modules.py:89-96     x = dropout(LN(embs + pos_emb))
//   attention block       CC/model/modules.py:21-32, 54-64 ; mask CC/model/encoders.py:53-58
//   feed-forward block    CC/model/modules.py:14-18
//
// The per-operator path (user_encoder.cu) launches ~45 kernels for 512 users x 10 positions x 64 features: pure launch and
// latency overhead.  Here one CTA owns 4 users (40 rows): the activations of the whole encoder stay in shared memory, the
// weights (100 k parameters, 400 KB) are read through L1/L2 by every CTA, and each linear layer is either a register-tiled FMA
// loop (exact mode: one weight row per thread against the CTA's rows, broadcast 128-bit shared-memory reads) or a set of
// warp-level mma.sync TF32 tiles (fast mode).  The forward writes the same
// stash as the per-operator path; the backward recomputes nothing, accumulates the weight gradients of its 40 rows in
// registers and adds them to the global gradient with one red.add per element and CTA.
#include "common.cuh"
#include "launch.cuh"
#include "user_encoder.cuh"

namespace iisan {

constexpr int FE = 64;            // embedding width
constexpr int FL = 10;            // sequence length
constexpr int FUPC = 4;           // users per CTA
constexpr int FR = FUPC * FL;     // rows per CTA
constexpr int FF = 4 * FE;        // FFN width
constexpr int FTHREADS = 512;
constexpr int FMAXH = 4;

struct FuArgs {
  int users, H, n_blocks;
  const float* embs; int64_t ld_user;
  const float* log_mask;
  iisan_ue_params P;
  iisan_ue_params G;              // backward: gradient pointers
  UeLayout W;
  const float* d_out; float* d_embs;
  float* out;
  DropCfg dc;
};

// TC = false: fp32 FMA linears (exact mode).  TC = true (fast mode): the linears run on the tensor cores as warp-level
// mma.sync.m16n8k8 TF32 tiles (fp32 operands straight from shared memory / the fp32 weights, fp32 accumulation): the rows of a
// CTA are padded to 3 MMA row tiles and the row strides to a multiple of 32 banks + 16 so that the 128-bit fragment loads of
// the 8 rows x 4 k-quads of a quarter-warp fall into distinct banks.
template <bool TC>
struct FuSmem {
  static constexpr int RP = TC ? 48 : FR;            // rows allocated per buffer
  static constexpr int SX = TC ? FE + 16 : FE;       // row stride of the [., FE] buffers
  static constexpr int SH = TC ? FF + 16 : FF;       // row stride of the [., FF] buffers
  static constexpr int kX = 0;                       // [FR][FE] block input / running activation
  static constexpr int kQ = kX + RP * SX;            // [FR][FE]
  static constexpr int kK = kQ + RP * SX;
  static constexpr int kV = kK + RP * SX;
  static constexpr int kC = kV + RP * SX;            // [FR][FE] ctx / scratch
  static constexpr int kM = kC + RP * SX;            // [FR][FE] xmid
  static constexpr int kP = kM + RP * SX;            // [FUPC][FMAXH][FL][FL]
  static constexpr int kH = kP + FUPC * FMAXH * FL * FL;   // [FR][FF]
  static constexpr int kD = kH + RP * SH;            // backward only: second [FR][FF] buffer
  static constexpr int kFwdFloats = kD;
  static constexpr int kBwdFloats = kD + RP * SH;
};
template <bool TC, int W>
__host__ __device__ constexpr int fu_stride() { return TC ? W + 16 : W; }

// ---------------------------------------------------------------------------------------------------------------
// fp32 FMA linears (exact mode)
// ---------------------------------------------------------------------------------------------------------------
// ys[r][o] = act(bias[o] + sum_k xs[r][k] W[o][k]) for the CTA's FR rows.  One output feature per thread, RT rows per thread.
template <int IN, int OUT>
__device__ __forceinline__ void fma_linear(const float* __restrict__ Wg, const float* __restrict__ bias, const float* xs, float* ys,
                                           bool relu) {
  constexpr int GROUPS = FTHREADS / OUT;     // row groups
  constexpr int RT = FR / GROUPS;            // rows per thread
  static_assert(FTHREADS % OUT == 0 && FR % GROUPS == 0, "tiling");
  const int o = threadIdx.x % OUT, rg = threadIdx.x / OUT;
  float acc[RT];
#pragma unroll
  for (int i = 0; i < RT; ++i) acc[i] = 0.f;
  const float4* wrow = reinterpret_cast<const float4*>(Wg + (size_t)o * IN);
#pragma unroll 4
  for (int k4 = 0; k4 < IN / 4; ++k4) {
    const float4 w = __ldg(wrow + k4);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xs + (rg * RT + i) * IN + k4 * 4);
      acc[i] = fmaf(x.w, w.w, fmaf(x.z, w.z, fmaf(x.y, w.y, fmaf(x.x, w.x, acc[i]))));
    }
  }
  const float b = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    float v = acc[i] + b;
    if (relu) v = fmaxf(v, 0.f);
    ys[(rg * RT + i) * OUT + o] = v;
  }
}

// dxs[r][i] (+)= sum_o dys[r][o] W[o][i]   (data gradient of a linear layer).  One input feature per thread.
template <int IN, int OUT>
__device__ __forceinline__ void fma_linear_t(const float* __restrict__ Wg, const float* dys, float* dxs, bool accumulate) {
  constexpr int GROUPS = FTHREADS / IN;
  constexpr int RT = FR / GROUPS;
  static_assert(FTHREADS % IN == 0 && FR % GROUPS == 0, "tiling");
  const int i = threadIdx.x % IN, rg = threadIdx.x / IN;
  float acc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
#pragma unroll 4
  for (int o4 = 0; o4 < OUT / 4; ++o4) {
    const float w0 = __ldg(Wg + (size_t)(o4 * 4 + 0) * IN + i), w1 = __ldg(Wg + (size_t)(o4 * 4 + 1) * IN + i);
    const float w2 = __ldg(Wg + (size_t)(o4 * 4 + 2) * IN + i), w3 = __ldg(Wg + (size_t)(o4 * 4 + 3) * IN + i);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float4 d = *reinterpret_cast<const float4*>(dys + (rg * RT + r) * OUT + o4 * 4);
      acc[r] = fmaf(d.w, w3, fmaf(d.z, w2, fmaf(d.y, w1, fmaf(d.x, w0, acc[r]))));
    }
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    float* p = dxs + (rg * RT + r) * IN + i;
    *p = accumulate ? *p + acc[r] : acc[r];
  }
}

// dW[o][i] += sum_r dys[r][o] xs[r][i] over the CTA's rows; 4x4 register tiles, one red.add per element.
template <int IN, int OUT>
__device__ __forceinline__ void fma_wgrad(float* __restrict__ dWg, const float* dys, const float* xs) {
  constexpr int TI = IN / 4, TILES = (OUT / 4) * TI;
  for (int t = threadIdx.x; t < TILES; t += FTHREADS) {
    const int o0 = (t / TI) * 4, i0 = (t % TI) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int r = 0; r < FR; ++r) {
      const float4 d = *reinterpret_cast<const float4*>(dys + r * OUT + o0);
      const float4 x = *reinterpret_cast<const float4*>(xs + r * IN + i0);
      const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dv[a], xv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) atomicAdd(dWg + (size_t)(o0 + a) * IN + i0 + b, acc[a][b]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// tensor-core linears (fast mode): mma.sync.m16n8k8 TF32, fp32 accumulate.  Fragment coordinates (g = lane / 4, t = lane % 4):
//   A (16 x 8): a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)       B (8 x 8): b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   C (16 x 8): c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
// A contraction is a sum, so the k slots of one MMA may hold any 8 distinct k as long as A and B agree: a thread loads FOUR
// consecutive k (one 128-bit load) and feeds slots (t, t + 4) of two MMAs with (4t, 4t + 1) and (4t + 2, 4t + 3).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FNW = FTHREADS / 32;           // warps per CTA
constexpr int FMT = 3;                       // MMA row tiles per CTA (48 >= FR rows)
static_assert(FR <= FMT * 16 && FR % 8 == 0, "row tiling of the tensor-core path");

// round-to-nearest TF32 (10 explicit mantissa bits: the operand precision of the reference's fp16 autocast GEMMs)
__device__ __forceinline__ uint32_t tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], float a0, float a1, float a2, float a3, float b0, float b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(tf32(a0)), "r"(tf32(a1)), "r"(tf32(a2)), "r"(tf32(a3)), "r"(tf32(b0)), "r"(tf32(b1)));
}

// rows [m_first*16, (m_first+MTW)*16) x columns [nt*8, +8) of ys = act(xs W^T + bias)
template <int IN, int OUT, int MTW>
__device__ __forceinline__ void tc_linear_tile(const float* __restrict__ Wg, const float* __restrict__ bias, const float* xs, float* ys,
                                               bool relu, int nt, int m_first) {
  constexpr int SXI = IN + 16, SYO = OUT + 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[MTW][4];
#pragma unroll
  for (int mi = 0; mi < MTW; ++mi)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[mi][j] = 0.f;
  const float4* wrow = reinterpret_cast<const float4*>(Wg + (size_t)(nt * 8 + g) * IN) + t;
#pragma unroll 8
  for (int kk = 0; kk < IN / 16; ++kk) {
    const float4 w = __ldg(wrow + kk * 4);
#pragma unroll
    for (int mi = 0; mi < MTW; ++mi) {
      const float* xr = xs + ((m_first + mi) * 16 + g) * SXI + kk * 16 + 4 * t;
      const float4 xa = *reinterpret_cast<const float4*>(xr);
      const float4 xb = *reinterpret_cast<const float4*>(xr + 8 * SXI);
      mma_tf32(acc[mi], xa.x, xb.x, xa.y, xb.y, w.x, w.y);
      mma_tf32(acc[mi], xa.z, xb.z, xa.w, xb.w, w.z, w.w);
    }
  }
  const int col = nt * 8 + 2 * t;
  const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
  for (int mi = 0; mi < MTW; ++mi) {
    const int r0 = (m_first + mi) * 16 + g;
    float2 v0 = make_float2(acc[mi][0] + b0, acc[mi][1] + b1), v1 = make_float2(acc[mi][2] + b0, acc[mi][3] + b1);
    if (relu) { v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f); }
    if (r0 < FR) *reinterpret_cast<float2*>(ys + r0 * SYO + col) = v0;
    if (r0 + 8 < FR) *reinterpret_cast<float2*>(ys + (r0 + 8) * SYO + col) = v1;
  }
}
template <int IN, int OUT>
__device__ __forceinline__ void tc_linear(const float* __restrict__ Wg, const float* __restrict__ bias, const float* xs, float* ys,
                                          bool relu) {
  constexpr int NT = OUT / 8;
  const int warp = threadIdx.x >> 5;
  if constexpr (NT >= FNW) {
    for (int nt = warp; nt < NT; nt += FNW) tc_linear_tile<IN, OUT, FMT>(Wg, bias, xs, ys, relu, nt, 0);
  } else {                                      // 8 column tiles: two warp groups split the row tiles {0, 1} | {2}
    static_assert(FNW == 2 * NT && FMT == 3, "warp split");
    if (warp < NT) tc_linear_tile<IN, OUT, 2>(Wg, bias, xs, ys, relu, warp, 0);
    else tc_linear_tile<IN, OUT, 1>(Wg, bias, xs, ys, relu, warp - NT, 2);
  }
}

// rows x columns [nt*8, +8) of dxs (+)= dys W        (K = OUT; B[k = o][n = i] = W[o][i])
template <int IN, int OUT, int MTW>
__device__ __forceinline__ void tc_linear_t_tile(const float* __restrict__ Wg, const float* dys, float* dxs, bool accumulate, int nt,
                                                 int m_first) {
  constexpr int SXI = IN + 16, SYO = OUT + 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[MTW][4];
#pragma unroll
  for (int mi = 0; mi < MTW; ++mi)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[mi][j] = 0.f;
  const float* wp = Wg + (size_t)(4 * t) * IN + nt * 8 + g;
#pragma unroll 4
  for (int kk = 0; kk < OUT / 16; ++kk) {
    const float* w = wp + (size_t)kk * 16 * IN;
    const float w0 = __ldg(w), w1 = __ldg(w + IN), w2 = __ldg(w + 2 * IN), w3 = __ldg(w + 3 * IN);
#pragma unroll
    for (int mi = 0; mi < MTW; ++mi) {
      const float* dr = dys + ((m_first + mi) * 16 + g) * SYO + kk * 16 + 4 * t;
      const float4 da = *reinterpret_cast<const float4*>(dr);
      const float4 db = *reinterpret_cast<const float4*>(dr + 8 * SYO);
      mma_tf32(acc[mi], da.x, db.x, da.y, db.y, w0, w1);
      mma_tf32(acc[mi], da.z, db.z, da.w, db.w, w2, w3);
    }
  }
  const int col = nt * 8 + 2 * t;
#pragma unroll
  for (int mi = 0; mi < MTW; ++mi) {
    const int r0 = (m_first + mi) * 16 + g;
    if (r0 < FR) {
      float2* p = reinterpret_cast<float2*>(dxs + r0 * SXI + col);
      float2 v = make_float2(acc[mi][0], acc[mi][1]);
      if (accumulate) { const float2 o = *p; v.x += o.x; v.y += o.y; }
      *p = v;
    }
    if (r0 + 8 < FR) {
      float2* p = reinterpret_cast<float2*>(dxs + (r0 + 8) * SXI + col);
      float2 v = make_float2(acc[mi][2], acc[mi][3]);
      if (accumulate) { const float2 o = *p; v.x += o.x; v.y += o.y; }
      *p = v;
    }
  }
}
template <int IN, int OUT>
__device__ __forceinline__ void tc_linear_t(const float* __restrict__ Wg, const float* dys, float* dxs, bool accumulate) {
  constexpr int NT = IN / 8;
  const int warp = threadIdx.x >> 5;
  if constexpr (NT >= FNW) {
    for (int nt = warp; nt < NT; nt += FNW) tc_linear_t_tile<IN, OUT, FMT>(Wg, dys, dxs, accumulate, nt, 0);
  } else {
    static_assert(FNW == 2 * NT && FMT == 3, "warp split");
    if (warp < NT) tc_linear_t_tile<IN, OUT, 2>(Wg, dys, dxs, accumulate, warp, 0);
    else tc_linear_t_tile<IN, OUT, 1>(Wg, dys, dxs, accumulate, warp - NT, 2);
  }
}

// dW[o][i] += sum_r dys[r][o] xs[r][i]: M = OUT, N = IN, K = the CTA's FR rows (5 k-steps); a warp keeps the A fragments of its
// 16 output features and walks column tiles; one vector red.add per pair of elements.
template <int IN, int OUT>
__device__ __forceinline__ void tc_wgrad(float* __restrict__ dWg, const float* dys, const float* xs) {
  constexpr int SXI = IN + 16, SYO = OUT + 16, MT = OUT / 16, NTT = IN / 8, KS = FR / 8;
  constexpr int WPM = MT >= FNW ? 1 : FNW / MT;          // warps sharing one row tile of dW
  static_assert(MT % (FNW / WPM) == 0 && NTT % WPM == 0, "tiling");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int mt = warp / WPM; mt < MT; mt += FNW / WPM) {
    float a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float* dp = dys + (8 * ks + t) * SYO + mt * 16 + g;
      a[ks][0] = dp[0]; a[ks][1] = dp[8]; a[ks][2] = dp[4 * SYO]; a[ks][3] = dp[4 * SYO + 8];
    }
    for (int nt = warp % WPM; nt < NTT; nt += WPM) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const float* xp = xs + (8 * ks + t) * SXI + nt * 8 + g;
        mma_tf32(acc, a[ks][0], a[ks][1], a[ks][2], a[ks][3], xp[0], xp[4 * SXI]);
      }
      float* o = dWg + (size_t)(mt * 16 + g) * IN + nt * 8 + 2 * t;
      atomicAdd(reinterpret_cast<float2*>(o), make_float2(acc[0], acc[1]));
      atomicAdd(reinterpret_cast<float2*>(o + (size_t)8 * IN), make_float2(acc[2], acc[3]));
    }
  }
}

// ---- mode dispatch: TC picks the tensor-core tiles and the padded shared-memory strides ----
template <bool TC, int IN, int OUT>
__device__ __forceinline__ void cta_linear(const float* __restrict__ Wg, const float* __restrict__ bias, const float* xs, float* ys,
                                           bool relu) {
  if constexpr (TC) tc_linear<IN, OUT>(Wg, bias, xs, ys, relu);
  else fma_linear<IN, OUT>(Wg, bias, xs, ys, relu);
}
template <bool TC, int IN, int OUT>
__device__ __forceinline__ void cta_linear_t(const float* __restrict__ Wg, const float* dys, float* dxs, bool accumulate) {
  if constexpr (TC) tc_linear_t<IN, OUT>(Wg, dys, dxs, accumulate);
  else fma_linear_t<IN, OUT>(Wg, dys, dxs, accumulate);
}
template <bool TC, int IN, int OUT>
__device__ __forceinline__ void cta_wgrad(float* __restrict__ dWg, float* __restrict__ dbg, const float* dys, const float* xs) {
  if constexpr (TC) tc_wgrad<IN, OUT>(dWg, dys, xs);
  else fma_wgrad<IN, OUT>(dWg, dys, xs);
  if (dbg) {
    constexpr int SYO = fu_stride<TC, OUT>();
    for (int o = threadIdx.x; o < OUT; o += FTHREADS) {
      float s = 0.f;
      for (int r = 0; r < FR; ++r) s += dys[r * SYO + o];
      atomicAdd(dbg + o, s);
    }
  }
}

// sum / max over the 16 lanes of a half-warp (xor shuffles stay inside the half)
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
static_assert(FE == 64 && FR % 8 == 0 && (FR - FTHREADS / 16) % 2 == 0, "half-warp row mapping: both halves of a warp are active together");

// LayerNorm over FE = 64 of the CTA's rows, half-warp per row (lane owns e = 4 l .. 4 l + 3: 128-bit accesses, one Philox
// block per lane and dropout site).  SX: row stride of the smem buffers.
// MODE 0: pre = a_row + b_row                      ; out = dropout_site(LN(pre))     (entry)
// MODE 1: pre = xs[r] + dropout_site(fs[r])        ; out = LN(pre)
template <int MODE, int SX>
__device__ __forceinline__ void cta_ln_fwd(const FuArgs& a, const DropRt& dr, int u0, const float* xs, const float* fs,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, float* pre_g, float* stat_g,
                                           float* outs, float* out_g, uint32_t site) {
  const int hw = threadIdx.x >> 4, l = threadIdx.x & 15;
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + l), bt = __ldg(reinterpret_cast<const float4*>(beta) + l);
  for (int r = hw; r < FR; r += FTHREADS / 16) {
    const int ul = r / FL, t = r % FL, u = u0 + ul;
    const bool ok = u < a.users;
    const int64_t gr = (int64_t)u * FL + t;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
      if (MODE == 0) {
        const float4 x = *reinterpret_cast<const float4*>(a.embs + (int64_t)u * a.ld_user + (int64_t)t * FE + 4 * l);
        const float4 p = __ldg(reinterpret_cast<const float4*>(a.P.pos_emb + t * FE) + l);
        v[0] = x.x + p.x; v[1] = x.y + p.y; v[2] = x.z + p.z; v[3] = x.w + p.w;
      } else {
        const float4 x = *reinterpret_cast<const float4*>(xs + r * SX + 4 * l);
        const float4 f = *reinterpret_cast<const float4*>(fs + r * SX + 4 * l);
        float fv[4] = {f.x, f.y, f.z, f.w};
        drop4(dr, site, (uint64_t)gr * (FE / 4) + l, fv);
        v[0] = x.x + fv[0]; v[1] = x.y + fv[1]; v[2] = x.z + fv[2]; v[3] = x.w + fv[3];
      }
    }
    const float mean = half_sum((v[0] + v[1]) + (v[2] + v[3])) / (float)FE;
    const float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean, d3 = v[3] - mean;
    const float var = half_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) / (float)FE;
    const float rstd = 1.0f / sqrtf(var + kLnEps);
    float y[4] = {d0 * rstd * gm.x + bt.x, d1 * rstd * gm.y + bt.y, d2 * rstd * gm.z + bt.z, d3 * rstd * gm.w + bt.w};
    if (MODE == 0 && ok) drop4(dr, site, (uint64_t)gr * (FE / 4) + l, y);
    if (!ok) { y[0] = 0.f; y[1] = 0.f; y[2] = 0.f; y[3] = 0.f; }
    *reinterpret_cast<float4*>(outs + r * SX + 4 * l) = make_float4(y[0], y[1], y[2], y[3]);
    if (ok) {
      reinterpret_cast<float4*>(pre_g + gr * FE)[l] = make_float4(v[0], v[1], v[2], v[3]);
      if (out_g) reinterpret_cast<float4*>(out_g + gr * FE)[l] = make_float4(y[0], y[1], y[2], y[3]);
      if (l == 0) *reinterpret_cast<float2*>(stat_g + 2 * gr) = make_float2(mean, rstd);
    }
  }
}

// copy the CTA's rows of a [FR][W] shared buffer (row stride SW) to the global stash (row-contiguous, coalesced)
template <int W, int SW>
__device__ __forceinline__ void cta_store_rows(const FuArgs& a, int u0, const float* s, float* g) {
  for (int idx = threadIdx.x; idx < FR * W / 4; idx += FTHREADS) {
    const int r = idx / (W / 4), c4 = idx % (W / 4);
    const int u = u0 + r / FL;
    if (u < a.users) reinterpret_cast<float4*>(g + ((int64_t)u * FL + r % FL) * W)[c4] = reinterpret_cast<const float4*>(s + r * SW)[c4];
  }
}
template <int W, int SW>
__device__ __forceinline__ void cta_load_rows(const FuArgs& a, int u0, float* s, const float* g) {
  for (int idx = threadIdx.x; idx < FR * W / 4; idx += FTHREADS) {
    const int r = idx / (W / 4), c4 = idx % (W / 4);
    const int u = u0 + r / FL;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < a.users) v = reinterpret_cast<const float4*>(g + ((int64_t)u * FL + r % FL) * W)[c4];
    reinterpret_cast<float4*>(s + r * SW)[c4] = v;
  }
}

// asynchronous variant (cp.async, 16 B per request): issued as soon as the destination buffer is free, completed by
// cp_async_wait_all() + __syncthreads() right before the first use, so that the stash reloads overlap the preceding phases
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int W, int SW>
__device__ __forceinline__ void cta_load_rows_async(const FuArgs& a, int u0, float* s, const float* g) {
  for (int idx = threadIdx.x; idx < FR * W / 4; idx += FTHREADS) {
    const int r = idx / (W / 4), c4 = idx % (W / 4);
    const int u = u0 + r / FL;
    if (u < a.users) cp_async16(s + r * SW + 4 * c4, g + ((int64_t)u * FL + r % FL) * W + 4 * c4);
    else reinterpret_cast<float4*>(s + r * SW)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool TC>
__global__ void __launch_bounds__(FTHREADS, 1) ue_fused_fwd_kernel(const __grid_constant__ FuArgs a) {
  using S = FuSmem<TC>;
  constexpr int SX = S::SX, SH = S::SH;
  extern __shared__ __align__(16) float fsm[];
  float* sX = fsm + S::kX; float* sQ = fsm + S::kQ; float* sK = fsm + S::kK; float* sV = fsm + S::kV;
  float* sC = fsm + S::kC; float* sM = fsm + S::kM; float* sP = fsm + S::kP; float* sH = fsm + S::kH;
  __shared__ float keyok[FR];
  const int u0 = blockIdx.x * FUPC;
  const int H = a.H, dk = FE / H;
  const float temp = sqrtf((float)dk);
  const DropRt drt = drop_rt(a.dc);
  if (threadIdx.x < FR) {
    const int u = u0 + threadIdx.x / FL;
    keyok[threadIdx.x] = (u < a.users && a.log_mask[(int64_t)u * FL + threadIdx.x % FL] != 0.f) ? 1.f : 0.f;
  }
  // ---- entry: x = dropout(LN(embs + pos)) ----
  cta_ln_fwd<0, SX>(a, drt, u0, nullptr, nullptr, a.P.ln_w, a.P.ln_b, a.W.pre0, a.W.stat0, sX, a.W.b[0].x_in, 0u);
  __syncthreads();
  for (int b = 0; b < a.n_blocks; ++b) {
    const iisan_ue_block_ptrs& bp = a.P.blocks[b];
    const UeBlockBufs& X = a.W.b[b];
    // ---- q, k, v ----
    cta_linear<TC, FE, FE>(bp.w_q, nullptr, sX, sQ, false);
    cta_linear<TC, FE, FE>(bp.w_k, nullptr, sX, sK, false);
    cta_linear<TC, FE, FE>(bp.w_v, nullptr, sX, sV, false);
    __syncthreads();
    cta_store_rows<FE, SX>(a, u0, sQ, X.q); cta_store_rows<FE, SX>(a, u0, sK, X.k); cta_store_rows<FE, SX>(a, u0, sV, X.v);
    // ---- scores + mask + softmax + attention dropout: a 16-lane group per (user, head, query) row, lane j = key ----
    {
      const int grp = threadIdx.x >> 4, j = threadIdx.x & 15;
      for (int row = grp; row < FUPC * H * FL; row += FTHREADS / 16) {
        const int ul = row / (H * FL), h = (row / FL) % H, i = row % FL, u = u0 + ul;
        float sc = -3.0e38f;
        if (j < FL) {
          const float4* qr = reinterpret_cast<const float4*>(sQ + (ul * FL + i) * SX + h * dk);
          const float4* kr = reinterpret_cast<const float4*>(sK + (ul * FL + j) * SX + h * dk);
          float d = 0.f;
          for (int c = 0; c < dk / 4; ++c) {
            const float4 q = qr[c], k = kr[c];
            d = fmaf(q.x, k.x, d); d = fmaf(q.y, k.y, d); d = fmaf(q.z, k.z, d); d = fmaf(q.w, k.w, d);
          }
          const float m = (j <= i && keyok[ul * FL + j] != 0.f) ? 0.f : kAttNeg;
          sc = __fadd_rn(__fdiv_rn(d, temp), m);
        }
        const float mx = half_max(sc);
        const float ex = j < FL ? expf(sc - mx) : 0.f;
        const float p = ex / half_sum(ex);
        if (j < FL) {
          const int64_t gi = (int64_t)u * H * FL * FL + (row % (H * FL)) * FL + j;
          float pd = 0.f;
          if (u < a.users) { X.p[gi] = p; pd = drop1(drt, 1u + 4u * b, (uint64_t)gi, p); }
          sP[row * FL + j] = pd;
        }
      }
    }
    __syncthreads();
    // ---- ctx = p v ----
    for (int idx = threadIdx.x; idx < FR * FE; idx += FTHREADS) {
      const int r = idx / FE, e = idx % FE, ul = r / FL, i = r % FL, h = e / dk;
      const float* pr = sP + ((ul * H + h) * FL + i) * FL;
      float acc = 0.f;
      for (int j = 0; j < FL; ++j) acc = fmaf(pr[j], sV[(ul * FL + j) * SX + e], acc);
      sC[r * SX + e] = acc;
    }
    __syncthreads();
    cta_store_rows<FE, SX>(a, u0, sC, X.ctx);
    // ---- fc + residual + LN1 ----
    cta_linear<TC, FE, FE>(bp.w_fc, nullptr, sC, sQ, false);          // sQ reused as the linear output
    __syncthreads();
    cta_ln_fwd<1, SX>(a, drt, u0, sX, sQ, bp.ln1_w, bp.ln1_b, X.pre1, X.stat1, sM, X.xmid, 2u + 4u * b);
    __syncthreads();
    // ---- FFN ----
    cta_linear<TC, FE, FF>(bp.w1, bp.b1, sM, sH, true);
    __syncthreads();
    cta_store_rows<FF, SH>(a, u0, sH, X.h1);
    cta_linear<TC, FF, FE>(bp.w2, bp.b2, sH, sQ, false);
    __syncthreads();
    float* dst_g = (b + 1 < a.n_blocks) ? a.W.b[b + 1].x_in : a.out;
    cta_ln_fwd<1, SX>(a, drt, u0, sM, sQ, bp.ln2_w, bp.ln2_b, X.pre2, X.stat2, sX, dst_g, 3u + 4u * b);
    __syncthreads();
  }
}

// LayerNorm backward over the CTA's rows (half-warp per row).  dys: [FR][FE] gradient of the LN output.
// MODE 0 (entry): dys <- dropout_bwd(dys) first; d pre -> d_embs (global, strided by user); d pos via dfs (dense copy).
// MODE 1: d pre -> dpre_s (residual branch) and dfs = dropout_bwd(d pre) (linear branch).
template <int MODE, int SX>
__device__ __forceinline__ void cta_ln_bwd(const FuArgs& a, const DropRt& dr, int u0, const float* dys, const float* pre_g,
                                           const float* stat_g, const float* __restrict__ gamma, float* dgamma_s, float* dbeta_s,
                                           float* dpre_s, float* dfs, uint32_t site) {
  const int hw = threadIdx.x >> 4, l = threadIdx.x & 15;
  const float4 gm4 = __ldg(reinterpret_cast<const float4*>(gamma) + l);
  const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
  float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = hw; r < FR; r += FTHREADS / 16) {
    const int ul = r / FL, t = r % FL, u = u0 + ul;
    const bool ok = u < a.users;
    const int64_t gr = (int64_t)u * FL + t;
    float mean = 0.f, rstd = 0.f;
    float d[4] = {0.f, 0.f, 0.f, 0.f}, xh[4] = {0.f, 0.f, 0.f, 0.f}, g[4];
    if (ok) {
      const float2 st = *reinterpret_cast<const float2*>(stat_g + 2 * gr);
      mean = st.x; rstd = st.y;
      const float4 dv = *reinterpret_cast<const float4*>(dys + r * SX + 4 * l);
      d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
      if (MODE == 0) drop4(dr, site, (uint64_t)gr * (FE / 4) + l, d);
      const float4 pv = reinterpret_cast<const float4*>(pre_g + gr * FE)[l];
      xh[0] = (pv.x - mean) * rstd; xh[1] = (pv.y - mean) * rstd; xh[2] = (pv.z - mean) * rstd; xh[3] = (pv.w - mean) * rstd;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ag[i] += d[i] * xh[i]; ab[i] += d[i];
      g[i] = d[i] * gm[i];
      s1 += g[i]; s2 += g[i] * xh[i];
    }
    s1 = half_sum(s1) / (float)FE; s2 = half_sum(s2) / (float)FE;
    float dp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dp[i] = ok ? rstd * (g[i] - s1 - xh[i] * s2) : 0.f;
    if (MODE == 0) {
      if (ok) *reinterpret_cast<float4*>(a.d_embs + (int64_t)u * a.ld_user + (int64_t)t * FE + 4 * l) = make_float4(dp[0], dp[1], dp[2], dp[3]);
      *reinterpret_cast<float4*>(dfs + r * SX + 4 * l) = make_float4(dp[0], dp[1], dp[2], dp[3]);
    } else {
      *reinterpret_cast<float4*>(dpre_s + r * SX + 4 * l) = make_float4(dp[0], dp[1], dp[2], dp[3]);
      if (ok) drop4(dr, site, (uint64_t)gr * (FE / 4) + l, dp);
      *reinterpret_cast<float4*>(dfs + r * SX + 4 * l) = make_float4(dp[0], dp[1], dp[2], dp[3]);
    }
  }
  // gamma / beta partials: fold the two half-warps, then one row per warp (summed by ln_partials_flush: no atomics)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ag[i] += __shfl_xor_sync(0xffffffffu, ag[i], 16);
    ab[i] += __shfl_xor_sync(0xffffffffu, ab[i], 16);
  }
  if ((threadIdx.x & 31) < 16) {
    const int w = threadIdx.x >> 5;
    *reinterpret_cast<float4*>(dgamma_s + w * FE + 4 * l) = make_float4(ag[0], ag[1], ag[2], ag[3]);
    *reinterpret_cast<float4*>(dbeta_s + w * FE + 4 * l) = make_float4(ab[0], ab[1], ab[2], ab[3]);
  }
}
// after a __syncthreads(): d gamma / d beta of the CTA's rows -> global (one red.add per feature)
__device__ __forceinline__ void ln_partials_flush(const float* dgamma_s, const float* dbeta_s, float* g_w, float* g_b) {
  if (threadIdx.x < 2 * FE) {
    const int e = threadIdx.x & (FE - 1);
    const float* src = (threadIdx.x < FE) ? dgamma_s : dbeta_s;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < FTHREADS / 32; ++w) sum += src[w * FE + e];
    atomicAdd(((threadIdx.x < FE) ? g_w : g_b) + e, sum);
  }
}

template <bool TC>
__global__ void __launch_bounds__(FTHREADS, 1) ue_fused_bwd_kernel(const __grid_constant__ FuArgs a) {
  using S = FuSmem<TC>;
  constexpr int SX = S::SX, SH = S::SH;
  extern __shared__ __align__(16) float fsm[];
  float* sX = fsm + S::kX;      // x_in of the block / d x_in accumulation target (see below)
  float* sQ = fsm + S::kQ; float* sK = fsm + S::kK; float* sV = fsm + S::kV;
  float* sC = fsm + S::kC; float* sM = fsm + S::kM; float* sP = fsm + S::kP;
  float* sH = fsm + S::kH;      // h1
  float* sD = fsm + S::kD;      // d h1
  __shared__ __align__(16) float sgam[(FTHREADS / 32) * FE], sbet[(FTHREADS / 32) * FE];   // per-warp LayerNorm gamma / beta partials
  const int u0 = blockIdx.x * FUPC;
  const int H = a.H, dk = FE / H;
  const float temp = sqrtf((float)dk);
  const DropRt drt = drop_rt(a.dc);
  // running gradient dY [FR][FE] lives in sC at block entry
  cta_load_rows<FE, SX>(a, u0, sC, a.d_out);
  __syncthreads();
  for (int b = a.n_blocks - 1; b >= 0; --b) {
    const iisan_ue_block_ptrs& bp = a.P.blocks[b];
    const iisan_ue_block_ptrs& bg = a.G.blocks[b];
    const UeBlockBufs& X = a.W.b[b];
    // ---- stash reloads whose buffers are free now: h1 -> sH, xmid -> sK, ctx -> sV, x_in -> sX (asynchronous) ----
    cta_load_rows_async<FF, SH>(a, u0, sH, X.h1);
    cta_load_rows_async<FE, SX>(a, u0, sK, X.xmid);
    cta_load_rows_async<FE, SX>(a, u0, sV, X.ctx);
    cta_load_rows_async<FE, SX>(a, u0, sX, X.x_in);
    // ---- LN2 backward: sC = dY -> sM = d pre2 (residual to xmid), sQ = df (w2 branch) ----
    cta_ln_bwd<1, SX>(a, drt, u0, sC, X.pre2, X.stat2, bp.ln2_w, sgam, sbet, sM, sQ, 3u + 4u * b);
    cp_async_wait_all();
    __syncthreads();
    ln_partials_flush(sgam, sbet, bg.ln2_w, bg.ln2_b);
    // ---- W2: dW2 += df^T h1 ; db2 ; d h1 = (df W2) * (h1 > 0) ----
    cta_wgrad<TC, FF, FE>(bg.w2, bg.b2, sQ, sH);
    cta_linear_t<TC, FF, FE>(bp.w2, sQ, sD, false);
    __syncthreads();
    for (int idx = threadIdx.x; idx < FR * FF; idx += FTHREADS) {
      const int o = (idx / FF) * SH + idx % FF;
      if (!(sH[o] > 0.f)) sD[o] = 0.f;
    }
    __syncthreads();
    // ---- W1: dW1 += dh1^T xmid ; db1 ; d xmid = d pre2 + dh1 W1 ----
    cta_wgrad<TC, FE, FF>(bg.w1, bg.b1, sD, sK);
    cta_linear_t<TC, FE, FF>(bp.w1, sD, sM, true);
    __syncthreads();
    cta_load_rows_async<FE, SX>(a, u0, sK, X.k);       // xmid is dead
    // ---- LN1 backward: sM = d xmid -> sC = d pre1 (residual to x_in), sQ = df (fc branch) ----
    cta_ln_bwd<1, SX>(a, drt, u0, sM, X.pre1, X.stat1, bp.ln1_w, sgam, sbet, sC, sQ, 2u + 4u * b);
    __syncthreads();
    ln_partials_flush(sgam, sbet, bg.ln1_w, bg.ln1_b);
    // ---- fc: dWfc += df^T ctx ; d ctx = df Wfc -> sM ----
    cta_wgrad<TC, FE, FE>(bg.w_fc, nullptr, sQ, sV);
    cta_linear_t<TC, FE, FE>(bp.w_fc, sQ, sM, false);
    __syncthreads();
    // ---- attention backward: q,k,v,p from the stash, d ctx in sM -> dq, dk, dv ----
    cta_load_rows_async<FE, SX>(a, u0, sQ, X.q); cta_load_rows_async<FE, SX>(a, u0, sV, X.v);      // k is already on its way
    float* sPd = sH;                               // dropout(p)      [FUPC][H][FL][FL]   (h1 is dead)
    float* sDs = sH + FUPC * FMAXH * FL * FL;      // d scores / temp
    for (int idx = threadIdx.x; idx < FUPC * H * FL * FL; idx += FTHREADS) {
      const int ul = idx / (H * FL * FL), u = u0 + ul;
      const int64_t gi = (int64_t)u * H * FL * FL + idx % (H * FL * FL);
      float p = 0.f, pd = 0.f;
      if (u < a.users) { p = X.p[gi]; pd = drop1(drt, 1u + 4u * b, (uint64_t)gi, p); }
      sP[idx] = p; sPd[idx] = pd;
    }
    cp_async_wait_all();
    __syncthreads();
    for (int idx = threadIdx.x; idx < FUPC * H * FL * FL; idx += FTHREADS) {       // d(dropout(p)) = dctx_i . v_j, then mask * scale
      const int ul = idx / (H * FL * FL), rem = idx % (H * FL * FL), u = u0 + ul;
      const int h = rem / (FL * FL), i = (rem / FL) % FL, j = rem % FL;
      const float4* dr = reinterpret_cast<const float4*>(sM + (ul * FL + i) * SX + h * dk);
      const float4* vr = reinterpret_cast<const float4*>(sV + (ul * FL + j) * SX + h * dk);
      float acc = 0.f;
      for (int c = 0; c < dk / 4; ++c) {
        const float4 d4 = dr[c], v4 = vr[c];
        acc = fmaf(d4.x, v4.x, acc); acc = fmaf(d4.y, v4.y, acc); acc = fmaf(d4.z, v4.z, acc); acc = fmaf(d4.w, v4.w, acc);
      }
      const int64_t gi = (int64_t)u * H * FL * FL + rem;
      sDs[idx] = (u < a.users) ? drop1(drt, 1u + 4u * b, (uint64_t)gi, acc) : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < FUPC * H * FL; r += FTHREADS) {                    // softmax backward, / temp
      float dot = 0.f;
      for (int j = 0; j < FL; ++j) dot = fmaf(sDs[r * FL + j], sP[r * FL + j], dot);
      for (int j = 0; j < FL; ++j) sDs[r * FL + j] = sP[r * FL + j] * (sDs[r * FL + j] - dot) / temp;
    }
    __syncthreads();
    float* sDq = sD; float* sDk = sD + S::RP * SX; float* sDv = sD + 2 * S::RP * SX;   // d h1 is dead
    static_assert(3 * S::RP * SX <= S::RP * SH, "dq | dk | dv alias the d h1 buffer");
    for (int idx = threadIdx.x; idx < FR * FE; idx += FTHREADS) {
      const int r = idx / FE, e = idx % FE, ul = r / FL, i = r % FL, h = e / dk;
      const float* ds = sDs + (ul * H + h) * FL * FL; const float* pd = sPd + (ul * H + h) * FL * FL;
      float aq = 0.f, ak = 0.f, av = 0.f;
      for (int j = 0; j < FL; ++j) {
        aq = fmaf(ds[i * FL + j], sK[(ul * FL + j) * SX + e], aq);       // dq[i] = sum_j ds[i,j] k[j]
        ak = fmaf(ds[j * FL + i], sQ[(ul * FL + j) * SX + e], ak);       // dk[i] = sum_j ds[j,i] q[j]
        av = fmaf(pd[j * FL + i], sM[(ul * FL + j) * SX + e], av);       // dv[i] = sum_j pd[j,i] dctx[j]
      }
      sDq[r * SX + e] = aq; sDk[r * SX + e] = ak; sDv[r * SX + e] = av;
    }
    __syncthreads();
    // ---- q/k/v projections: dW += d^T x_in ; d x_in = d pre1 + dq Wq + dk Wk + dv Wv (accumulated into sC) ----
    cta_wgrad<TC, FE, FE>(bg.w_q, nullptr, sDq, sX);
    cta_wgrad<TC, FE, FE>(bg.w_k, nullptr, sDk, sX);
    cta_wgrad<TC, FE, FE>(bg.w_v, nullptr, sDv, sX);
    cta_linear_t<TC, FE, FE>(bp.w_q, sDq, sC, true);
    __syncthreads();
    cta_linear_t<TC, FE, FE>(bp.w_k, sDk, sC, true);
    __syncthreads();
    cta_linear_t<TC, FE, FE>(bp.w_v, sDv, sC, true);
    __syncthreads();
  }
  // ---- entry LN + position embedding ----
  cta_ln_bwd<0, SX>(a, drt, u0, sC, a.W.pre0, a.W.stat0, a.P.ln_w, sgam, sbet, nullptr, sQ, 0u);
  __syncthreads();
  ln_partials_flush(sgam, sbet, a.G.ln_w, a.G.ln_b);
  for (int idx = threadIdx.x; idx < FL * FE; idx += FTHREADS) {       // d pos[t] = sum over the CTA's users
    const int t = idx / FE, e = idx % FE;
    float s = 0.f;
    for (int ul = 0; ul < FUPC; ++ul) s += sQ[(ul * FL + t) * SX + e];
    atomicAdd(a.G.pos_emb + idx, s);
  }
}

bool ue_fused_supported(const iisan_ue_desc& D) {
  return D.emb == FE && D.seq_len == FL && D.heads >= 1 && D.heads <= FMAXH && FE % D.heads == 0;
}

template <bool TC>
static int fused_forward_t(const FuArgs& a, int grid, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const size_t smem = FuSmem<TC>::kFwdFloats * sizeof(float);
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(ue_fused_fwd_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  { LaunchScope ls_(IISAN_K_USER, st); ue_fused_fwd_kernel<TC><<<grid, FTHREADS, smem, st>>>(a); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}
template <bool TC>
static int fused_backward_t(const FuArgs& a, int grid, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const size_t smem = FuSmem<TC>::kBwdFloats * sizeof(float);
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(ue_fused_bwd_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  { LaunchScope ls_(IISAN_K_USER, st); ue_fused_bwd_kernel<TC><<<grid, FTHREADS, smem, st>>>(a); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int ue_fused_forward(const iisan_ue_desc& D, const iisan_ue_params* P, const float* embs, int64_t ld_user, const float* log_mask,
                     void* workspace, float* out, cudaStream_t st) {
  FuArgs a{};
  a.users = D.users; a.H = D.heads; a.n_blocks = D.n_blocks; a.embs = embs; a.ld_user = ld_user; a.log_mask = log_mask;
  a.P = *P; a.W = UeLayout(D, workspace); a.out = out; a.dc = drop_cfg(D);
  const int grid = (D.users + FUPC - 1) / FUPC;
  return D.compute == IISAN_COMPUTE_BF16 ? fused_forward_t<true>(a, grid, st) : fused_forward_t<false>(a, grid, st);
}

int ue_fused_backward(const iisan_ue_desc& D, const iisan_ue_params* P, const iisan_ue_params* G, const float* embs, int64_t ld_user,
                      const float* log_mask, void* workspace, const float* d_out, float* d_embs, cudaStream_t st) {
  FuArgs a{};
  a.users = D.users; a.H = D.heads; a.n_blocks = D.n_blocks; a.embs = embs; a.ld_user = ld_user; a.log_mask = log_mask;
  a.P = *P; a.G = *G; a.W = UeLayout(D, workspace); a.d_out = d_out; a.d_embs = d_embs; a.dc = drop_cfg(D);
  const int grid = (D.users + FUPC - 1) / FUPC;
  return D.compute == IISAN_COMPUTE_BF16 ? fused_backward_t<true>(a, grid, st) : fused_backward_t<false>(a, grid, st);
}

}  // namespace iisan
