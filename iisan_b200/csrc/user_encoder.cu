// SASRec user encoder forward/backward (exact fp32 mode) + C-ABI entry points.
//
// Algorithm restated from the reference's PyTorch modules (nothing ported):
//   mask + entry              CC/model/encoders.py:53-58   key k visible to query q iff k<=q and log_mask[k]!=0,
//                                                          otherwise -1e9 is ADDED to the score
//   embedding + LN + dropout  CC/model/modules.py:89-96
//   attention block           CC/model/modules.py:21-32, 54-64  (bias-free Q/K/V/fc, softmax(QK^T/sqrt(dk)+mask),
//                                                          attn-dropout, LN(x + dropout(fc(ctx))), eps 1e-6)
//   feed-forward block        CC/model/modules.py:14-18    LN(y + dropout(W2 relu(W1 y)))
// Dropout uses a counter-based Philox stream (philox.cuh) instead of torch's generator; with
// drop_rate 0 / eval mode the result is the reference's.
#include <cstdlib>

#include "common.cuh"
#include "launch.cuh"
#include "gemm_simt.cuh"
#include "philox.cuh"
#include "user_encoder.cuh"

namespace iisan {

// ---- LayerNorm forward -----------------------------------------------------------------------------------
// MODE 0: pre = embs[u, t, :] + pos[t, :] ; out = dropout_site(LN(pre))
// MODE 1: pre = x[row, :] + dropout_site(f[row, :]) ; out = LN(pre)
template <int MODE>
__global__ void __launch_bounds__(256) ue_ln_fwd_kernel(int rows, int L, int E, const float* __restrict__ a, int64_t a_ld_user,
                                                        const float* __restrict__ b, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ pre,
                                                        float* __restrict__ stats, float* __restrict__ out, DropCfg dc, uint32_t site) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int per = E / 32;
  float v[kMaxEPerLane];
  float s = 0.f;
  const int u = warp / L, t = warp % L;
#pragma unroll
  for (int i = 0; i < kMaxEPerLane; ++i) {
    if (i < per) {
      const int e = lane + i * 32;
      float x;
      if (MODE == 0) x = a[(int64_t)u * a_ld_user + (int64_t)t * E + e] + b[t * E + e];
      else x = a[(int64_t)warp * E + e] + drop_apply(dc, site, (uint64_t)warp * E + e, b[(int64_t)warp * E + e]);
      v[i] = x; s += x;
    }
  }
  const float mean = warp_sum(s) / (float)E;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxEPerLane; ++i) if (i < per) { const float d = v[i] - mean; q += d * d; }
  const float var = warp_sum(q) / (float)E;
  const float rstd = 1.0f / sqrtf(var + kLnEps);
#pragma unroll
  for (int i = 0; i < kMaxEPerLane; ++i) {
    if (i < per) {
      const int e = lane + i * 32;
      pre[(int64_t)warp * E + e] = v[i];
      float y = (v[i] - mean) * rstd * gamma[e] + beta[e];
      if (MODE == 0) y = drop_apply(dc, site, (uint64_t)warp * E + e, y);
      out[(int64_t)warp * E + e] = y;
    }
  }
  if (lane == 0) { stats[2 * warp] = mean; stats[2 * warp + 1] = rstd; }
}

// ---- LayerNorm backward ------------------------------------------------------------------------------------
// MODE 0 (entry LN): dy <- dropout_bwd(dy) first; writes dpre to d_embs (strided by user) .
// MODE 1 (residual LN): writes dpre (residual branch gradient) and df = dropout_bwd(dpre) (linear branch).
template <int MODE>
__global__ void __launch_bounds__(256) ue_ln_bwd_kernel(int rows, int L, int E, const float* __restrict__ dy,
                                                        const float* __restrict__ pre, const float* __restrict__ stats,
                                                        const float* __restrict__ gamma, float* __restrict__ dgamma,
                                                        float* __restrict__ dbeta, float* __restrict__ dpre, int64_t dpre_ld_user,
                                                        float* __restrict__ df, DropCfg dc, uint32_t site) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int per = E / 32;
  float ag[kMaxEPerLane], ab[kMaxEPerLane];
#pragma unroll
  for (int i = 0; i < kMaxEPerLane; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  for (int row = blockIdx.x * 8 + wib; row < rows; row += gridDim.x * 8) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[kMaxEPerLane], g[kMaxEPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxEPerLane; ++i) {
      if (i < per) {
        const int e = lane + i * 32;
        float d = dy[(int64_t)row * E + e];
        if (MODE == 0) d = drop_apply(dc, site, (uint64_t)row * E + e, d);
        xh[i] = (pre[(int64_t)row * E + e] - mean) * rstd;
        ag[i] += d * xh[i]; ab[i] += d;
        g[i] = d * gamma[e];
        s1 += g[i]; s2 += g[i] * xh[i];
      }
    }
    s1 = warp_sum(s1) / (float)E; s2 = warp_sum(s2) / (float)E;
    const int u = row / L, t = row % L;
#pragma unroll
    for (int i = 0; i < kMaxEPerLane; ++i) {
      if (i < per) {
        const int e = lane + i * 32;
        const float dp = rstd * (g[i] - s1 - xh[i] * s2);
        if (MODE == 0) {
          dpre[(int64_t)u * dpre_ld_user + (int64_t)t * E + e] = dp;
          if (df) df[(int64_t)row * E + e] = dp;   // dense copy for the position-embedding column sums
        } else {
          dpre[(int64_t)row * E + e] = dp;
          df[(int64_t)row * E + e] = drop_apply(dc, site, (uint64_t)row * E + e, dp);
        }
      }
    }
  }
  __shared__ float rg[8][256], rb[8][256];
#pragma unroll
  for (int i = 0; i < kMaxEPerLane; ++i) if (i < per) { rg[wib][lane + i * 32] = ag[i]; rb[wib][lane + i * 32] = ab[i]; }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tg += rg[w][e]; tb += rb[w][e]; }
    atomicAdd(dgamma + e, tg); atomicAdd(dbeta + e, tb);
  }
}

// ---- attention ---------------------------------------------------------------------------------------------
constexpr int kMaxL = 16, kMaxE = 256, kMaxH = 8;

__global__ void __launch_bounds__(128) ue_attn_fwd_kernel(int L, int E, int H, const float* __restrict__ q, const float* __restrict__ k,
                                                          const float* __restrict__ v, const float* __restrict__ log_mask,
                                                          float* __restrict__ p_out, float* __restrict__ ctx, DropCfg dc, uint32_t site) {
  extern __shared__ float sm[];
  float* sq = sm; float* sk = sq + L * E; float* sv = sk + L * E; float* sp = sv + L * E;  // sp [H][L][L]
  __shared__ float keyok[kMaxL];
  const int u = blockIdx.x, dk = E / H;
  const float temp = sqrtf((float)dk);
  for (int i = threadIdx.x; i < L * E; i += blockDim.x) {
    sq[i] = q[(int64_t)u * L * E + i]; sk[i] = k[(int64_t)u * L * E + i]; sv[i] = v[(int64_t)u * L * E + i];
  }
  if (threadIdx.x < L) keyok[threadIdx.x] = (log_mask[u * L + threadIdx.x] != 0.f) ? 1.f : 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < H * L * L; idx += blockDim.x) {
    const int h = idx / (L * L), i = (idx / L) % L, j = idx % L;
    float d = 0.f;
    for (int c = 0; c < dk; ++c) d = fmaf(sq[i * E + h * dk + c], sk[j * E + h * dk + c], d);
    const float m = (j <= i && keyok[j] != 0.f) ? 0.f : kAttNeg;
    sp[idx] = __fadd_rn(__fdiv_rn(d, temp), m);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < H * L; r += blockDim.x) {
    float* row = sp + r * L;
    float mx = row[0];
    for (int j = 1; j < L; ++j) mx = fmaxf(mx, row[j]);
    float s = 0.f;
    for (int j = 0; j < L; ++j) { row[j] = expf(row[j] - mx); s += row[j]; }
    for (int j = 0; j < L; ++j) {
      const float p = row[j] / s;
      p_out[(int64_t)u * H * L * L + r * L + j] = p;
      row[j] = drop_apply(dc, site, (uint64_t)u * H * L * L + r * L + j, p);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < L * E; idx += blockDim.x) {
    const int i = idx / E, e = idx % E, h = e / dk;
    float a = 0.f;
    for (int j = 0; j < L; ++j) a = fmaf(sp[(h * L + i) * L + j], sv[j * E + e], a);
    ctx[(int64_t)u * L * E + idx] = a;
  }
}

__global__ void __launch_bounds__(128) ue_attn_bwd_kernel(int L, int E, int H, const float* __restrict__ q, const float* __restrict__ k,
                                                          const float* __restrict__ v, const float* __restrict__ p_in,
                                                          const float* __restrict__ dctx, float* __restrict__ dq, float* __restrict__ dk_,
                                                          float* __restrict__ dv, DropCfg dc, uint32_t site) {
  extern __shared__ float sm[];
  float* sq = sm; float* sk = sq + L * E; float* sv = sk + L * E; float* sd = sv + L * E;  // dctx
  float* sp = sd + L * E;            // p (softmax)          [H][L][L]
  float* spd = sp + H * L * L;       // dropout(p)           [H][L][L]
  float* sds = spd + H * L * L;      // d scores / temp      [H][L][L]
  const int u = blockIdx.x, dk = E / H;
  const float temp = sqrtf((float)dk);
  for (int i = threadIdx.x; i < L * E; i += blockDim.x) {
    const int64_t g = (int64_t)u * L * E + i;
    sq[i] = q[g]; sk[i] = k[g]; sv[i] = v[g]; sd[i] = dctx[g];
  }
  for (int i = threadIdx.x; i < H * L * L; i += blockDim.x) {
    const float p = p_in[(int64_t)u * H * L * L + i];
    sp[i] = p;
    spd[i] = drop_apply(dc, site, (uint64_t)u * H * L * L + i, p);
  }
  __syncthreads();
  // d(dropout(p)) -> dp
  for (int idx = threadIdx.x; idx < H * L * L; idx += blockDim.x) {
    const int h = idx / (L * L), i = (idx / L) % L, j = idx % L;
    float a = 0.f;
    for (int c = 0; c < dk; ++c) a = fmaf(sd[i * E + h * dk + c], sv[j * E + h * dk + c], a);
    sds[idx] = drop_apply(dc, site, (uint64_t)u * H * L * L + idx, a);   // mask * scale
  }
  __syncthreads();
  for (int r = threadIdx.x; r < H * L; r += blockDim.x) {
    float dot = 0.f;
    for (int j = 0; j < L; ++j) dot = fmaf(sds[r * L + j], sp[r * L + j], dot);
    for (int j = 0; j < L; ++j) sds[r * L + j] = sp[r * L + j] * (sds[r * L + j] - dot) / temp;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < L * E; idx += blockDim.x) {
    const int i = idx / E, e = idx % E, h = e / dk;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int j = 0; j < L; ++j) {
      aq = fmaf(sds[(h * L + i) * L + j], sk[j * E + e], aq);       // dq[i] = sum_j ds[i,j] k[j]
      ak = fmaf(sds[(h * L + j) * L + i], sq[j * E + e], ak);       // dk[i] = sum_j ds[j,i] q[j]
      av = fmaf(spd[(h * L + j) * L + i], sd[j * E + e], av);       // dv[i] = sum_j pd[j,i] dctx[j]
    }
    const int64_t g = (int64_t)u * L * E + idx;
    dq[g] = aq; dk_[g] = ak; dv[g] = av;
  }
}

static int ue_validate(const iisan_ue_desc* D) {
  if (!D) return IISAN_EINVAL;
  if (D->users <= 0 || D->seq_len <= 0 || D->seq_len > kMaxL) return IISAN_EINVAL;
  if (D->emb % 32 || D->emb > kMaxE || D->heads <= 0 || D->heads > kMaxH || D->emb % D->heads) return IISAN_EINVAL;
  if (D->n_blocks <= 0 || D->n_blocks > IISAN_MAX_BLOCKS) return IISAN_EINVAL;
  if (D->training && (D->dropout_p < 0.f || D->dropout_p >= 1.f)) return IISAN_EINVAL;
  return IISAN_OK;
}

static size_t attn_smem(const iisan_ue_desc& D, bool bwd) {
  const size_t LE = (size_t)D.seq_len * D.emb, HLL = (size_t)D.heads * D.seq_len * D.seq_len;
  return sizeof(float) * (bwd ? (4 * LE + 3 * HLL) : (3 * LE + HLL));
}

}  // namespace iisan

using namespace iisan;

// test / profiling switch: IISAN_B200_NO_FUSED_UE=1 forces the per-operator path
static const bool g_no_fused_ue = [] { const char* e = getenv("IISAN_B200_NO_FUSED_UE"); return e && e[0] == '1'; }();

extern "C" size_t iisan_user_encoder_workspace_bytes(const iisan_ue_desc* desc) {
  if (ue_validate(desc) != IISAN_OK) return 0;
  UeLayout L(*desc, nullptr);
  return L.bytes;
}

extern "C" int iisan_user_encoder_forward(const iisan_ue_desc* desc, const iisan_ue_params* P, const float* embs,
                                          int64_t ld_user, const float* log_mask, void* workspace, size_t workspace_bytes,
                                          float* out, iisan_stream_t stream) {
  IISAN_TRY(ue_validate(desc));
  if (!P || !embs || !log_mask || !workspace || !out) return IISAN_EINVAL;
  if (workspace_bytes < iisan_user_encoder_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  const iisan_ue_desc& D = *desc;
  cudaStream_t st = as_stream(stream);
  if (ue_fused_supported(D) && !g_no_fused_ue) return ue_fused_forward(D, P, embs, ld_user, log_mask, workspace, out, st);
  UeLayout W(D, workspace);
  const int R = D.users * D.seq_len, E = D.emb, L = D.seq_len, H = D.heads;
  const DropCfg dc = drop_cfg(D);
  const int ln_blocks = (R * 32 + 255) / 256;
  if (attn_smem(D, true) > 48 * 1024) return IISAN_EUNSUPPORTED;
  { LaunchScope ls_(IISAN_K_USER, st); ue_ln_fwd_kernel<0><<<ln_blocks, 256, 0, st>>>(R, L, E, embs, ld_user, P->pos_emb, P->ln_w, P->ln_b, W.pre0, W.stat0,
                                                W.b[0].x_in, dc, 0u); }
  IISAN_LAUNCH_OK();
  for (int b = 0; b < D.n_blocks; ++b) {
    const iisan_ue_block_ptrs& bp = P->blocks[b];
    UeBlockBufs& X = W.b[b];
    GemmBatch qkv{}; qkv.n = 3;
    qkv.p[0] = prob_linear(X.x_in, E, bp.w_q, nullptr, X.q, E, R, E, E);
    qkv.p[1] = prob_linear(X.x_in, E, bp.w_k, nullptr, X.k, E, R, E, E);
    qkv.p[2] = prob_linear(X.x_in, E, bp.w_v, nullptr, X.v, E, R, E, E);
    IISAN_TRY(launch_gemm(qkv, st));
    { LaunchScope ls_(IISAN_K_USER, st); ue_attn_fwd_kernel<<<D.users, 128, attn_smem(D, false), st>>>(L, E, H, X.q, X.k, X.v, log_mask, X.p, X.ctx, dc, 1u + 4u * b); }
    IISAN_LAUNCH_OK();
    GemmBatch fc{}; fc.n = 1;
    fc.p[0] = prob_linear(X.ctx, E, bp.w_fc, nullptr, W.lin, E, R, E, E);
    IISAN_TRY(launch_gemm(fc, st));
    { LaunchScope ls_(IISAN_K_USER, st); ue_ln_fwd_kernel<1><<<ln_blocks, 256, 0, st>>>(R, L, E, X.x_in, 0, W.lin, bp.ln1_w, bp.ln1_b, X.pre1, X.stat1, X.xmid, dc, 2u + 4u * b); }
    IISAN_LAUNCH_OK();
    GemmBatch f1{}; f1.n = 1;
    f1.p[0] = prob_linear(X.xmid, E, bp.w1, bp.b1, X.h1, 4 * E, R, 4 * E, E, 1);
    IISAN_TRY(launch_gemm(f1, st));
    GemmBatch f2{}; f2.n = 1;
    f2.p[0] = prob_linear(X.h1, 4 * E, bp.w2, bp.b2, W.lin, E, R, E, 4 * E);
    IISAN_TRY(launch_gemm(f2, st));
    float* dst = (b + 1 < D.n_blocks) ? W.b[b + 1].x_in : out;
    { LaunchScope ls_(IISAN_K_USER, st); ue_ln_fwd_kernel<1><<<ln_blocks, 256, 0, st>>>(R, L, E, X.xmid, 0, W.lin, bp.ln2_w, bp.ln2_b, X.pre2, X.stat2, dst, dc, 3u + 4u * b); }
    IISAN_LAUNCH_OK();
  }
  return IISAN_OK;
}

extern "C" int iisan_user_encoder_backward(const iisan_ue_desc* desc, const iisan_ue_params* P, const iisan_ue_params* G,
                                           const float* embs, int64_t ld_user, const float* log_mask, void* workspace,
                                           size_t workspace_bytes, const float* d_out, float* d_embs, iisan_stream_t stream) {
  IISAN_TRY(ue_validate(desc));
  if (!P || !G || !embs || !log_mask || !workspace || !d_out || !d_embs) return IISAN_EINVAL;
  if (workspace_bytes < iisan_user_encoder_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  const iisan_ue_desc& D = *desc;
  cudaStream_t st = as_stream(stream);
  if (ue_fused_supported(D) && !g_no_fused_ue) return ue_fused_backward(D, P, G, embs, ld_user, log_mask, workspace, d_out, d_embs, st);
  UeLayout W(D, workspace);
  const int R = D.users * D.seq_len, E = D.emb, L = D.seq_len, H = D.heads;
  const DropCfg dc = drop_cfg(D);
  const int lnb = min(148 * 4, (R + 7) / 8);
  const float* dy = d_out;
  for (int b = D.n_blocks - 1; b >= 0; --b) {
    const iisan_ue_block_ptrs& bp = P->blocks[b];
    const iisan_ue_block_ptrs& bg = G->blocks[b];
    UeBlockBufs& X = W.b[b];
    // LN2: dA = d pre2 (residual), df = dropout_bwd(dA) (w2 branch)
    { LaunchScope ls_(IISAN_K_USER, st); ue_ln_bwd_kernel<1><<<lnb, 256, 0, st>>>(R, L, E, dy, X.pre2, X.stat2, bp.ln2_w, bg.ln2_w, bg.ln2_b, W.dA, 0, W.df, dc, 3u + 4u * b); }
    IISAN_LAUNCH_OK();
    {
      GemmBatch g{}; g.n = 1; g.p[0] = prob_wgrad(W.df, E, X.h1, 4 * E, bg.w2, R, E, 4 * E); IISAN_TRY(launch_gemm(g, st));
      ColsumBatch c{}; c.n = 1; c.p[0] = {W.df, E, R, E, bg.b2}; IISAN_TRY(launch_colsum(c, st));
      GemmBatch d{}; d.n = 1; d.p[0] = prob_dgrad(W.df, E, bp.w2, W.dh1, 4 * E, R, E, 4 * E, X.h1, 4 * E); IISAN_TRY(launch_gemm(d, st));
      GemmBatch g1{}; g1.n = 1; g1.p[0] = prob_wgrad(W.dh1, 4 * E, X.xmid, E, bg.w1, R, 4 * E, E); IISAN_TRY(launch_gemm(g1, st));
      ColsumBatch c1{}; c1.n = 1; c1.p[0] = {W.dh1, 4 * E, R, 4 * E, bg.b1}; IISAN_TRY(launch_colsum(c1, st));
      // d xmid = dA + dh1 W1
      GemmBatch d1{}; d1.n = 1; d1.p[0] = prob_dgrad(W.dh1, 4 * E, bp.w1, W.dB, E, R, 4 * E, E, nullptr, 0, W.dA, E); IISAN_TRY(launch_gemm(d1, st));
    }
    // LN1: dA = d pre1 (residual to x_in), df = dropout_bwd (fc branch)
    { LaunchScope ls_(IISAN_K_USER, st); ue_ln_bwd_kernel<1><<<lnb, 256, 0, st>>>(R, L, E, W.dB, X.pre1, X.stat1, bp.ln1_w, bg.ln1_w, bg.ln1_b, W.dA, 0, W.df, dc, 2u + 4u * b); }
    IISAN_LAUNCH_OK();
    {
      GemmBatch g{}; g.n = 1; g.p[0] = prob_wgrad(W.df, E, X.ctx, E, bg.w_fc, R, E, E); IISAN_TRY(launch_gemm(g, st));
      GemmBatch d{}; d.n = 1; d.p[0] = prob_dgrad(W.df, E, bp.w_fc, W.dctx, E, R, E, E); IISAN_TRY(launch_gemm(d, st));
    }
    { LaunchScope ls_(IISAN_K_USER, st); ue_attn_bwd_kernel<<<D.users, 128, attn_smem(D, true), st>>>(L, E, H, X.q, X.k, X.v, X.p, W.dctx, W.dq, W.dk, W.dv, dc, 1u + 4u * b); }
    IISAN_LAUNCH_OK();
    {
      GemmBatch g{}; g.n = 3;
      g.p[0] = prob_wgrad(W.dq, E, X.x_in, E, bg.w_q, R, E, E);
      g.p[1] = prob_wgrad(W.dk, E, X.x_in, E, bg.w_k, R, E, E);
      g.p[2] = prob_wgrad(W.dv, E, X.x_in, E, bg.w_v, R, E, E);
      IISAN_TRY(launch_gemm(g, st));
      // d x_in = dA + dq Wq + dk Wk + dv Wv   -> dB
      GemmBatch d0{}; d0.n = 1; d0.p[0] = prob_dgrad(W.dq, E, bp.w_q, W.dB, E, R, E, E, nullptr, 0, W.dA, E); IISAN_TRY(launch_gemm(d0, st));
      GemmBatch d1{}; d1.n = 1; d1.p[0] = prob_dgrad(W.dk, E, bp.w_k, W.dB, E, R, E, E); d1.p[0].accumulate = 1; IISAN_TRY(launch_gemm(d1, st));
      GemmBatch d2{}; d2.n = 1; d2.p[0] = prob_dgrad(W.dv, E, bp.w_v, W.dB, E, R, E, E); d2.p[0].accumulate = 1; IISAN_TRY(launch_gemm(d2, st));
    }
    // dB is now d x_in of this block == d out of the previous block; keep it in dctx's sibling to free dB
    IISAN_CUDA_OK(cudaMemcpyAsync(W.dctx, W.dB, sizeof(float) * (size_t)R * E, cudaMemcpyDeviceToDevice, st));
    dy = W.dctx;
  }
  // entry LN + position embedding
  { LaunchScope ls_(IISAN_K_USER, st); ue_ln_bwd_kernel<0><<<lnb, 256, 0, st>>>(R, L, E, dy, W.pre0, W.stat0, P->ln_w, G->ln_w, G->ln_b, d_embs, ld_user, W.df, dc, 0u); }
  IISAN_LAUNCH_OK();
  ColsumBatch c{}; c.n = 1; c.p[0] = {W.df, (int64_t)L * E, D.users, L * E, G->pos_emb};
  IISAN_TRY(launch_colsum(c, st));
  return IISAN_OK;
}
