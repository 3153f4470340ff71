// In-batch softmax cross-entropy with debias, column-pad mask and reject mask (exact fp32 mode).
//
// Algorithm restated from CC/model/model.py:63-64 and :81-105 (nothing ported):
//   logit(r, c)  = <prec[r], score[c]> - log(pop_prob[ids_cols[c]])
//   column mask  : column (u, p) with p < L and log_mask_cols[u, p] == 0        -> -1e4   (:88-89)
//   reject mask  : ids_cols[c] occurs among the 11 ids of the row's user and
//                  c is not the row's label column                                -> -1e4   (:91-100)
//   label(i, j)  = (user_offset + i) * (L+1) + j + 1                                       (:82-85)
//   loss         = mean over rows with log_mask_rows != 0 of CE(row)                       (:102-104)
// The reference's O(B) python loop with B in-place index_puts becomes 11 integer compares per
// (row-user, column), evaluated in-tile; the [B*L, C] logits are never written to memory: the
// forward keeps one log-sum-exp per row, the backward recomputes the tile.
#include "common.cuh"
#include "launch.cuh"
#include "ce_umma.cuh"

namespace iisan {

constexpr float kNegMask = -1e4f;
constexpr int kMaxS = 17;  // slots per user (L + 1), L <= 16
constexpr int kCeMaxE = 256;

struct CeArgs {
  int B, Bc, L, S, E;
  int64_t C;            // Bc * S columns
  int64_t user_offset;
  const float* prec; const float* score;
  const int64_t* ids_rows; const int64_t* ids_cols;
  const float* lm_rows; const float* lm_cols;
  const float* pop;
};

__device__ __forceinline__ bool col_masked(const CeArgs& a, int64_t c) {
  const int p = (int)(c % a.S);
  if (p >= a.L) return false;                      // the appended all-ones column (:88)
  return a.lm_cols[(c / a.S) * a.L + p] == 0.f;
}

__device__ __forceinline__ bool in_reject(const int64_t* rid, int S, int64_t id) {
  bool hit = false;
#pragma unroll 1
  for (int k = 0; k < S; ++k) hit |= (rid[k] == id);
  return hit;
}

__device__ __forceinline__ float dot_row(const float* __restrict__ a_smem, const float* __restrict__ b, int E) {
  float d = 0.f;
  for (int e = 0; e < E; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(b + e);
    d = fmaf(a_smem[e], v.x, d); d = fmaf(a_smem[e + 1], v.y, d);
    d = fmaf(a_smem[e + 2], v.z, d); d = fmaf(a_smem[e + 3], v.w, d);
  }
  return d;
}

// ---- forward: one CTA per row --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_fwd_kernel(const CeArgs a, float* __restrict__ lse_out, double* __restrict__ acc_sum,
                                                     int* __restrict__ acc_cnt) {
  const int row = blockIdx.x;
  const int i = row / a.L, j = row % a.L;
  if (a.lm_rows[row] == 0.f) { if (threadIdx.x == 0) lse_out[row] = 0.f; return; }
  __shared__ float sp[kCeMaxE];
  __shared__ int64_t rid[kMaxS];
  __shared__ float red_m[8], red_s[8];
  __shared__ float lab_logit;
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) sp[e] = a.prec[(int64_t)row * a.E + e];
  if (threadIdx.x < a.S) rid[threadIdx.x] = a.ids_rows[(int64_t)i * a.S + threadIdx.x];
  __syncthreads();
  const int64_t label = (a.user_offset + i) * a.S + j + 1;
  float m = -INFINITY, s = 0.f;
  for (int64_t c = threadIdx.x; c < a.C; c += blockDim.x) {
    float lg;
    const int64_t id = a.ids_cols[c];
    if (col_masked(a, c) || (c != label && in_reject(rid, a.S, id))) lg = kNegMask;
    else lg = dot_row(sp, a.score + c * a.E, a.E) - logf(a.pop[id]);
    if (c == label) lab_logit = lg;
    if (lg > m) { s = s * expf(m - lg) + 1.f; m = lg; } else { s += expf(lg - m); }
  }
  // block reduce (max, sum)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float wm = warp_max(m);
  float ws = warp_sum(m == -INFINITY ? 0.f : s * expf(m - wm));
  if (lane == 0) { red_m[w] = wm; red_s[w] = ws; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float gm = red_m[0];
    for (int k = 1; k < 8; ++k) gm = fmaxf(gm, red_m[k]);
    float gs = 0.f;
    for (int k = 0; k < 8; ++k) gs += (red_m[k] == -INFINITY) ? 0.f : red_s[k] * expf(red_m[k] - gm);
    const float lse = gm + logf(gs);
    lse_out[row] = lse;
    atomicAdd(acc_sum, (double)(lse - lab_logit));
    atomicAdd(acc_cnt, 1);
  }
}

__global__ void ce_finalize_kernel(const double* acc_sum, const int* acc_cnt, float* loss_sum, int32_t* n_valid, float* loss) {
  const double s = *acc_sum; const int n = *acc_cnt;
  if (loss_sum) *loss_sum = (float)s;
  if (n_valid) *n_valid = n;
  if (loss) *loss = (float)(s / (double)n);
}

__device__ __forceinline__ float ce_scale(const float* g_sum, const float* g_mean, const int32_t* n_valid) {
  float s = 0.f;
  if (g_sum) s += __ldg(g_sum);
  if (g_mean) s += __ldg(g_mean) / (float)__ldg(n_valid);
  return s;
}

// ---- backward --------------------------------------------------------------------------------------------
// w(r, c) = scale * (exp(logit - lse_r) - [c == label_r]) for valid rows, 0 for masked entries.
// d_prec[r] = sum_c w(r,c) score[c] : one CTA per row, column chunks of 256 staged through smem.
__global__ void __launch_bounds__(256) ce_bwd_rows_kernel(const CeArgs a, const float* __restrict__ lse, const float* __restrict__ g_sum, const float* __restrict__ g_mean,
                                                          const int32_t* __restrict__ n_valid, float* __restrict__ d_prec) {
  const int row = blockIdx.x;
  const int i = row / a.L, j = row % a.L;
  if (a.lm_rows[row] == 0.f) {
    for (int e = threadIdx.x; e < a.E; e += blockDim.x) d_prec[(int64_t)row * a.E + e] = 0.f;
    return;
  }
  __shared__ float sp[kCeMaxE];
  __shared__ int64_t rid[kMaxS];
  __shared__ float wbuf[256];
  __shared__ float part[256];
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) sp[e] = a.prec[(int64_t)row * a.E + e];
  if (threadIdx.x < a.S) rid[threadIdx.x] = a.ids_rows[(int64_t)i * a.S + threadIdx.x];
  __syncthreads();
  const int64_t label = (a.user_offset + i) * a.S + j + 1;
  const float scale = ce_scale(g_sum, g_mean, n_valid);
  const float l = lse[row];
  const int groups = blockDim.x / a.E;           // E <= 256, E | 256 for E in {32,64,128,256}
  const int e = threadIdx.x % a.E, grp = threadIdx.x / a.E;
  float acc = 0.f;
  for (int64_t c0 = 0; c0 < a.C; c0 += blockDim.x) {
    const int64_t c = c0 + threadIdx.x;
    float w = 0.f;
    if (c < a.C) {
      const int64_t id = a.ids_cols[c];
      const bool masked = col_masked(a, c) || (c != label && in_reject(rid, a.S, id));
      const float lg = masked ? kNegMask : dot_row(sp, a.score + c * a.E, a.E) - logf(a.pop[id]);
      w = expf(lg - l) - (c == label ? 1.f : 0.f);
    }
    wbuf[threadIdx.x] = w * scale;
    __syncthreads();
    if (grp < groups) {
      const int nc = (int)imin64(blockDim.x, a.C - c0);
      for (int k = grp; k < nc; k += groups) {
        const float wk = wbuf[k];
        if (wk != 0.f) acc = fmaf(wk, a.score[(c0 + k) * a.E + e], acc);
      }
    }
    __syncthreads();
  }
  part[threadIdx.x] = (grp < groups) ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < a.E) {
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += part[g * a.E + threadIdx.x];
    d_prec[(int64_t)row * a.E + threadIdx.x] = t;
  }
}

// d_score[c] = sum_r w(r,c) prec[r] : one CTA per column, row chunks of 256 staged through smem.
__global__ void __launch_bounds__(256) ce_bwd_cols_kernel(const CeArgs a, const float* __restrict__ lse, const float* __restrict__ g_sum, const float* __restrict__ g_mean,
                                                          const int32_t* __restrict__ n_valid, float* __restrict__ d_score) {
  const int64_t c = blockIdx.x;
  __shared__ float sc[kCeMaxE];
  __shared__ float wbuf[256];
  __shared__ float part[256];
  const bool masked_c = col_masked(a, c);   // stays general: a masked column may still be some row's label
  for (int e = threadIdx.x; e < a.E; e += blockDim.x) sc[e] = a.score[c * a.E + e];
  __syncthreads();
  const int64_t id = a.ids_cols[c];
  const float debias = logf(a.pop[id]);
  const float scale = ce_scale(g_sum, g_mean, n_valid);
  const int R = a.B * a.L;
  const int groups = blockDim.x / a.E;
  const int e = threadIdx.x % a.E, grp = threadIdx.x / a.E;
  float acc = 0.f;
  for (int r0 = 0; r0 < R; r0 += blockDim.x) {
    const int r = r0 + threadIdx.x;
    float w = 0.f;
    if (r < R && a.lm_rows[r] != 0.f) {
      const int i = r / a.L, j = r % a.L;
      const int64_t label = (a.user_offset + i) * a.S + j + 1;
      bool rej = false;
      if (c != label) rej = in_reject(a.ids_rows + (int64_t)i * a.S, a.S, id);
      const float lg = (masked_c || rej) ? kNegMask : dot_row(sc, a.prec + (int64_t)r * a.E, a.E) - debias;
      w = expf(lg - lse[r]) - (c == label ? 1.f : 0.f);
    }
    wbuf[threadIdx.x] = w * scale;
    __syncthreads();
    if (grp < groups) {
      const int nr = min((int)blockDim.x, R - r0);
      for (int k = grp; k < nr; k += groups) {
        const float wk = wbuf[k];
        if (wk != 0.f) acc = fmaf(wk, a.prec[(int64_t)(r0 + k) * a.E + e], acc);
      }
    }
    __syncthreads();
  }
  part[threadIdx.x] = (grp < groups) ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < a.E) {
    float t = 0.f;
    for (int g = 0; g < groups; ++g) t += part[g * a.E + threadIdx.x];
    d_score[c * a.E + threadIdx.x] = t;
  }
}

// ---- mask probe (parity tests) ------------------------------------------------------------------------------
__global__ void ce_masks_kernel(const CeArgs a, uint8_t* __restrict__ out) {
  const int row = blockIdx.x;
  const int i = row / a.L, j = row % a.L;
  const int64_t label = (a.user_offset + i) * a.S + j + 1;
  const bool valid = a.lm_rows[row] != 0.f;
  for (int64_t c = threadIdx.x; c < a.C; c += blockDim.x) {
    uint8_t b = 0;
    if (col_masked(a, c)) b |= 1;
    if (c != label && in_reject(a.ids_rows + (int64_t)i * a.S, a.S, a.ids_cols[c])) b |= 2;
    if (c == label) b |= 4;
    if (valid) b |= 8;
    out[(int64_t)row * a.C + c] = b;
  }
}

static int ce_validate(const iisan_ce_desc* d) {
  if (!d) return IISAN_EINVAL;
  if (d->row_users <= 0 || d->col_users <= 0 || d->seq_len <= 0 || d->seq_len + 1 > kMaxS) return IISAN_EINVAL;
  if (d->emb <= 0 || d->emb > kCeMaxE || (256 % d->emb) != 0 || d->emb % 4) return IISAN_EINVAL;
  if (d->user_offset < 0 || d->user_offset + d->row_users > d->col_users) return IISAN_EINVAL;
  return IISAN_OK;
}

static CeArgs make_args(const iisan_ce_desc& d, const float* prec, const float* score, const int64_t* ids_rows,
                        const int64_t* ids_cols, const float* lm_rows, const float* lm_cols, const float* pop) {
  CeArgs a;
  a.B = d.row_users; a.Bc = d.col_users; a.L = d.seq_len; a.S = d.seq_len + 1; a.E = d.emb;
  a.C = (int64_t)d.col_users * a.S; a.user_offset = d.user_offset;
  a.prec = prec; a.score = score; a.ids_rows = ids_rows; a.ids_cols = ids_cols;
  a.lm_rows = lm_rows; a.lm_cols = lm_cols; a.pop = pop;
  return a;
}

struct CeLayout {
  float* lse; double* acc_sum; int* acc_cnt; size_t bytes;
  CeLayout(const iisan_ce_desc& d, void* ws) {
    Arena a(ws);
    lse = a.take<float>((size_t)d.row_users * d.seq_len);
    acc_sum = a.take<double>(1); acc_cnt = a.take<int>(1);
    bytes = a.off;
  }
};

}  // namespace iisan

using namespace iisan;

extern "C" size_t iisan_inbatch_ce_workspace_bytes(const iisan_ce_desc* desc) {
  if (ce_validate(desc) != IISAN_OK) return 0;
  if (desc->compute == IISAN_COMPUTE_BF16 && ce_fast_supported(*desc)) return ce_fast_workspace_bytes(*desc);
  CeLayout L(*desc, nullptr);
  return L.bytes;
}

extern "C" int iisan_inbatch_ce_forward(const iisan_ce_desc* desc, const float* prec, const float* score, const int64_t* ids_rows,
                                        const int64_t* ids_cols, const float* log_mask_rows, const float* log_mask_cols,
                                        const float* pop_prob, void* workspace, size_t workspace_bytes, float* loss_sum,
                                        int32_t* n_valid, float* loss, iisan_stream_t stream) {
  IISAN_TRY(ce_validate(desc));
  if (!prec || !score || !ids_rows || !ids_cols || !log_mask_rows || !log_mask_cols || !pop_prob || !workspace) return IISAN_EINVAL;
  if (workspace_bytes < iisan_inbatch_ce_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (desc->compute == IISAN_COMPUTE_BF16 && ce_fast_supported(*desc))
    return ce_fast_forward(*desc, prec, score, ids_rows, ids_cols, log_mask_rows, log_mask_cols, pop_prob, workspace, loss_sum, n_valid, loss, st);
  CeLayout W(*desc, workspace);
  const CeArgs a = make_args(*desc, prec, score, ids_rows, ids_cols, log_mask_rows, log_mask_cols, pop_prob);
  IISAN_CUDA_OK(cudaMemsetAsync(W.acc_sum, 0, 256 + sizeof(int), st));   // acc_sum and acc_cnt are adjacent 256 B slots
  { LaunchScope ls_(IISAN_K_CE, st); ce_fwd_kernel<<<a.B * a.L, 256, 0, st>>>(a, W.lse, W.acc_sum, W.acc_cnt); }
  IISAN_LAUNCH_OK();
  { LaunchScope ls_(IISAN_K_CE, st); ce_finalize_kernel<<<1, 1, 0, st>>>(W.acc_sum, W.acc_cnt, loss_sum, n_valid, loss); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

extern "C" int iisan_inbatch_ce_backward(const iisan_ce_desc* desc, const float* prec, const float* score, const int64_t* ids_rows,
                                         const int64_t* ids_cols, const float* log_mask_rows, const float* log_mask_cols,
                                         const float* pop_prob, void* workspace, size_t workspace_bytes, const float* grad_loss_sum,
                                         const float* grad_loss_mean, const int32_t* n_valid, float* d_prec, float* d_score,
                                         iisan_stream_t stream) {
  IISAN_TRY(ce_validate(desc));
  if (!prec || !score || !ids_rows || !ids_cols || !log_mask_rows || !log_mask_cols || !pop_prob || !workspace ||
      (!grad_loss_sum && !grad_loss_mean) || !n_valid || !d_prec || !d_score)
    return IISAN_EINVAL;
  if (workspace_bytes < iisan_inbatch_ce_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (desc->compute == IISAN_COMPUTE_BF16 && ce_fast_supported(*desc))
    return ce_fast_backward(*desc, log_mask_rows, log_mask_cols, workspace, grad_loss_sum, grad_loss_mean, n_valid, d_prec, d_score, st);
  CeLayout W(*desc, workspace);
  const CeArgs a = make_args(*desc, prec, score, ids_rows, ids_cols, log_mask_rows, log_mask_cols, pop_prob);
  { LaunchScope ls_(IISAN_K_CE, st); ce_bwd_rows_kernel<<<a.B * a.L, 256, 0, st>>>(a, W.lse, grad_loss_sum, grad_loss_mean, n_valid, d_prec); }
  IISAN_LAUNCH_OK();
  { LaunchScope ls_(IISAN_K_CE, st); ce_bwd_cols_kernel<<<(unsigned)a.C, 256, 0, st>>>(a, W.lse, grad_loss_sum, grad_loss_mean, n_valid, d_score); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

extern "C" int iisan_inbatch_ce_masks(const iisan_ce_desc* desc, const int64_t* ids_rows, const int64_t* ids_cols,
                                      const float* log_mask_rows, const float* log_mask_cols, uint8_t* out, iisan_stream_t stream) {
  IISAN_TRY(ce_validate(desc));
  if (!ids_rows || !ids_cols || !log_mask_rows || !log_mask_cols || !out) return IISAN_EINVAL;
  const CeArgs a = make_args(*desc, nullptr, nullptr, ids_rows, ids_cols, log_mask_rows, log_mask_cols, nullptr);
  { LaunchScope ls_(IISAN_K_CE, as_stream(stream)); ce_masks_kernel<<<a.B * a.L, 256, 0, as_stream(stream)>>>(a, out); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

extern "C" int iisan_inbatch_ce_masks_fast(const iisan_ce_desc* desc, const int64_t* ids_rows, const int64_t* ids_cols,
                                           const float* log_mask_rows, const float* log_mask_cols, void* workspace,
                                           size_t workspace_bytes, uint8_t* out, iisan_stream_t stream) {
  IISAN_TRY(ce_validate(desc));
  if (!ids_rows || !ids_cols || !log_mask_rows || !log_mask_cols || !workspace || !out) return IISAN_EINVAL;
  if (desc->compute != IISAN_COMPUTE_BF16 || !ce_fast_supported(*desc)) return IISAN_EUNSUPPORTED;
  if (workspace_bytes < ce_fast_workspace_bytes(*desc)) return IISAN_EWORKSPACE;
  return ce_fast_masks(*desc, ids_rows, ids_cols, log_mask_rows, log_mask_cols, workspace, out, as_stream(stream));
}
