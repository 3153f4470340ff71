// Evaluation scoring: rank of the held-out item among the whole catalogue, per user.
//
// Algorithm restated from the reference's Python loop (nothing ported): CC/data_utils/metrics.py:212-222 computes, per user,
//   scores = prec_emb . item_embeddings^T  over ids 0..item_num ; scores[history] = -inf ; drop id 0 ;
// and metrics_topK (:59-67) sorts the scores, finds the 1-based position `rank` of the target and reports Hit@K = [rank <= K],
// nDCG@K = 1/log2(rank + 1).  The position is rank = 1 + #{ i in 1..item_num, i not in history : score_i > score_target }
// (strict: exact ties have no defined order in the reference's argsort either).  No sort and no [users, items] score matrix:
// one CTA per 8 users streams the item table (L2-resident: 19 k x 64 fp32 = 4.9 MB) and counts.
#include "common.cuh"
#include "launch.cuh"

namespace iisan {

constexpr int ER_THREADS = 256;
constexpr int ER_MAX_E = 256;
constexpr int ER_UPC = 8;          // users per CTA: an item row is loaded once and scored against 8 user vectors

// score of item row `row` for the CTA's users (fixed summation order: every thread that scores the same pair gets the same bits)
__device__ __forceinline__ void er_dots(const float* __restrict__ row, const float (*su)[ER_MAX_E], int E, float (&acc)[ER_UPC]) {
#pragma unroll
  for (int q = 0; q < ER_UPC; ++q) acc[q] = 0.f;
  for (int k = 0; k < E; k += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
#pragma unroll
    for (int q = 0; q < ER_UPC; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(&su[q][k]);
      acc[q] = fmaf(v.x, w.x, acc[q]); acc[q] = fmaf(v.y, w.y, acc[q]); acc[q] = fmaf(v.z, w.z, acc[q]); acc[q] = fmaf(v.w, w.w, acc[q]);
    }
  }
}
__device__ __forceinline__ float er_dot1(const float* __restrict__ row, const float* u, int E) {      // same order as er_dots
  float acc = 0.f;
  for (int k = 0; k < E; k += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row + k));
    acc = fmaf(v.x, u[k], acc); acc = fmaf(v.y, u[k + 1], acc); acc = fmaf(v.z, u[k + 2], acc); acc = fmaf(v.w, u[k + 3], acc);
  }
  return acc;
}

__global__ void __launch_bounds__(ER_THREADS) eval_rank_kernel(const float* __restrict__ prec, const float* __restrict__ items,
                                                               const int64_t* __restrict__ targets, const int64_t* __restrict__ history,
                                                               int users, int n_items1, int E, int H, int32_t* __restrict__ ranks) {
  __shared__ __align__(16) float su[ER_UPC][ER_MAX_E];
  __shared__ float s_t[ER_UPC];
  __shared__ int s_cnt[ER_UPC];
  const int u0 = blockIdx.x * ER_UPC;
  for (int idx = threadIdx.x; idx < ER_UPC * E; idx += ER_THREADS) {
    const int q = idx / E, k = idx % E;
    su[q][k] = (u0 + q < users) ? prec[(int64_t)(u0 + q) * E + k] : 0.f;
  }
  if (threadIdx.x < ER_UPC) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < ER_UPC) {
    const int q = threadIdx.x;
    // a target outside [0, n_items1) cannot be scored: +inf makes no item beat it -- the host side rejects such inputs
    // (iisan_b200/eval.py), this only keeps the kernel inside the table
    const int64_t tg = (u0 + q < users) ? targets[u0 + q] : -1;
    s_t[q] = (tg >= 0 && tg < n_items1) ? er_dot1(items + tg * E, su[q], E) : INFINITY;
  }
  __syncthreads();
  float st[ER_UPC];
  int cnt[ER_UPC];
#pragma unroll
  for (int q = 0; q < ER_UPC; ++q) { st[q] = s_t[q]; cnt[q] = 0; }
  for (int i = 1 + threadIdx.x; i < n_items1; i += ER_THREADS) {
    float acc[ER_UPC];
    er_dots(items + (int64_t)i * E, su, E, acc);
#pragma unroll
    for (int q = 0; q < ER_UPC; ++q) cnt[q] += acc[q] > st[q] ? 1 : 0;
  }
  // history items do not compete (their score is -inf): take back those that were counted, each distinct id once; id 0 (padding
  // of the history list, and the padding item, which is dropped anyway) is skipped
  for (int idx = threadIdx.x; idx < ER_UPC * H; idx += ER_THREADS) {
    const int q = idx / H, j = idx % H;
    if (u0 + q >= users) continue;
    const int64_t* hu = history + (int64_t)(u0 + q) * H;
    const int64_t h = hu[j];
    if (h <= 0 || h >= n_items1) continue;
    bool dup = false;
    for (int p = 0; p < j; ++p) dup |= (hu[p] == h);
    if (!dup && er_dot1(items + h * E, su[q], E) > s_t[q]) atomicSub(&s_cnt[q], 1);
  }
#pragma unroll
  for (int q = 0; q < ER_UPC; ++q) {
    int c = cnt[q];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt[q], c);
  }
  __syncthreads();
  if (threadIdx.x < ER_UPC && u0 + threadIdx.x < users) ranks[u0 + threadIdx.x] = 1 + s_cnt[threadIdx.x];
}

}  // namespace iisan

using namespace iisan;

extern "C" int iisan_eval_ranks(const float* prec, const float* item_embs, const int64_t* targets, const int64_t* history, int32_t users,
                                int32_t n_items1, int32_t emb, int32_t hist_len, int32_t* ranks, iisan_stream_t stream) {
  if (!prec || !item_embs || !targets || !ranks || users <= 0 || n_items1 <= 1 || emb <= 0 || emb > ER_MAX_E || (emb & 3) || hist_len < 0 ||
      (hist_len > 0 && !history))
    return IISAN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(item_embs) & 15) != 0) return IISAN_EINVAL;
  cudaStream_t st = as_stream(stream);
  { LaunchScope ls_(IISAN_K_MISC, st); eval_rank_kernel<<<(users + ER_UPC - 1) / ER_UPC, ER_THREADS, 0, st>>>(prec, item_embs, targets, history, users, n_items1, emb, hist_len, ranks); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}
