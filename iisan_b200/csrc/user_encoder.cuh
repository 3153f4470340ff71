// Shared pieces of the SASRec user-encoder kernels (user_encoder.cu: one kernel per operator; user_encoder_fused.cu: whole
// encoder per CTA): dropout configuration, forward->backward stash layout.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace iisan {

constexpr float kLnEps = 1e-6f;
constexpr float kAttNeg = -1e9f;
constexpr int kMaxEPerLane = 8;  // E <= 256

struct DropCfg { int on; float p; float scale; uint64_t seed, offset; const uint64_t* offset_dev; uint32_t thr; };

__device__ __forceinline__ float drop_apply(const DropCfg& c, uint32_t site, uint64_t idx, float v) {
  if (!c.on) return v;
  const uint64_t off = c.offset_dev ? __ldg(reinterpret_cast<const unsigned long long*>(c.offset_dev)) : c.offset;
  return dropout_keep_thr(c.seed, off, site, idx, c.thr) ? v * c.scale : 0.f;
}

// Per-thread dropout state of the fused kernels: the step offset is read once, and one Philox block serves the four
// consecutive elements 4*idx4 .. 4*idx4+3 (same mask function as drop_apply).
struct DropRt { bool on; uint32_t thr; float scale; uint64_t seed, off; };
__device__ __forceinline__ DropRt drop_rt(const DropCfg& c) {
  DropRt d;
  d.on = c.on != 0; d.thr = c.thr; d.scale = c.scale; d.seed = c.seed;
  d.off = (c.on && c.offset_dev) ? __ldg(reinterpret_cast<const unsigned long long*>(c.offset_dev)) : c.offset;
  return d;
}
__device__ __forceinline__ void drop4(const DropRt& d, uint32_t site, uint64_t idx4, float (&v)[4]) {
  if (!d.on) return;
  uint32_t r[4];
  philox4x32_10(d.seed, d.off, site, idx4, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = r[i] >= d.thr ? v[i] * d.scale : 0.f;
}
__device__ __forceinline__ float drop1(const DropRt& d, uint32_t site, uint64_t idx, float v) {
  if (!d.on) return v;
  return dropout_keep_thr(d.seed, d.off, site, idx, d.thr) ? v * d.scale : 0.f;
}


// ---- workspace layout --------------------------------------------------------------------------------------
struct UeBlockBufs {
  float *x_in, *q, *k, *v, *p, *ctx, *pre1, *stat1, *xmid, *h1, *pre2, *stat2;
};
struct UeLayout {
  float *pre0, *stat0;
  UeBlockBufs b[IISAN_MAX_BLOCKS];
  float *lin;                       // [R, E] scratch: fc / w2 outputs
  float *dA, *dB, *dq, *dk, *dv, *dctx, *dh1, *df;
  size_t bytes;
  UeLayout() = default;
  UeLayout(const iisan_ue_desc& D, void* ws) {
    Arena a(ws);
    const size_t R = (size_t)D.users * D.seq_len, E = D.emb;
    pre0 = a.take<float>(R * E); stat0 = a.take<float>(R * 2);
    for (int i = 0; i < D.n_blocks; ++i) {
      UeBlockBufs& x = b[i];
      x.x_in = a.take<float>(R * E); x.q = a.take<float>(R * E); x.k = a.take<float>(R * E); x.v = a.take<float>(R * E);
      x.p = a.take<float>((size_t)D.users * D.heads * D.seq_len * D.seq_len);
      x.ctx = a.take<float>(R * E); x.pre1 = a.take<float>(R * E); x.stat1 = a.take<float>(R * 2);
      x.xmid = a.take<float>(R * E); x.h1 = a.take<float>(R * 4 * E); x.pre2 = a.take<float>(R * E); x.stat2 = a.take<float>(R * 2);
    }
    lin = a.take<float>(R * E);
    dA = a.take<float>(R * E); dB = a.take<float>(R * E); dq = a.take<float>(R * E); dk = a.take<float>(R * E);
    dv = a.take<float>(R * E); dctx = a.take<float>(R * E); dh1 = a.take<float>(R * 4 * E); df = a.take<float>(R * E);
    bytes = a.off;
  }
};


inline DropCfg drop_cfg(const iisan_ue_desc& D) {
  DropCfg c;
  c.on = (D.training && D.dropout_p > 0.f) ? 1 : 0;
  c.p = D.dropout_p; c.scale = c.on ? 1.0f / (1.0f - D.dropout_p) : 1.0f;
  c.seed = D.seed; c.offset = D.offset; c.offset_dev = D.offset_dev; c.thr = dropout_threshold(D.dropout_p);
  return c;
}

// fused whole-encoder kernels (user_encoder_fused.cu)
bool ue_fused_supported(const iisan_ue_desc& D);
int ue_fused_forward(const iisan_ue_desc& D, const iisan_ue_params* P, const float* embs, int64_t ld_user, const float* log_mask,
                     void* workspace, float* out, cudaStream_t st);
int ue_fused_backward(const iisan_ue_desc& D, const iisan_ue_params* P, const iisan_ue_params* G, const float* embs, int64_t ld_user,
                      const float* log_mask, void* workspace, const float* d_out, float* d_embs, cudaStream_t st);

}  // namespace iisan
