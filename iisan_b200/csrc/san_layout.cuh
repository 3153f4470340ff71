// Workspace layout of the side-adapter network: forward->backward stash + backward scratch.
// Built identically by iisan_san_workspace_bytes(), the forward and the backward.
#pragma once
#include "common.cuh"

namespace iisan {

int san_validate(const iisan_san_desc* d);
// bf16 tensor-core mode (san_bf16.cu)
size_t san_bf16_workspace_bytes(const iisan_san_desc& D);
int san_chain_eligible_if_bf16(const iisan_san_desc& D);
int san_bf16_supported(const iisan_san_desc& D);
int san_forward_bf16(const iisan_san_desc* D, const iisan_san_params* P, const void* image, const void* text, void* ws,
                     float* out, cudaStream_t st);
int san_backward_bf16(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G, const void* image,
                      const void* text, void* ws, const float* d_out, cudaStream_t st);

struct SanLayout {
  // per stage stash (null where the tower is idle in that stage)
  float* x_t[IISAN_MAX_STAGES]; float* z_t[IISAN_MAX_STAGES]; float* last_t[IISAN_MAX_STAGES];
  float* x_i[IISAN_MAX_STAGES]; float* z_i[IISAN_MAX_STAGES]; float* last_i[IISAN_MAX_STAGES];
  float* x_m[IISAN_MAX_STAGES]; float* z_m[IISAN_MAX_STAGES]; float* last_m[IISAN_MAX_STAGES];
  float* dp[IISAN_MAX_STAGES];      // down_project output of the wide modality (Versa)
  float* a_t[IISAN_MAX_STAGES]; float* a_i[IISAN_MAX_STAGES]; float* a_m[IISAN_MAX_STAGES];   // GELU: pre-activations [N, r] (null for ReLU)
  float *head_t, *head_i, *head_m;  // fc outputs
  // scratch
  float* wide_dense;                // [N, max(d)] fp32 copy of one wide hidden-state layer (Versa)
  float *dhead_t, *dhead_i, *dhead_m;
  float *dy_t, *dx_t, *dz_t, *dy_i, *dx_i, *dz_i, *dy_m, *dx_m, *dz_m, *ddp;
  size_t bytes;

  SanLayout(const iisan_san_desc& D, void* ws) {
    Arena a(ws);
    const size_t N = (size_t)D.n_items;
    const int E = D.emb;
    const bool dimdiff = D.d_text != D.d_img;
    for (int s = 0; s < IISAN_MAX_STAGES; ++s) {
      x_t[s] = z_t[s] = last_t[s] = x_i[s] = z_i[s] = last_i[s] = x_m[s] = z_m[s] = last_m[s] = dp[s] = nullptr;
      a_t[s] = a_i[s] = a_m[s] = nullptr;
    }
    const bool gelu = D.activation == IISAN_ACT_GELU;
    for (int s = 0; s < D.n_stages; ++s) {
      if (D.text_adapter[s] >= 0) { x_t[s] = a.take<float>(N * D.d_text); z_t[s] = a.take<float>(N * D.r_text); last_t[s] = a.take<float>(N * D.d_text); }
      if (D.img_adapter[s] >= 0) { x_i[s] = a.take<float>(N * D.d_img); z_i[s] = a.take<float>(N * D.r_img); last_i[s] = a.take<float>(N * D.d_img); }
      if (gelu && D.text_adapter[s] >= 0) a_t[s] = a.take<float>(N * D.r_text);
      if (gelu && D.img_adapter[s] >= 0) a_i[s] = a.take<float>(N * D.r_img);
      if (gelu && D.mm_index[s] >= 0) a_m[s] = a.take<float>(N * D.r_mm);
      if (D.mm_index[s] >= 0) {
        x_m[s] = a.take<float>(N * D.d_mm); z_m[s] = a.take<float>(N * D.r_mm); last_m[s] = a.take<float>(N * D.d_mm);
        if (dimdiff) dp[s] = a.take<float>(N * D.d_mm);
      }
    }
    const int ft = D.asym ? E : D.d_text, fi = D.asym ? E : D.d_img, fm = D.d_mm;
    head_t = a.take<float>(N * ft); head_i = a.take<float>(N * fi); head_m = a.take<float>(N * fm);
    const int dmax = D.d_text > D.d_img ? D.d_text : D.d_img;
    wide_dense = dimdiff ? a.take<float>(N * dmax) : nullptr;
    dhead_t = a.take<float>(N * ft); dhead_i = a.take<float>(N * fi); dhead_m = a.take<float>(N * fm);
    dy_t = a.take<float>(N * D.d_text); dx_t = a.take<float>(N * D.d_text); dz_t = a.take<float>(N * D.r_text);
    dy_i = a.take<float>(N * D.d_img); dx_i = a.take<float>(N * D.d_img); dz_i = a.take<float>(N * D.r_img);
    dy_m = a.take<float>(N * D.d_mm); dx_m = a.take<float>(N * D.d_mm); dz_m = a.take<float>(N * D.r_mm);
    ddp = dimdiff ? a.take<float>(N * D.d_mm) : nullptr;
    bytes = a.off;
  }
};

}  // namespace iisan
