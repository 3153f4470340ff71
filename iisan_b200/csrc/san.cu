// Side-adapter network forward/backward (exact fp32 mode) + C-ABI entry points.
//
// Algorithm restated from the reference's PyTorch modules (nothing ported):
//   gated fusion       CC/model/model.py:319-326   x = g*h_l + (1-g)*last,  g = sigmoid(p/0.1)
//   inter-modal mix    CC/model/model.py:335-337   x = last + g*h_cv + (1-g)*h_text
//   AdapterBlock       CC/model/modules.py:113-116 last' = up(relu(down(x))) + x
//   heads              CC/model/model.py:340-347   e = pre(fc(last))
//   Versa extensions   CA/model/model.py:353-417   solo stages + down_project dim alignment
//
// HBM layout: the cached states stay in the caller's [N, layers, d] tensors and are read in place
// (row pitch layers*d, 128-bit loads of the selected layer only).  The forward stashes, per stage
// and tower, x_s [N,d], z_s [N,r] and last_s [N,d] in the workspace for the backward.
#include "common.cuh"
#include "launch.cuh"
#include "gemm_simt.cuh"
#include "san_layout.cuh"
#include "san_mix.cuh"

namespace iisan {

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
int san_validate(const iisan_san_desc* d) {
  if (!d) return IISAN_EINVAL;
  if (d->n_items <= 0 || d->n_stages <= 0 || d->n_stages > IISAN_MAX_STAGES) return IISAN_EINVAL;
  if (d->d_text % 4 || d->d_img % 4 || d->d_mm % 4 || d->emb % 4) return IISAN_EINVAL;
  if (d->d_mm != (d->d_text < d->d_img ? d->d_text : d->d_img)) return IISAN_EINVAL;
  if (d->out_ld < 3 * d->emb) return IISAN_EINVAL;
  if (d->state_dtype < IISAN_F32 || d->state_dtype > IISAN_F16) return IISAN_EINVAL;
  if (d->compute != IISAN_COMPUTE_FP32 && d->compute != IISAN_COMPUTE_BF16) return IISAN_EINVAL;
  if (d->activation != IISAN_ACT_RELU && d->activation != IISAN_ACT_GELU) return IISAN_EINVAL;
  for (int s = 0; s < d->n_stages; ++s) {
    if (d->text_adapter[s] >= 0 && (d->text_layer[s] < 0 || d->text_layer[s] >= d->layers_text)) return IISAN_EINVAL;
    if (d->img_adapter[s] >= 0 && (d->img_layer[s] < 0 || d->img_layer[s] >= d->layers_img)) return IISAN_EINVAL;
    if (d->mm_index[s] >= 0 && (d->text_adapter[s] < 0 || d->img_adapter[s] < 0)) return IISAN_EINVAL;
  }
  return IISAN_OK;
}

template <typename T>
static int san_forward_fp32(const iisan_san_desc* D, const iisan_san_params* P, const void* image,
                            const void* text, void* ws, float* out, cudaStream_t st) {
  SanLayout L(*D, ws);
  const int N = D->n_items;
  const float* last_t = nullptr; const float* last_i = nullptr; const float* last_m = nullptr;
  bool first_t = true, first_i = true;
  const int act = D->activation == IISAN_ACT_GELU ? 2 : 1;      // epilogue of the down-projection (GELU also stores the pre-activation)
  for (int s = 0; s < D->n_stages; ++s) {
    const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
    // ---- optional dim alignment GEMM (CA/model/model.py:406-411) ----
    const float* dp = nullptr;
    if (mi >= 0 && D->d_text != D->d_img) {
      const bool text_wide = D->d_text > D->d_img;
      const int dw = text_wide ? D->d_text : D->d_img;
      MixBatch gb{}; gb.n = 1;
      MixProb& g = gb.p[0];
      g.P = text_wide ? state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s])
                      : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      g.mode = 2; g.X = L.wide_dense; g.N = N; g.d = dw;
      IISAN_TRY(launch_mix<T>(gb, st));
      GemmBatch b{}; b.n = 1;
      b.p[0] = prob_linear(L.wide_dense, dw, P->down_project[mi].w, P->down_project[mi].b, L.dp[s], D->d_mm, N, D->d_mm, dw);
      IISAN_TRY(launch_gemm(b, st));
      dp = L.dp[s];
    }
    // ---- fusion ----
    MixBatch mb{}; GemmBatch down{}, up{};
    if (ta >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      if (first_t && D->remove_first) m.R = state_src<T>(text, D->layers_text, D->d_text, 0);
      else m.R = dense_src(last_t, D->d_text);
      m.gate = P->gate_text[ta]; m.mode = 0; m.X = L.x_t[s]; m.N = N; m.d = D->d_text;
      down.p[down.n] = prob_linear(L.x_t[s], D->d_text, P->text[ta].w_down, P->text[ta].b_down, L.z_t[s], D->r_text, N, D->r_text, D->d_text, act);
      down.p[down.n].pre = L.a_t[s]; down.p[down.n++].ldp = D->r_text;
      up.p[up.n++] = prob_linear(L.z_t[s], D->r_text, P->text[ta].w_up, P->text[ta].b_up, L.last_t[s], D->d_text, N, D->d_text, D->r_text, 0, L.x_t[s], D->d_text);
    }
    if (ia >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      if (first_i && D->remove_first) m.R = state_src<T>(image, D->layers_img, D->d_img, 0);
      else m.R = dense_src(last_i, D->d_img);
      m.gate = P->gate_img[ia]; m.mode = 0; m.X = L.x_i[s]; m.N = N; m.d = D->d_img;
      down.p[down.n] = prob_linear(L.x_i[s], D->d_img, P->img[ia].w_down, P->img[ia].b_down, L.z_i[s], D->r_img, N, D->r_img, D->d_img, act);
      down.p[down.n].pre = L.a_i[s]; down.p[down.n++].ldp = D->r_img;
      up.p[up.n++] = prob_linear(L.z_i[s], D->r_img, P->img[ia].w_up, P->img[ia].b_up, L.last_i[s], D->d_img, N, D->d_img, D->r_img, 0, L.x_i[s], D->d_img);
    }
    if (mi >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = (dp && D->d_img > D->d_text) ? dense_src(dp, D->d_mm) : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      m.Q = (dp && D->d_text > D->d_img) ? dense_src(dp, D->d_mm) : state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      m.R = dense_src(last_m, D->d_mm);
      m.gate = P->gate_mm[mi]; m.mode = 1; m.X = L.x_m[s]; m.N = N; m.d = D->d_mm;
      down.p[down.n] = prob_linear(L.x_m[s], D->d_mm, P->mm[mi].w_down, P->mm[mi].b_down, L.z_m[s], D->r_mm, N, D->r_mm, D->d_mm, act);
      down.p[down.n].pre = L.a_m[s]; down.p[down.n++].ldp = D->r_mm;
      up.p[up.n++] = prob_linear(L.z_m[s], D->r_mm, P->mm[mi].w_up, P->mm[mi].b_up, L.last_m[s], D->d_mm, N, D->d_mm, D->r_mm, 0, L.x_m[s], D->d_mm);
    }
    IISAN_TRY(launch_mix<T>(mb, st));
    IISAN_TRY(launch_gemm(down, st));
    IISAN_TRY(launch_gemm(up, st));
    if (ta >= 0) { last_t = L.last_t[s]; first_t = false; }
    if (ia >= 0) { last_i = L.last_i[s]; first_i = false; }
    if (mi >= 0) last_m = L.last_m[s];
  }
  if (!last_t || !last_i || !last_m) return IISAN_EINVAL;
  // ---- heads: e = pre(fc(last)) ----
  const int E = D->emb;
  const int ft = D->asym ? E : D->d_text, fi = D->asym ? E : D->d_img, fm = D->d_mm;
  GemmBatch fc{}; fc.n = 3;
  fc.p[0] = prob_linear(last_i, D->d_img, P->fc_img.w, P->fc_img.b, L.head_i, fi, N, fi, D->d_img);
  fc.p[1] = prob_linear(last_t, D->d_text, P->fc_text.w, P->fc_text.b, L.head_t, ft, N, ft, D->d_text);
  fc.p[2] = prob_linear(last_m, D->d_mm, P->fc_mm.w, P->fc_mm.b, L.head_m, fm, N, fm, D->d_mm);
  IISAN_TRY(launch_gemm(fc, st));
  GemmBatch pre{}; pre.n = 3;
  pre.p[0] = prob_linear(L.head_i, fi, P->pre_img.w, P->pre_img.b, out, D->out_ld, N, E, fi);
  pre.p[1] = prob_linear(L.head_t, ft, P->pre_text.w, P->pre_text.b, out + E, D->out_ld, N, E, ft);
  pre.p[2] = prob_linear(L.head_m, fm, P->mm_down.w, P->mm_down.b, out + 2 * E, D->out_ld, N, E, fm);
  IISAN_TRY(launch_gemm(pre, st));
  return IISAN_OK;
}

template <typename T>
static int san_backward_fp32(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G,
                             const void* image, const void* text, void* ws, const float* d_out, cudaStream_t st) {
  SanLayout L(*D, ws);
  const int N = D->n_items, E = D->emb;
  const int ft = D->asym ? E : D->d_text, fi = D->asym ? E : D->d_img, fm = D->d_mm;
  // last stage index per tower
  int ls_t = -1, ls_i = -1, ls_m = -1;
  for (int s = 0; s < D->n_stages; ++s) {
    if (D->text_adapter[s] >= 0) ls_t = s;
    if (D->img_adapter[s] >= 0) ls_i = s;
    if (D->mm_index[s] >= 0) ls_m = s;
  }
  if (ls_t < 0 || ls_i < 0 || ls_m < 0) return IISAN_EINVAL;
  // ---- heads backward ----
  {
    GemmBatch w{}; w.n = 3;  // d pre weights
    w.p[0] = prob_wgrad(d_out, D->out_ld, L.head_i, fi, G->pre_img.w, N, E, fi);
    w.p[1] = prob_wgrad(d_out + E, D->out_ld, L.head_t, ft, G->pre_text.w, N, E, ft);
    w.p[2] = prob_wgrad(d_out + 2 * E, D->out_ld, L.head_m, fm, G->mm_down.w, N, E, fm);
    IISAN_TRY(launch_gemm(w, st));
    ColsumBatch c{}; c.n = 3;
    c.p[0] = {d_out, D->out_ld, N, E, G->pre_img.b};
    c.p[1] = {d_out + E, D->out_ld, N, E, G->pre_text.b};
    c.p[2] = {d_out + 2 * E, D->out_ld, N, E, G->mm_down.b};
    IISAN_TRY(launch_colsum(c, st));
    GemmBatch dh{}; dh.n = 3;  // d head = d_out W_pre
    dh.p[0] = prob_dgrad(d_out, D->out_ld, P->pre_img.w, L.dhead_i, fi, N, E, fi);
    dh.p[1] = prob_dgrad(d_out + E, D->out_ld, P->pre_text.w, L.dhead_t, ft, N, E, ft);
    dh.p[2] = prob_dgrad(d_out + 2 * E, D->out_ld, P->mm_down.w, L.dhead_m, fm, N, E, fm);
    IISAN_TRY(launch_gemm(dh, st));
    GemmBatch wf{}; wf.n = 3;
    wf.p[0] = prob_wgrad(L.dhead_i, fi, L.last_i[ls_i], D->d_img, G->fc_img.w, N, fi, D->d_img);
    wf.p[1] = prob_wgrad(L.dhead_t, ft, L.last_t[ls_t], D->d_text, G->fc_text.w, N, ft, D->d_text);
    wf.p[2] = prob_wgrad(L.dhead_m, fm, L.last_m[ls_m], D->d_mm, G->fc_mm.w, N, fm, D->d_mm);
    IISAN_TRY(launch_gemm(wf, st));
    ColsumBatch cf{}; cf.n = 3;
    cf.p[0] = {L.dhead_i, fi, N, fi, G->fc_img.b};
    cf.p[1] = {L.dhead_t, ft, N, ft, G->fc_text.b};
    cf.p[2] = {L.dhead_m, fm, N, fm, G->fc_mm.b};
    IISAN_TRY(launch_colsum(cf, st));
    GemmBatch dl{}; dl.n = 3;  // d last = d head W_fc
    dl.p[0] = prob_dgrad(L.dhead_i, fi, P->fc_img.w, L.dy_i, D->d_img, N, fi, D->d_img);
    dl.p[1] = prob_dgrad(L.dhead_t, ft, P->fc_text.w, L.dy_t, D->d_text, N, ft, D->d_text);
    dl.p[2] = prob_dgrad(L.dhead_m, fm, P->fc_mm.w, L.dy_m, D->d_mm, N, fm, D->d_mm);
    IISAN_TRY(launch_gemm(dl, st));
  }
  // ---- stages in reverse ----
  // dy_* holds d last_s on entry to stage s and d last_{s-1} on exit (ping-pong with dx_*).
  float* dy_t = L.dy_t; float* dx_t = L.dx_t;
  float* dy_i = L.dy_i; float* dx_i = L.dx_i;
  float* dy_m = L.dy_m; float* dx_m = L.dx_m;
  for (int s = D->n_stages - 1; s >= 0; --s) {
    const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
    // previous stage of each tower
    int ps_t = -1, ps_i = -1, ps_m = -1;
    for (int q = 0; q < s; ++q) {
      if (D->text_adapter[q] >= 0) ps_t = q;
      if (D->img_adapter[q] >= 0) ps_i = q;
      if (D->mm_index[q] >= 0) ps_m = q;
    }
    GemmBatch wu{}, dz{}, wd{}, dx{}; ColsumBatch cu{}, cd{};
    MixBwdBatch mb{};
    auto add = [&](const iisan_adapter_ptrs& p, const iisan_adapter_ptrs& g, const float* x, const float* z, const float* pre,
                   float* dz_buf, const float* dy, float* dxb, int d, int r) {
      wu.p[wu.n++] = prob_wgrad(dy, d, z, r, g.w_up, N, d, r);                 // dWu += dy^T z
      cu.p[cu.n++] = {dy, d, N, d, g.b_up};
      if (pre) dz.p[dz.n++] = prob_dgrad(dy, d, p.w_up, dz_buf, r, N, d, r, pre, r, nullptr, 0, 1);   // GELU: dz = (dy Wu) * gelu'(pre)
      else dz.p[dz.n++] = prob_dgrad(dy, d, p.w_up, dz_buf, r, N, d, r, z, r);  // ReLU: dz = (dy Wu) * (z>0)
      wd.p[wd.n++] = prob_wgrad(dz_buf, r, x, d, g.w_down, N, r, d);            // dWd += dz^T x
      cd.p[cd.n++] = {dz_buf, r, N, r, g.b_down};
      dx.p[dx.n++] = prob_dgrad(dz_buf, r, p.w_down, dxb, d, N, r, d, nullptr, 0, dy, d);  // dx = dy + dz Wd
    };
    if (ta >= 0) add(P->text[ta], G->text[ta], L.x_t[s], L.z_t[s], L.a_t[s], L.dz_t, dy_t, dx_t, D->d_text, D->r_text);
    if (ia >= 0) add(P->img[ia], G->img[ia], L.x_i[s], L.z_i[s], L.a_i[s], L.dz_i, dy_i, dx_i, D->d_img, D->r_img);
    if (mi >= 0) add(P->mm[mi], G->mm[mi], L.x_m[s], L.z_m[s], L.a_m[s], L.dz_m, dy_m, dx_m, D->d_mm, D->r_mm);
    IISAN_TRY(launch_gemm(wu, st));
    IISAN_TRY(launch_colsum(cu, st));
    IISAN_TRY(launch_gemm(dz, st));
    IISAN_TRY(launch_gemm(wd, st));
    IISAN_TRY(launch_colsum(cd, st));
    IISAN_TRY(launch_gemm(dx, st));
    // fusion backward
    const bool has_dp = (mi >= 0 && D->d_text != D->d_img);
    const bool text_wide = D->d_text > D->d_img;
    if (ta >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      m.P = state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      if (ps_t < 0) { if (D->remove_first) m.R = state_src<T>(text, D->layers_text, D->d_text, 0); else m.R = MixSrc{nullptr, 0, 0}; }
      else m.R = dense_src(L.last_t[ps_t], D->d_text);
      m.gate = P->gate_text[ta]; m.dgate = G->gate_text[ta]; m.mode = 0; m.dX = dx_t;
      m.dPrev = (ps_t >= 0) ? dy_t : nullptr; m.N = N; m.d = D->d_text;
    }
    if (ia >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      m.P = state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      if (ps_i < 0) { if (D->remove_first) m.R = state_src<T>(image, D->layers_img, D->d_img, 0); else m.R = MixSrc{nullptr, 0, 0}; }
      else m.R = dense_src(L.last_i[ps_i], D->d_img);
      m.gate = P->gate_img[ia]; m.dgate = G->gate_img[ia]; m.mode = 0; m.dX = dx_i;
      m.dPrev = (ps_i >= 0) ? dy_i : nullptr; m.N = N; m.d = D->d_img;
    }
    if (mi >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      const float* dp = has_dp ? L.dp[s] : nullptr;
      m.P = (dp && !text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      m.Q = (dp && text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      m.gate = P->gate_mm[mi]; m.dgate = G->gate_mm[mi]; m.mode = 1; m.dX = dx_m;
      m.dP_out = (has_dp && !text_wide) ? L.ddp : nullptr;
      m.dQ_out = (has_dp && text_wide) ? L.ddp : nullptr;
      m.N = N; m.d = D->d_mm;
    }
    IISAN_TRY(launch_mix_bwd<T>(mb, st));
    if (mi >= 0) { float* t = dy_m; dy_m = dx_m; dx_m = t; (void)ps_m; }   // d last_mm_{s-1} = dx_mm
    if (has_dp) {
      // down_project gradients: dW += ddp^T h_wide ; db += colsum(ddp)
      const int dw = text_wide ? D->d_text : D->d_img;
      MixBatch gb{}; gb.n = 1;
      MixProb& g = gb.p[0];
      g.P = text_wide ? state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s])
                      : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      g.mode = 2; g.X = L.wide_dense; g.N = N; g.d = dw;
      IISAN_TRY(launch_mix<T>(gb, st));
      GemmBatch w{}; w.n = 1;
      w.p[0] = prob_wgrad(L.ddp, D->d_mm, L.wide_dense, dw, G->down_project[mi].w, N, D->d_mm, dw);
      IISAN_TRY(launch_gemm(w, st));
      ColsumBatch c{}; c.n = 1;
      c.p[0] = {L.ddp, D->d_mm, N, D->d_mm, G->down_project[mi].b};
      IISAN_TRY(launch_colsum(c, st));
    }
  }
  return IISAN_OK;
}

}  // namespace iisan

using namespace iisan;

extern "C" size_t iisan_san_workspace_bytes(const iisan_san_desc* desc) {
  if (san_validate(desc) != IISAN_OK) return 0;
  if (desc->compute == IISAN_COMPUTE_BF16) return san_bf16_supported(*desc) ? san_bf16_workspace_bytes(*desc) : 0;
  SanLayout L(*desc, nullptr);
  return L.bytes;
}

extern "C" int iisan_san_fused_eligible(const iisan_san_desc* desc) {
  if (san_validate(desc) != IISAN_OK || desc->compute != IISAN_COMPUTE_BF16 || !san_bf16_supported(*desc)) return 0;
  return san_chain_eligible_if_bf16(*desc);
}

extern "C" int iisan_san_forward(const iisan_san_desc* desc, const iisan_san_params* params, const void* image,
                                 const void* text, void* workspace, size_t workspace_bytes, float* out,
                                 iisan_stream_t stream) {
  IISAN_TRY(san_validate(desc));
  if (!params || !image || !text || !workspace || !out) return IISAN_EINVAL;
  if (desc->compute == IISAN_COMPUTE_BF16 && !san_bf16_supported(*desc)) return IISAN_EUNSUPPORTED;
  if (workspace_bytes < iisan_san_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (desc->compute == IISAN_COMPUTE_BF16) return san_forward_bf16(desc, params, image, text, workspace, out, st);
  switch (desc->state_dtype) {
    case IISAN_F32: return san_forward_fp32<float>(desc, params, image, text, workspace, out, st);
    case IISAN_BF16: return san_forward_fp32<__nv_bfloat16>(desc, params, image, text, workspace, out, st);
    case IISAN_F16: return san_forward_fp32<__half>(desc, params, image, text, workspace, out, st);
  }
  return IISAN_EINVAL;
}

extern "C" int iisan_san_backward(const iisan_san_desc* desc, const iisan_san_params* params,
                                  const iisan_san_params* grads, const void* image, const void* text,
                                  void* workspace, size_t workspace_bytes, const float* d_out,
                                  iisan_stream_t stream) {
  IISAN_TRY(san_validate(desc));
  if (!params || !grads || !image || !text || !workspace || !d_out) return IISAN_EINVAL;
  if (desc->compute == IISAN_COMPUTE_BF16 && !san_bf16_supported(*desc)) return IISAN_EUNSUPPORTED;
  if (workspace_bytes < iisan_san_workspace_bytes(desc)) return IISAN_EWORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (desc->compute == IISAN_COMPUTE_BF16) return san_backward_bf16(desc, params, grads, image, text, workspace, d_out, st);
  switch (desc->state_dtype) {
    case IISAN_F32: return san_backward_fp32<float>(desc, params, grads, image, text, workspace, d_out, st);
    case IISAN_BF16: return san_backward_fp32<__nv_bfloat16>(desc, params, grads, image, text, workspace, d_out, st);
    case IISAN_F16: return san_backward_fp32<__half>(desc, params, grads, image, text, workspace, d_out, st);
  }
  return IISAN_EINVAL;
}
