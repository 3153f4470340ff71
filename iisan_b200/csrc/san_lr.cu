// Side-adapter network, fast mode, third generation: resident-state chain forward (san_chain3.cu) + LOW-RANK ADJOINT backward.
//
// The network leaves its width-d residual stream only through rank-64 bottlenecks (AdapterBlock fc_down, CC/model/modules.py:
// 113-116; fc o pre_fc = ONE [E, d] matrix, CC/model/model.py:340-347), so with  beta_s = 1 - g_s  (intra-modal; 1 for the
// inter-modal tower),  pi(j, s) = prod_{k = s+1..j} beta_k  and stage A = the merged head (Wd_A = M = W_pre W_fc, dz_A = dL/dy):
//
//   d x_s   = sum_{j >= s} pi(j, s) dz_j Wd_j                          never materialised
//   dz_s    = [z_s > 0] * sum_{j > s} pi(j, s) dz_j (Wd_j Wu_s)        rank space: [N, 64] x [64, 64] products
//   dWd_s   = sum_{j <= s} pi(s, j) c_j (dz_s^T h_j)  +  sum_{j < s} pi(s, j) ((dz_s^T r_j) Wu_j^T + colsum(dz_s) (x) bu_j)
//   dWu_s   = sum_{j > s} pi(j, s) Wd_j^T (dz_j^T r_s) ,   dbu_s = sum_{j > s} pi(j, s) colsum(dz_j) Wd_j
//   gates   : Q_s = <dx_s, h_s> = sum_{j >= s} pi(j, s) <dz_j^T h_s, Wd_j> ;  intra-modal  d g_s = Q_s - R_{s-1} / (1 - g_s) with
//             R_s = <d last_s, last_s> = R_{s-1} + g_s Q_s - <dWd_s, Wd_s> + <dWu_s, Wu_s> + <dbu_s, bu_s>   (additive, R_{-1} = 0)
//
// i.e. the whole backward is ONE pass over the cached hidden states (G = dz^T h, a plain tensor-core GEMM with the item index as
// the reduction) plus rank-space work; no x_s / last_s / d last_s stash exists.  The algebra is restated in torch and checked
// against autograd of the oracle in float64: tests/lowrank_reference.py, tests/test_lowrank_adjoint_cpu.py.
//
// Buffers (workspace).  Per tower t (0 text, 1 image, 2 inter-modal), compact:  Rc_t [N, A, 64] relu(z_s) (forward),
// Dc_t [N, A+1, 64] dz_s with stage A = dL/dy (backward).  Per modality X (0 text, 1 image), interleaved: D_X [N, A+1, 2, 64] with
// slot 0 = the modality's own tower and slot 1 = the inter-modal tower (written into both modalities' buffers), so that ONE GEMM
// over a modality's hidden states serves both towers that read them:
//   GT_X[j] [(A+1-j)*128, d] = D_X[:, j..A]^T h^X_{layer j} ;   Pc_t [(A+1)*64, A*64] = Dc_t^T Rc_t   (fp32)
#include "san_lr.cuh"

#include <cstdlib>

#include "gemm_simt.cuh"
#include "launch.cuh"
#include "san_chain2.cuh"
#include "san_chain3.cuh"
#include "umma_gemm.cuh"

namespace iisan {

using bf16 = __nv_bfloat16;
constexpr int LE = 64;                    // rank of every bottleneck handled here (adapters and merged head)
constexpr int LMAXA = kChainMaxStages;    // 8

extern thread_local int g_umma_launch_class;
int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_inner, int box_outer);


bool san_lr_eligible(const iisan_san_desc& D) {
  // d == 768: the only width the Code_Cached tree (the non-asym heads this path merges) can have (CC/model/model.py:260,301-302);
  // the kernels themselves take any even number of 64-column chunks up to 12, which nothing reachable exercises
  if (D.d_text != D.d_img || D.d_text != 768 || !chain3_shape_supported(D.d_text, D.emb)) return false;
  if (D.r_text != 64 || D.r_img != 64 || D.r_mm != 64 || D.asym) return false;
  if (D.state_dtype != IISAN_BF16 || D.remove_first || D.n_stages > kChainMaxStages || D.n_stages < 1) return false;
  for (int s = 0; s < D.n_stages; ++s)
    if (D.text_adapter[s] < 0 || D.img_adapter[s] < 0 || D.mm_index[s] < 0) return false;
  return true;
}

bool san_lr_usable(const iisan_san_desc& D, const iisan_san_params& P) {
  if (!san_lr_eligible(D)) return false;
  uintptr_t al = 0;
  auto add = [&](const void* p) { al |= reinterpret_cast<uintptr_t>(p); };
  for (int s = 0; s < D.n_stages; ++s) {
    const iisan_adapter_ptrs* ad[3] = {&P.text[D.text_adapter[s]], &P.img[D.img_adapter[s]], &P.mm[D.mm_index[s]]};
    for (int t = 0; t < 3; ++t) { add(ad[t]->w_down); add(ad[t]->b_down); add(ad[t]->w_up); add(ad[t]->b_up); }
  }
  const iisan_linear_ptrs* lin[6] = {&P.fc_text, &P.fc_img, &P.fc_mm, &P.pre_text, &P.pre_img, &P.mm_down};
  for (int i = 0; i < 6; ++i) { add(lin[i]->w); add(lin[i]->b); }
  return (al & 15) == 0;
}

struct LrLayout {
  bf16 *wd_pack[3], *wu_pack[3], *wu_rows[3], *fcb[3], *preb[3];
  float *M32[3], *hb[3], *bsc[3];     // bsc: up-projection biases times the gate factor of the next fusion [A, d]
  bf16 *Rc[3], *Dc[3], *Dst[2];   // relu(z) [N, A, 64] and dz [N, A+1, 64] per tower (compact) ; dz interleaved per modality [N, A+1, 2, 64]
  float* KK[3];
  bf16* KB[3];                   // rank-space B operands: [32 A (A+1), 64] per tower
  float* tab;                    // GateTab
  float* zero_begin;
  float *GT[2], *Pc[3], *cs[2], *dWd[3], *dWu[3], *scal;
  size_t zero_bytes;
  bf16 *Pd[3], *Pu[3], *dMb[3];
  size_t gt_off[LMAXA];          // element offset of GT_X[j] inside GT[X]
  size_t bytes;

  LrLayout(const iisan_san_desc& D, void* ws) {
    Arena a(ws);
    const size_t A = D.n_stages, d = D.d_mm, N = D.n_items;
    const size_t f = D.d_mm;     // CC heads: fc d -> d
    for (int t = 0; t < 3; ++t) {
      wd_pack[t] = a.take<bf16>((A + 1) * LE * d); wu_pack[t] = a.take<bf16>(A * d * LE); wu_rows[t] = a.take<bf16>(d * A * LE);
      fcb[t] = a.take<bf16>(f * d); preb[t] = a.take<bf16>(LE * f);
      M32[t] = a.take<float>(LE * d); hb[t] = a.take<float>(LE); bsc[t] = a.take<float>(A * d);
    }
    for (int t = 0; t < 3; ++t) { Rc[t] = a.take<bf16>(N * A * LE); Dc[t] = a.take<bf16>(N * (A + 1) * LE); }
    for (int x = 0; x < 2; ++x) Dst[x] = a.take<bf16>(N * (A + 1) * 128);
    for (int t = 0; t < 3; ++t) KK[t] = a.take<float>((A + 1) * LE * A * LE);
    size_t gt = 0;
    for (size_t s = 0; s < A; ++s) { gt_off[s] = gt; gt += 128 * (A + 1 - s) * d; }
    for (int t = 0; t < 3; ++t) KB[t] = a.take<bf16>(32 * A * (A + 1) * LE);
    tab = a.take<float>(512);
    const size_t z0 = a.off;
    zero_begin = reinterpret_cast<float*>(a.base + a.off);
    for (int x = 0; x < 2; ++x) { GT[x] = a.take<float>(gt); cs[x] = a.take<float>((A + 1) * 128); }
    for (int t = 0; t < 3; ++t) Pc[t] = a.take<float>((A + 1) * LE * A * LE);
    for (int t = 0; t < 3; ++t) { dWd[t] = a.take<float>((A + 1) * LE * d); dWu[t] = a.take<float>(A * d * LE); }
    scal = a.take<float>(256);
    zero_bytes = a.off - z0;
    for (int t = 0; t < 3; ++t) { Pd[t] = a.take<bf16>((A + 1) * LE * A * LE); Pu[t] = a.take<bf16>(A * (A + 1) * LE * LE); dMb[t] = a.take<bf16>(LE * d); }
    bytes = a.off;
  }
};

size_t san_lr_workspace_bytes(const iisan_san_desc& D) {
  LrLayout L(D, nullptr);
  return L.bytes;
}

// ------------------------------------------------------------------------------------------------
// per-tower parameter views (tower 0 text, 1 image, 2 inter-modal)
// ------------------------------------------------------------------------------------------------
struct LrTowerPtrs {
  const float* wd[LMAXA]; const float* wu[LMAXA]; const float* bu[LMAXA]; const float* gate[LMAXA];
  const float *w_fc, *b_fc, *w_pre, *b_pre;
};
struct LrTowerGrads {
  float* wd[LMAXA]; float* bd[LMAXA]; float* wu[LMAXA]; float* bu[LMAXA]; float* gate[LMAXA];
  float *w_fc, *b_fc, *w_pre, *b_pre;
};

static void tower_ptrs(const iisan_san_desc& D, const iisan_san_params& P, int t, LrTowerPtrs* o) {
  for (int s = 0; s < D.n_stages; ++s) {
    const iisan_adapter_ptrs& ad = t == 0 ? P.text[D.text_adapter[s]] : (t == 1 ? P.img[D.img_adapter[s]] : P.mm[D.mm_index[s]]);
    o->wd[s] = ad.w_down; o->wu[s] = ad.w_up; o->bu[s] = ad.b_up;
    o->gate[s] = t == 0 ? P.gate_text[D.text_adapter[s]] : (t == 1 ? P.gate_img[D.img_adapter[s]] : P.gate_mm[D.mm_index[s]]);
  }
  const iisan_linear_ptrs& fc = t == 0 ? P.fc_text : (t == 1 ? P.fc_img : P.fc_mm);
  const iisan_linear_ptrs& pre = t == 0 ? P.pre_text : (t == 1 ? P.pre_img : P.mm_down);
  o->w_fc = fc.w; o->b_fc = fc.b; o->w_pre = pre.w; o->b_pre = pre.b;
}
static void tower_grads(const iisan_san_desc& D, const iisan_san_params& G, int t, LrTowerGrads* o) {
  for (int s = 0; s < D.n_stages; ++s) {
    const iisan_adapter_ptrs& ad = t == 0 ? G.text[D.text_adapter[s]] : (t == 1 ? G.img[D.img_adapter[s]] : G.mm[D.mm_index[s]]);
    o->wd[s] = ad.w_down; o->bd[s] = ad.b_down; o->wu[s] = ad.w_up; o->bu[s] = ad.b_up;
    o->gate[s] = t == 0 ? G.gate_text[D.text_adapter[s]] : (t == 1 ? G.gate_img[D.img_adapter[s]] : G.gate_mm[D.mm_index[s]]);
  }
  const iisan_linear_ptrs& fc = t == 0 ? G.fc_text : (t == 1 ? G.fc_img : G.fc_mm);
  const iisan_linear_ptrs& pre = t == 0 ? G.pre_text : (t == 1 ? G.pre_img : G.mm_down);
  o->w_fc = fc.w; o->b_fc = fc.b; o->w_pre = pre.w; o->b_pre = pre.b;
}
static const iisan_adapter_ptrs& tower_adapter(const iisan_san_desc& D, const iisan_san_params& P, int t, int s) {
  return t == 0 ? P.text[D.text_adapter[s]] : (t == 1 ? P.img[D.img_adapter[s]] : P.mm[D.mm_index[s]]);
}
__host__ __device__ static inline int tower_out_col(int t, int E) { return t == 0 ? E : (t == 1 ? 0 : 2 * E); }   // torch.cat((cv, text, mm)): CC/model/model.py:72

// ------------------------------------------------------------------------------------------------
// forward preparation: bf16 operand copies (one launch) + merged head bias
// ------------------------------------------------------------------------------------------------
struct LrCastJob { const float* src; bf16* dst; int rows, cols; int64_t ld_dst; };
constexpr int kLrCastJobs = 72;
struct LrPrepArgs {
  LrCastJob j[kLrCastJobs]; int n;
  const float* w_pre[3]; const float* b_fc[3]; const float* b_pre[3]; float* hb[3]; int f;
  const float* gate[3][LMAXA]; const float* bu[3][LMAXA]; float* bsc[3]; int A, d;
};

__global__ void __launch_bounds__(256) lr_prep_fwd_kernel(const __grid_constant__ LrPrepArgs a) {
  if ((int)blockIdx.y >= a.n + 3) {          // (1 - g_{s+1}) b_up_s for the intra-modal towers (the fusion that consumes last_s), else b_up_s
    const int i = blockIdx.y - a.n - 3, t = i / a.A, s = i % a.A;
    if (blockIdx.x != 0) return;
    const float coef = (t < 2 && s + 1 < a.A) ? 1.0f - gate_value(a.gate[t][s + 1]) : 1.0f;
    for (int k = threadIdx.x; k < a.d; k += 256) a.bsc[t][(size_t)s * a.d + k] = coef * a.bu[t][s][k];
    return;
  }
  if ((int)blockIdx.y >= a.n) {              // merged head bias  c = W_pre b_fc + b_pre  (blocks 0..7 of a tower, one warp per output)
    const int t = blockIdx.y - a.n;
    if (blockIdx.x >= LE / 8) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, e = blockIdx.x * 8 + warp;
    const float* w = a.w_pre[t] + (int64_t)e * a.f;
    const float* b = a.b_fc[t];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = lane;
    for (; k + 96 < a.f; k += 128) { s0 += w[k] * b[k]; s1 += w[k + 32] * b[k + 32]; s2 += w[k + 64] * b[k + 64]; s3 += w[k + 96] * b[k + 96]; }
    for (; k < a.f; k += 32) s0 += w[k] * b[k];
    const float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) a.hb[t][e] = s + a.b_pre[t][e];
    return;
  }
  const LrCastJob& J = a.j[blockIdx.y];
  const uint32_t n4 = (uint32_t)(J.rows * J.cols) / 4u, cols = (uint32_t)J.cols;          // cols % 4 == 0 (d, 64)
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n4; i += gridDim.x * 256u) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(J.src) + i);
    const uint32_t e = i * 4u, r = e / cols, c = e - r * cols;
    uint2 q;
    *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(J.dst + (int64_t)r * J.ld_dst + c) = q;
  }
}

int san_lr_forward(const iisan_san_desc* D, const iisan_san_params* P, const void* image, const void* text, void* lr_ws, float* out,
                   cudaStream_t st) {
  LrLayout L(*D, lr_ws);
  const int N = D->n_items, A = D->n_stages, d = D->d_mm, E = D->emb, f = D->d_mm;
  static thread_local LrPrepArgs pa;
  pa.n = 0; pa.f = f;
  LrTowerPtrs tp[3];
  for (int t = 0; t < 3; ++t) {
    tower_ptrs(*D, *P, t, &tp[t]);
    for (int s = 0; s < A; ++s) {
      pa.j[pa.n++] = LrCastJob{tp[t].wd[s], L.wd_pack[t] + (size_t)s * LE * d, LE, d, d};
      pa.j[pa.n++] = LrCastJob{tp[t].wu[s], L.wu_pack[t] + (size_t)s * d * LE, d, LE, LE};
      pa.j[pa.n++] = LrCastJob{tp[t].wu[s], L.wu_rows[t] + (size_t)s * LE, d, LE, (int64_t)A * LE};
    }
    pa.j[pa.n++] = LrCastJob{tp[t].w_fc, L.fcb[t], f, d, d};
    pa.j[pa.n++] = LrCastJob{tp[t].w_pre, L.preb[t], E, f, f};
    pa.w_pre[t] = tp[t].w_pre; pa.b_fc[t] = tp[t].b_fc; pa.b_pre[t] = tp[t].b_pre; pa.hb[t] = L.hb[t];
    for (int s = 0; s < A; ++s) { pa.gate[t][s] = tp[t].gate[s]; pa.bu[t][s] = tp[t].bu[s]; }
    pa.bsc[t] = L.bsc[t];
  }
  pa.A = A; pa.d = d;
  if (pa.n > kLrCastJobs) return IISAN_EINVAL;
  { LaunchScope ls_(IISAN_K_MISC, st); lr_prep_fwd_kernel<<<dim3(48, pa.n + 3 + 3 * A), 256, 0, st>>>(pa); }
  IISAN_LAUNCH_OK();
  // ---- merged head  M = W_pre W_fc  (bf16 operands, fp32 accumulation): fp32 copy for the backward, bf16 as stage A of wd_pack ----
  {
    UmmaBatch hb{}; hb.n = 3;
    for (int t = 0; t < 3; ++t) {
      UmmaProblem& p = hb.p[t];
      p.A = UmmaOperand{L.preb[t], E, f, f};
      p.B = UmmaOperand{L.fcb[t], f, d, d};
      p.a_mn_major = 0; p.b_mn_major = 1;
      p.M = E; p.N = d; p.K = f; p.splitk = 1;
      p.epi.out_f32 = L.M32[t]; p.epi.ld_f32 = d; p.epi.out_bf16 = L.wd_pack[t] + (size_t)A * LE * d; p.epi.ld_bf16 = d;
    }
    IISAN_TRY(launch_umma_gemm(hb, st));
  }
  // ---- all stages of all three towers + the heads in one launch ----
  Chain3Args ca{};
  ca.out = out; ca.out_ld = D->out_ld; ca.n_items = N; ca.d = d; ca.n_stages = A;
  // ONE MMA-issuing thread per CTA by default.  Issuing tcgen05.mma from three threads of a CTA (down-projections, up-projections
  // of the even / odd chunks) is 12 us faster (76 vs 88 us) but produced an intermittent device fault: 4 of 4 data-parallel runs on
  // 8 x B200 and 1 of ~17 single-GPU bench processes failed with it, the same 8-GPU run passed with one issuer
  // (profiles/r02_8gpu_gen3_fault.md).  IISAN_B200_C3_ONE_ISSUER=0 selects the three-issuer schedule for measurements.
  static const int one_issuer = [] { const char* e = getenv("IISAN_B200_C3_ONE_ISSUER"); return (e && e[0] == '0') ? 0 : 1; }();
  ca.one_issuer = one_issuer;
  for (int t = 0; t < 3; ++t) {
    Chain3Tower& T = ca.tower[t];
    T.mode = t == 2 ? 1 : 0;
    const void* h = t == 0 ? text : image;
    const int64_t pitch = t == 0 ? (int64_t)D->layers_text * D->d_text : (int64_t)D->layers_img * D->d_img;
    IISAN_TRY(make_tensor_map_bf16(&T.map_h, h, N, pitch, pitch, 64, 128));
    if (t == 2) IISAN_TRY(make_tensor_map_bf16(&T.map_h2, text, N, (int64_t)D->layers_text * D->d_text, (int64_t)D->layers_text * D->d_text, 64, 128));
    else T.map_h2 = T.map_h;
    IISAN_TRY(make_tensor_map_bf16(&T.map_wd, L.wd_pack[t], (int64_t)(A + 1) * LE, d, d, 64, 64));
    IISAN_TRY(make_tensor_map_bf16(&T.map_wu, L.wu_pack[t], (int64_t)A * d, LE, LE, 64, 64));
    for (int s = 0; s < A; ++s) {
      const iisan_adapter_ptrs& ad = tower_adapter(*D, *P, t, s);
      T.layer[s] = t == 0 ? D->text_layer[s] : D->img_layer[s];
      T.layer2[s] = D->text_layer[s];
      T.gate[s] = tp[t].gate[s]; T.b_down[s] = ad.b_down; T.b_up[s] = L.bsc[t] + (size_t)s * d;
    }
    T.b_down[A] = L.hb[t];
    T.r_out = L.Rc[t];
    T.out_col = tower_out_col(t, E);
  }
  return launch_san_chain3_fwd(ca, 3, st);
}

// ------------------------------------------------------------------------------------------------
// gate table (device memory, written once per backward): g[t][s] and pi[t][j][s] = prod_{k = s+1..j} beta_k
// (j >= s ; beta_k = 1 - g_k for the intra-modal towers, 1 for the inter-modal one ; beta_A = 1)
// ------------------------------------------------------------------------------------------------
struct GateTab { float g[3][LMAXA + 1]; float pi[3][LMAXA + 1][LMAXA + 1]; };
struct GatePtrs { const float* p[3][LMAXA]; int A; };

__global__ void __launch_bounds__(256) lr_coef_kernel(const __grid_constant__ GatePtrs gp, GateTab* __restrict__ tab) {
  __shared__ float beta[3][LMAXA + 1];
  const int A = gp.A;
  if (threadIdx.x < 3 * (LMAXA + 1)) {
    const int t = threadIdx.x / (LMAXA + 1), s = threadIdx.x % (LMAXA + 1);
    float g = 0.f, b = 1.0f;
    if (s < A) { g = gate_value(gp.p[t][s]); b = (t == 2) ? 1.0f : 1.0f - g; }
    tab->g[t][s] = g; beta[t][s] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * (LMAXA + 1) * (LMAXA + 1); i += 256) {
    const int t = i / ((LMAXA + 1) * (LMAXA + 1)), r = i % ((LMAXA + 1) * (LMAXA + 1)), j = r / (LMAXA + 1), s = r % (LMAXA + 1);
    float v = (j >= s && j <= A) ? 1.0f : 0.f;
    for (int k = s + 1; k <= j && k <= A; ++k) v *= beta[t][k];
    tab->pi[t][j][s] = v;
  }
}

__global__ void __launch_bounds__(256) lr_zero_kernel(uint4* __restrict__ p, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// KB_t: the B operands of the rank-space chain.  Step j (A..1) multiplies dz_j [128 x 64] with the 64 j rows starting at row
// 32 j (j - 1):  KB_t[32 j (j-1) + s*64 + b][a] = pi_t(j, s) (Wd_j Wu_s)[a, b]   (s < j ; K-major: the reduction index a is contiguous)
struct LrKbArgs { const GateTab* tab; const float* KK[3]; bf16* KB[3]; int A; };
__global__ void __launch_bounds__(256) lr_kb_kernel(const __grid_constant__ LrKbArgs a) {
  const int A = a.A, t = blockIdx.z, j = blockIdx.y + 1;
  const int n = j * LE * LE;                                   // elements of this step's block
  const float* KK = a.KK[t];
  bf16* dst = a.KB[t] + (size_t)32 * j * (j - 1) * LE;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const int aa = i & 63, b = (i >> 6) & 63, s = i >> 12;
    dst[i] = __float2bfloat16_rn(a.tab->pi[t][j][s] * KK[(size_t)(s * LE + b) * ((A + 1) * LE) + j * LE + aa]);    // KK is stored transposed
  }
}

// ------------------------------------------------------------------------------------------------
// rank-space chain (tcgen05).  One CTA = 128 rows of one tower.  dz_A = dL/dy ; for j = A..1:
//   acc[:, s-block] += dz_j (pi(j, s) Wd_j Wu_s)   for ALL s < j in one MMA group (N = 64 j, K = 64, accumulators of all stages
//   live in tensor memory: 64 A <= 448 columns) ; then dz_{j-1} = [r_{j-1} > 0] * acc[:, (j-1)-block]  -> bf16 -> the A operand of
//   the next step (shared memory) and the dz stash D_X (global; the inter-modal tower writes both modalities' copies).
// warp roles: 0 TMA producer of the weight-product blocks | 1 TMEM allocator + MMA issuer | 2..9 epilogue (quadrant = warp % 4)
// ------------------------------------------------------------------------------------------------
struct LrRankArgs {
  CUtensorMap map_kb[3];            // KB_t as a [32 A (A+1), 64] bf16 matrix, boxes of 64 x 64
  const float* d_out; int64_t ld_out; int out_col[3];
  const bf16* R[3]; bf16* Dc[3]; bf16* D[3]; bf16* D2; int slot[3];     // R / Dc: compact per tower ; D (+ D2 for the inter-modal tower): interleaved per modality
  int n_items, A;
};
constexpr int RK_THREADS = 320;
constexpr int RK_A_BYTES = 128 * 64 * 2;          // 16 KB
constexpr int RK_B_BYTES = LMAXA * 64 * 64 * 2;   // up to 64 KB per step
constexpr int RK_SMEM = 2 * RK_A_BYTES + 2 * RK_B_BYTES + 1024 + 256;

__global__ void __launch_bounds__(RK_THREADS, 1) lr_rank_kernel(const __grid_constant__ LrRankArgs a) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase, sB = sbase + 2 * RK_A_BYTES, bar0 = sB + 2 * RK_B_BYTES;
  // barriers: aFull[2] (8 epilogue warps), bFull[2] (tx), bEmpty[2] (commit), accFull (commit), tmem slot
  // The barriers do NOT start right behind the TMA destination buffers: with aFull at bar0 + 0 compute-sanitizer's synccheck reports
  // "Missing init" for it and the kernel faults under the tool (and, rarely, without it: DESIGN 4.8); 128 bytes further on both are clean.
  const uint32_t bAFull = bar0 + 256, bBFull = bar0 + 128, bBEmpty = bar0 + 384, bAcc = bar0 + 512, bTmem = bar0 + 640;
  const int t = blockIdx.y, A = a.A;
  const int m0 = blockIdx.x * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&a.map_kb[t]);
    for (int i = 0; i < 2; ++i) { c2::mbar_init_a(bAFull + 8 * i, 8); c2::mbar_init_a(bBFull + 8 * i, 1); c2::mbar_init_a(bBEmpty + 8 * i, 1); }
    c2::mbar_init_a(bAcc, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bTmem), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bTmem) : "memory");

  if (warp == 0) {
    if (elect_one()) {
      for (int j = A, n = 0; j >= 1; --j, ++n) {                 // step n uses buffer n & 1
        const int buf = n & 1;
        c2::mbar_wait_park(bBEmpty + 8 * buf, ((uint32_t)(n >> 1) & 1u) ^ 1u);
        c2::mbar_expect_tx_a(bBFull + 8 * buf, (uint32_t)(j * 64 * 64 * 2));
        for (int s = 0; s < j; ++s)
          c2::tma_load_2d_a(sB + buf * RK_B_BYTES + s * 8192, &a.map_kb[t], bBFull + 8 * buf, 0, 32 * j * (j - 1) + s * 64);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int j = A, n = 0; j >= 1; --j, ++n) {
        const int buf = n & 1;
        c2::mbar_wait_park(bBFull + 8 * buf, (uint32_t)(n >> 1) & 1u);
        c2::mbar_wait_park(bAFull + 8 * buf, (uint32_t)(n >> 1) & 1u);
        tc_fence_after();
        const uint32_t sa = sA + buf * RK_A_BYTES, sb = sB + buf * RK_B_BYTES;
        const uint32_t acc = (j == A) ? 0u : 1u;
        // columns of stage j - 1 first: the epilogue turns them into dz_{j-1}, the operand of the next step
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_bf16_ss(tmem_base + (j - 1) * 64, smem_desc_sw128(sa + kk * 32, 16, 1024), smem_desc_sw128(sb + (j - 1) * 8192 + kk * 32, 16, 1024),
                      instr_desc_bf16(128, 64, 0, 0), (acc || kk > 0) ? 1u : 0u);
        c2::mma_commit_a(bAcc);
        const int N = 64 * (j - 1), N1 = N > 256 ? 256 : N, N2 = N - N1;      // the accumulators of the stages below
        if (N1 > 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ss(tmem_base, smem_desc_sw128(sa + kk * 32, 16, 1024), smem_desc_sw128(sb + kk * 32, 16, 1024),
                        instr_desc_bf16(128, N1, 0, 0), (acc || kk > 0) ? 1u : 0u);
        }
        if (N2 > 0) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ss(tmem_base + 256, smem_desc_sw128(sa + kk * 32, 16, 1024), smem_desc_sw128(sb + 256 * 128 + kk * 32, 16, 1024),
                        instr_desc_bf16(128, N2, 0, 0), (acc || kk > 0) ? 1u : 0u);
        }
        c2::mma_commit_a(bBEmpty + 8 * buf);
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int64_t grow = (int64_t)m0 + m;
    const bool live = grow < a.n_items;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw_row = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    uint32_t offq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) offq[q] = sw_row + (uint32_t)(((half * 4 + q) ^ (m & 7)) << 4);
    const int slot = a.slot[t];
    const int64_t ldD = (int64_t)(A + 1) * 128, ldR = (int64_t)A * 64, ldC = (int64_t)(A + 1) * 64;
    // this thread's 32 columns of dz_s -> global: interleaved per modality (operand of the G GEMM) and compact per tower (Gram GEMM)
    auto stash = [&](int s, const uint32_t (&o)[16]) {
      if (!live) return;
      const int64_t off = grow * ldD + s * 128 + slot * 64 + half * 32;
      uint4* p = reinterpret_cast<uint4*>(a.D[t] + off);
      uint4* pc = reinterpret_cast<uint4*>(a.Dc[t] + grow * ldC + s * 64 + half * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) { const uint4 v = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]); p[q] = v; pc[q] = v; }
      if (t == 2) {
        uint4* p2 = reinterpret_cast<uint4*>(a.D2 + off);
#pragma unroll
        for (int q = 0; q < 4; ++q) p2[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      }
    };
    // publish this thread's 32 columns of dz_s: shared-memory operand of step `n` and the global stash
    auto publish = [&](int s, int n, const uint32_t (&o)[16]) {
      const uint32_t dstA = sA + (n & 1) * RK_A_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) c2::sts128(dstA + offq[q], o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) c2::mbar_arrive_a(bAFull + 8 * (n & 1));
      stash(s, o);
    };
    {   // dz_A = dL/dy
      uint32_t o[16];
      if (live) {
        const float4* src = reinterpret_cast<const float4*>(a.d_out + grow * a.ld_out + a.out_col[t] + half * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float4 v = src[q]; o[2 * q] = c2::pack2(v.x, v.y); o[2 * q + 1] = c2::pack2(v.z, v.w); }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) o[q] = 0u;
      }
      publish(A, 0, o);
    }
    for (int j = A, n = 0; j >= 1; --j, ++n) {
      const int s = j - 1;
      uint4 rm[4];                                               // relu(z_s) of this thread's 32 columns (the ReLU mask)
      if (live) {
        const uint4* rp = reinterpret_cast<const uint4*>(a.R[t] + grow * ldR + s * 64 + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) rm[q] = rp[q];
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) rm[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      c2::mbar_wait_a(bAcc, (uint32_t)n & 1u);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(s * 64 + half * 32), raw);
      tmem_ld_wait();
      tc_fence_before();
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(rm);
      uint32_t o[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float lo = (rw[q] & 0x7fffu) && !(rw[q] & 0x8000u) ? __uint_as_float(raw[2 * q]) : 0.f;                 // bf16 > 0
        const float hi = (rw[q] & 0x7fff0000u) && !(rw[q] & 0x80000000u) ? __uint_as_float(raw[2 * q + 1]) : 0.f;
        o[q] = c2::pack2(lo, hi);
      }
      if (s > 0) publish(s, n + 1, o);
      else stash(0, o);                                          // dz_0: only the stash
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// pi-scaled bf16 copies of the Gram blocks: Pd[t][s][a, (j, b)] = pi(s, j) P_t[(s, a), (j, b)]  (j < s, else 0) ;
// Pu[t][s][(j - s - 1, a), b] = pi(j, s) P_t[(j, a), (s, b)]  (j > s).
struct LrScaleArgs { const GateTab* tab; const float* Pc[3]; bf16* Pd[3]; bf16* Pu[3]; int A; };
__global__ void __launch_bounds__(256) lr_scale_kernel(const __grid_constant__ LrScaleArgs a) {
  const int A = a.A, t = blockIdx.z;
  const float* Pc = a.Pc[t];                // [(A+1)*64, A*64]: P_t[(j, a), (s, b)] = sum_n dz_j[n, a] relu(z_s)[n, b]
  const int ldP = A * LE;
  auto P = [&](int j, int aa, int s, int b) { return Pc[(size_t)(j * LE + aa) * ldP + s * LE + b]; };
  const int s = blockIdx.y;               // 0..A
  const int AL = A * LE;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < LE * AL; i += gridDim.x * 256) {          // Pd block s: [64, A*64]
    const int aa = i / AL, c = i - aa * AL, j = c >> 6, b = c & 63;
    const float v = (j < s) ? a.tab->pi[t][s][j] * P(s, aa, j, b) : 0.f;
    a.Pd[t][(size_t)s * LE * AL + i] = __float2bfloat16_rn(v);
  }
  if (s < A) {                            // Pu block s: [(A - s) * 64, 64]
    bf16* dst = a.Pu[t] + (size_t)s * (A + 1) * LE * LE;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < (A - s) * LE * LE; i += gridDim.x * 256) {
      const int r = i >> 6, b = i & 63, j = s + 1 + (r >> 6), aa = r & 63;
      dst[i] = __float2bfloat16_rn(a.tab->pi[t][j][s] * P(j, aa, s, b));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// combine.  The scalar reductions are added (one red.add per block and scalar) into `scal` (zeroed per backward), t stride 64:
//   [0..7] Q_j (inter-modal: from the image states), [8..15] Q2_j (inter-modal: text states), [16..23] <dWd_s, Wd_s>,
//   [24..31] <dWu_s, Wu_s>, [32..39] <dbu_s, bu_s>
// ------------------------------------------------------------------------------------------------
constexpr int LR_NPART = 1 + 2 * LMAXA;
struct LrCombineArgs {
  const GateTab* tab;
  LrTowerPtrs P[3]; LrTowerGrads G[3];
  const float* GT[2]; size_t gt_off[LMAXA];
  const float* cs[2];
  float* dWd[3]; const float* dWu[3]; const float* M32[3]; bf16* dMb[3];
  float* scal;
  int d, f, A;
};

// grid (d / 64, A + 1, 3 * 4), 256 threads: block (kt, s, t, rq) owns rows [rq*16, +16) x columns [kt*64, +64) of dWd^t_s (s == A: dM):
// adds the G terms and the rank-1 bias term to the low-rank part in the scratch, accumulates into the parameter gradient
// (s < A) or writes dM (fp32 + bf16), and reduces <dWd_s, Wd_s> and the Q partials of the gate gradients.
__global__ void __launch_bounds__(256) lr_combine_kernel(const __grid_constant__ LrCombineArgs a) {
  __shared__ float red[8][LR_NPART];
  const int A = a.A, d = a.d, t = blockIdx.z >> 2, rq = blockIdx.z & 3, s = blockIdx.y;
  const int k = blockIdx.x * 64 + (threadIdx.x & 63);
  const int a0 = rq * 16 + (threadIdx.x >> 6) * 4;
  const int xo = t == 1 ? 1 : 0, slot = t == 2 ? 1 : 0;
  const GateTab& tab = *a.tab;
  const float* Wd_s = s < A ? a.P[t].wd[s] : a.M32[t];
  const int jmax = s < A ? s : A - 1;
  float q1[LMAXA], q2[LMAXA], ipd = 0.f, pij[LMAXA], gj[LMAXA];
  float bsum = 0.f;
#pragma unroll
  for (int j = 0; j < LMAXA; ++j) {
    q1[j] = 0.f; q2[j] = 0.f;
    pij[j] = (j <= jmax) ? tab.pi[t][s][j] : 0.f;
    gj[j] = (j <= jmax) ? tab.g[t][j] : 0.f;
    if (j < s && j < A) bsum += pij[j] * a.P[t].bu[j][k];
  }
  const float* cs_s = a.cs[xo] + (s * 2 + slot) * 64;
  const float* GTo = a.GT[xo]; const float* GT0 = a.GT[0]; const float* GT1 = a.GT[1];
  float accv[4], oldg[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {             // every load of the four rows first (the gradient stores below may alias as far as the compiler knows)
    const int aa = a0 + i;
    const float w = Wd_s[(size_t)aa * d + k];
    float v1[LMAXA], v2[LMAXA];
#pragma unroll
    for (int j = 0; j < LMAXA; ++j) {
      v1[j] = 0.f; v2[j] = 0.f;
      if (j <= jmax) {
        const size_t idx = a.gt_off[j] + (size_t)(((s - j) * 2 + slot) * 64 + aa) * d + k;
        if (t < 2) v1[j] = GTo[idx];
        else { v1[j] = GT1[idx]; v2[j] = GT0[idx]; }
      }
    }
    float acc = a.dWd[t][(size_t)(s * LE + aa) * d + k] + cs_s[aa] * bsum;
    oldg[i] = s < A ? a.G[t].wd[s][(size_t)aa * d + k] : 0.f;
#pragma unroll
    for (int j = 0; j < LMAXA; ++j) {
      if (t < 2) { acc += pij[j] * gj[j] * v1[j]; q1[j] += pij[j] * v1[j] * w; }
      else { acc += pij[j] * (gj[j] * v1[j] + (1.0f - gj[j]) * v2[j]); q1[j] += pij[j] * v1[j] * w; q2[j] += pij[j] * v2[j] * w; }
    }
    accv[i] = acc;
    if (s < A) ipd += acc * w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int aa = a0 + i;
    if (s < A) a.G[t].wd[s][(size_t)aa * d + k] = oldg[i] + accv[i];
    else a.dMb[t][(size_t)aa * d + k] = __float2bfloat16_rn(accv[i]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ipd = warp_sum(ipd);
  if (lane == 0) red[warp][0] = ipd;
#pragma unroll
  for (int j = 0; j < LMAXA; ++j) {
    const float s1 = warp_sum(q1[j]), s2 = warp_sum(q2[j]);
    if (lane == 0) { red[warp][1 + j] = s1; red[warp][1 + LMAXA + j] = s2; }
  }
  __syncthreads();
  if (threadIdx.x < LR_NPART) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    float* sc = a.scal + t * 64;
    if (threadIdx.x == 0) { if (s < A) atomicAdd(sc + 16 + s, v); }
    else if (threadIdx.x <= LMAXA) { const int j = threadIdx.x - 1; if (j <= jmax) atomicAdd(sc + j, v); }
    else { const int j = threadIdx.x - 1 - LMAXA; if (t == 2 && j <= jmax) atomicAdd(sc + 8 + j, v); }
  }
}

// grad Wu_s += dWu scratch ; partial <dWu_s, Wu_s>.  grid (LR_WU_BLOCKS, A, 3), 256 threads, float4 per thread and iteration.
constexpr int LR_WU_BLOCKS = 12;
__global__ void __launch_bounds__(256) lr_wu_kernel(const __grid_constant__ LrCombineArgs a) {
  __shared__ float red[8];
  const int A = a.A, t = blockIdx.z, s = blockIdx.y;
  const int n4 = a.d * LE / 4;
  const float4* su = reinterpret_cast<const float4*>(a.dWu[t] + (size_t)s * a.d * LE);
  const float4* wu = reinterpret_cast<const float4*>(a.P[t].wu[s]);
  float* gu = a.G[t].wu[s];
  float ip = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += LR_WU_BLOCKS * 256) {
    const float4 v = su[i], w = __ldg(wu + i);
    gu[4 * i] += v.x; gu[4 * i + 1] += v.y; gu[4 * i + 2] += v.z; gu[4 * i + 3] += v.w;     // (gradient tensors: no alignment assumed)
    ip += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
  }
  ip = warp_sum(ip);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ip;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w];
    atomicAdd(a.scal + t * 64 + 24 + s, v);
  }
}

// Bias-sized pieces.  grid (d / 32, 3), 256 threads = 32 columns x 8 row groups of 8.  v_j[k] = colsum(dz_j) . Wd_j[:, k] (j = 1..A),
// dbu_s[k] = sum_{j > s} pi(j, s) v_j[k] -> grad bu_s (row group s), <dbu_s, bu_s> ; db_fc[k] += W_pre[:, k] . colsum(e) ;
// dW_pre[:, k] += colsum(e) b_fc[k] ; block x == 0 also adds colsum(dz_s) to grad bd_s and colsum(e) to db_pre.
__global__ void __launch_bounds__(256) lr_bias_kernel(const __grid_constant__ LrCombineArgs a) {
  __shared__ float csj[(LMAXA + 1) * LE];
  __shared__ float vpart[8][LMAXA + 1][32];
  const int A = a.A, d = a.d, t = blockIdx.y;
  const int xo = t == 1 ? 1 : 0, slot = t == 2 ? 1 : 0;
  const GateTab& tab = *a.tab;
  for (int i = threadIdx.x; i < (A + 1) * LE; i += 256) csj[i] = a.cs[xo][((i >> 6) * 2 + slot) * 64 + (i & 63)];
  __syncthreads();
  const int kc = threadIdx.x & 31, ag = threadIdx.x >> 5, k = blockIdx.x * 32 + kc;
  for (int j = 1; j <= A; ++j) {
    const float* W = j < A ? a.P[t].wd[j] : a.M32[t];
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int aa = ag * 8 + i; v += csj[j * LE + aa] * W[(size_t)aa * d + k]; }
    vpart[ag][j][kc] = v;
  }
  float hb = 0.f;                      // head: partial over this thread's 8 rows e of W_pre[e, k] colsum(e)[e]   (f == d)
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int e = ag * 8 + i; hb += a.P[t].w_pre[(size_t)e * a.f + k] * csj[A * LE + e]; }
  vpart[ag][0][kc] = hb;
  __syncthreads();
  if (ag < A) {                        // row group ag: stage s = ag of dbu (one warp: the inner product reduces with one warp_sum)
    const int s = ag;
    float dbu = 0.f;
    for (int j = s + 1; j <= A; ++j) {
      float vj = 0.f;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) vj += vpart[g8][j][kc];
      dbu += tab.pi[t][j][s] * vj;
    }
    a.G[t].bu[s][k] += dbu;
    const float ipb = warp_sum(dbu * a.P[t].bu[s][k]);
    if (kc == 0) atomicAdd(a.scal + t * 64 + 32 + s, ipb);
  }
  if (ag == 7) {
    float v0 = 0.f;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) v0 += vpart[g8][0][kc];
    a.G[t].b_fc[k] += v0;
  }
  {
    const float bf = a.P[t].b_fc[k];
    float old[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) old[i] = a.G[t].w_pre[(size_t)(ag * 8 + i) * a.f + k];
#pragma unroll
    for (int i = 0; i < 8; ++i) a.G[t].w_pre[(size_t)(ag * 8 + i) * a.f + k] = old[i] + csj[A * LE + ag * 8 + i] * bf;
  }
  if (blockIdx.x == 0 && threadIdx.x < LE) {
    float old[LMAXA + 1];
#pragma unroll
    for (int s = 0; s < LMAXA; ++s) if (s < A) old[s] = a.G[t].bd[s][threadIdx.x];
    old[LMAXA] = a.G[t].b_pre[threadIdx.x];
#pragma unroll
    for (int s = 0; s < LMAXA; ++s) if (s < A) a.G[t].bd[s][threadIdx.x] = old[s] + csj[s * LE + threadIdx.x];
    a.G[t].b_pre[threadIdx.x] = old[LMAXA] + csj[A * LE + threadIdx.x];
  }
}

// gate gradients from the reduced scalars (one warp)
struct LrGateArgs { const GateTab* tab; float* g_gate[3][LMAXA]; const float* scal; int A; };
__global__ void lr_gate_kernel(const __grid_constant__ LrGateArgs a) {
  const int A = a.A, t = threadIdx.x;
  if (t >= 3) return;
  const GateTab& tab = *a.tab;
  const float* sc = a.scal + t * 64;
  if (t == 2) {
    for (int s = 0; s < A; ++s) { const float g = tab.g[2][s]; atomicAdd(a.g_gate[2][s], g * (1.0f - g) / 0.1f * (sc[s] - sc[8 + s])); }
    return;
  }
  float R = 0.f;
  for (int s = 0; s < A; ++s) {
    const float g = tab.g[t][s], Q = sc[s];
    atomicAdd(a.g_gate[t][s], (g / 0.1f) * ((1.0f - g) * Q - R));
    R += g * Q - sc[16 + s] + sc[24 + s] + sc[32 + s];
  }
}

// ------------------------------------------------------------------------------------------------
// backward host side
// ------------------------------------------------------------------------------------------------
static UmmaProblem lr_linear(const bf16* x, int64_t ldx, const bf16* w, int64_t ldw, int M, int N, int K) {   // y[M,N] = x[M,K] W[N,K]^T
  UmmaProblem p{};
  p.A = UmmaOperand{x, M, K, ldx};
  p.B = UmmaOperand{w, N, K, ldw};
  p.a_mn_major = p.b_mn_major = 0;
  p.M = M; p.N = N; p.K = K; p.splitk = 1;
  return p;
}
// Split-K factor of a persistent launch over 148 SMs: the one that minimises (rounds of tiles per SM) / split, i.e. the length
// of the longest SM's work list (ties: the smaller split, fewer red.add epilogues).
static int lr_pick_split(int K, int tiles_launch) {
  const int kb = (K + 63) / 64;
  int best = 1; double best_cost = 1e30;
  for (int s = 1; s <= 6 && s <= kb / 8; ++s) {
    const int rounds = (tiles_launch * s + 147) / 148;
    const double cost = (double)rounds / s + 0.02 * s;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}
// out[m, n] (+)= y[rows, m]^T x[rows, n]   (MN-major operands; the larger of m, n becomes the UMMA M dimension)
static UmmaProblem lr_wgrad(const bf16* y, int64_t ldy, int m, const bf16* x, int64_t ldx, int n, int rows, float* out, int split) {
  UmmaProblem p{};
  p.a_mn_major = p.b_mn_major = 1;
  p.K = rows;
  if (m >= n) {
    p.A = UmmaOperand{y, rows, m, ldy}; p.B = UmmaOperand{x, rows, n, ldx};
    p.M = m; p.N = n; p.epi.transpose_out = 0;
  } else {
    p.A = UmmaOperand{x, rows, n, ldx}; p.B = UmmaOperand{y, rows, m, ldy};
    p.M = n; p.N = m; p.epi.transpose_out = 1;
  }
  p.epi.out_f32 = out; p.epi.ld_f32 = n; p.epi.atomic = 1;
  p.splitk = split;
  return p;
}

int san_lr_backward(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G, const void* image, const void* text,
                    void* lr_ws, const float* d_out, cudaStream_t st) {
  LrLayout L(*D, lr_ws);
  const int N = D->n_items, A = D->n_stages, d = D->d_mm, E = D->emb, f = D->d_mm;
  const int64_t ldD = (int64_t)(A + 1) * 128;
  GateTab* tab = reinterpret_cast<GateTab*>(L.tab);
  GatePtrs gp{}; gp.A = A;
  static thread_local LrCombineArgs ca;
  for (int t = 0; t < 3; ++t) {
    tower_ptrs(*D, *P, t, &ca.P[t]); tower_grads(*D, *G, t, &ca.G[t]);
    for (int s = 0; s < A; ++s) gp.p[t][s] = ca.P[t].gate[s];
  }
  {   // zero the accumulation scratch (split-K / red.add targets) with a kernel of our own: one more node of the same kind in a captured step
    const size_t n16 = L.zero_bytes / 16;
    LaunchScope ls_(IISAN_K_MISC, st);
    lr_zero_kernel<<<296, 256, 0, st>>>(reinterpret_cast<uint4*>(L.zero_begin), n16);
  }
  IISAN_LAUNCH_OK();
  { LaunchScope ls_(IISAN_K_MISC, st); lr_coef_kernel<<<1, 256, 0, st>>>(gp, tab); }
  IISAN_LAUNCH_OK();
  // ---- weight products  KK_t[(s, b), (j, a)] = (Wd_j Wu_s)[a, b]  for j > s ----
  {
    static thread_local UmmaBatchBig kb;
    kb.n = 0;
    for (int t = 0; t < 3; ++t)
      for (int s = 0; s < A; ++s) {
        UmmaProblem& p = kb.p[kb.n++];
        p = UmmaProblem{};
        p.A = UmmaOperand{L.wd_pack[t] + (size_t)(s + 1) * LE * d, (int64_t)(A - s) * LE, d, d};
        p.B = UmmaOperand{L.wu_pack[t] + (size_t)s * d * LE, d, LE, LE};
        p.a_mn_major = 0; p.b_mn_major = 1;
        p.M = (A - s) * LE; p.N = LE; p.K = d; p.splitk = 1;
        p.epi.out_f32 = L.KK[t] + (size_t)s * LE * ((A + 1) * LE) + (size_t)(s + 1) * LE; p.epi.ld_f32 = (int64_t)(A + 1) * LE;
        p.epi.transpose_out = 1;          // stored [(s, b), (j, a)]: lr_kb_kernel reads along a
      }
    IISAN_TRY(launch_umma_gemm_many(kb, st));
  }
  {
    LrKbArgs ka{}; ka.tab = tab; ka.A = A;
    for (int t = 0; t < 3; ++t) { ka.KK[t] = L.KK[t]; ka.KB[t] = L.KB[t]; }
    LaunchScope ls_(IISAN_K_MISC, st);
    lr_kb_kernel<<<dim3(8, A, 3), 256, 0, st>>>(ka);
  }
  IISAN_LAUNCH_OK();
  // ---- rank-space chain: dz_s of every stage and tower ----
  {
    static thread_local LrRankArgs ra;
    for (int t = 0; t < 3; ++t) {
      IISAN_TRY(make_tensor_map_bf16(&ra.map_kb[t], L.KB[t], (int64_t)32 * A * (A + 1), LE, LE, 64, 64));
      ra.out_col[t] = tower_out_col(t, E);
      ra.R[t] = L.Rc[t]; ra.Dc[t] = L.Dc[t]; ra.D[t] = L.Dst[t == 1 ? 1 : 0]; ra.slot[t] = t == 2 ? 1 : 0;
    }
    ra.D2 = L.Dst[1];
    ra.d_out = d_out; ra.ld_out = D->out_ld; ra.n_items = N; ra.A = A;
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = device_bit();
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
      IISAN_CUDA_OK(cudaFuncSetAttribute(lr_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RK_SMEM));
      attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    LaunchScope ls_(IISAN_K_CHAIN_BWD, st);
    lr_rank_kernel<<<dim3((N + 127) / 128, 3), RK_THREADS, RK_SMEM, st>>>(ra);
  }
  IISAN_LAUNCH_OK();
  // ---- column sums of dz (bias gradients, rank-1 terms) ----
  {
    ColsumBatch cb{}; cb.n = 2;
    for (int x = 0; x < 2; ++x) cb.p[x] = ColsumProb{nullptr, ldD, N, (int)ldD, L.cs[x], L.Dst[x]};
    IISAN_TRY(launch_colsum(cb, st));
  }
  // ---- the one pass over the hidden states:  GT_X[j] = D_X[:, j..A]^T h^X_layer(j) ; Gram blocks  Pg_X = D_X^T R_X ----
  {
    static thread_local UmmaBatchBig gb;
    gb.n = 0;
    int tiles = 0;
    for (int j = 0; j < A; ++j) tiles += 2 * ((d + 127) / 128) * ((128 * (A + 1 - j) + 255) / 256);
    tiles += 3 * (((A + 1) * LE + 127) / 128) * ((A * LE + 255) / 256);
    const int split = lr_pick_split(N, tiles);
    for (int x = 0; x < 2; ++x) {
      const bf16* h = reinterpret_cast<const bf16*>(x == 0 ? text : image);
      const int64_t pitch = x == 0 ? (int64_t)D->layers_text * D->d_text : (int64_t)D->layers_img * D->d_img;
      for (int j = 0; j < A; ++j) {
        const int layer = x == 0 ? D->text_layer[j] : D->img_layer[j];
        gb.p[gb.n++] = lr_wgrad(L.Dst[x] + (size_t)j * 128, ldD, 128 * (A + 1 - j), h + (size_t)layer * d, pitch, d, N, L.GT[x] + L.gt_off[j], split);
      }
    }
    for (int t = 0; t < 3; ++t)
      gb.p[gb.n++] = lr_wgrad(L.Dc[t], (int64_t)(A + 1) * LE, (A + 1) * LE, L.Rc[t], (int64_t)A * LE, A * LE, N, L.Pc[t], split);
    g_umma_launch_class = IISAN_K_WGRAD_STREAM;
    const int gst = launch_umma_gemm_many(gb, st);
    g_umma_launch_class = IISAN_K_GEMM;
    IISAN_TRY(gst);
  }
  // ---- low-rank terms of dWd / dWu ----
  {
    LrScaleArgs sa{}; sa.tab = tab; sa.A = A;
    for (int t = 0; t < 3; ++t) { sa.Pc[t] = L.Pc[t]; sa.Pd[t] = L.Pd[t]; sa.Pu[t] = L.Pu[t]; }
    LaunchScope ls_(IISAN_K_MISC, st);
    lr_scale_kernel<<<dim3(16, A + 1, 3), 256, 0, st>>>(sa);
  }
  IISAN_LAUNCH_OK();
  {
    static thread_local UmmaBatchBig lb;
    lb.n = 0;
    for (int t = 0; t < 3; ++t)
      for (int s = 1; s <= A; ++s) {      // dWd_s (low rank) = Pd_s[:, (j < s, b)] Wu_rows[:, (j < s, b)]^T
        UmmaProblem& p = lb.p[lb.n++];
        p = lr_linear(L.Pd[t] + (size_t)s * LE * A * LE, (int64_t)A * LE, L.wu_rows[t], (int64_t)A * LE, LE, d, LE * s);
        p.epi.out_f32 = L.dWd[t] + (size_t)s * LE * d; p.epi.ld_f32 = d;
      }
    IISAN_TRY(launch_umma_gemm_many(lb, st));
    lb.n = 0;
    for (int t = 0; t < 3; ++t)
      for (int s = 0; s < A; ++s)         // dWu_s = sum_{j > s} Wd_j^T Pu_s[(j, a), :]
        lb.p[lb.n++] = lr_wgrad(L.wd_pack[t] + (size_t)(s + 1) * LE * d, d, d, L.Pu[t] + (size_t)s * (A + 1) * LE * LE, LE, LE, (A - s) * LE,
                                L.dWu[t] + (size_t)s * d * LE, 1);
    IISAN_TRY(launch_umma_gemm_many(lb, st));
  }
  // ---- combine ----
  ca.tab = tab; ca.A = A;
  for (int x = 0; x < 2; ++x) { ca.GT[x] = L.GT[x]; ca.cs[x] = L.cs[x]; }
  for (int j = 0; j < A; ++j) ca.gt_off[j] = L.gt_off[j];
  for (int t = 0; t < 3; ++t) { ca.dWd[t] = L.dWd[t]; ca.dWu[t] = L.dWu[t]; ca.M32[t] = L.M32[t]; ca.dMb[t] = L.dMb[t]; }
  ca.scal = L.scal; ca.d = d; ca.f = f;
  { LaunchScope ls_(IISAN_K_MISC, st); lr_combine_kernel<<<dim3(d / 64, A + 1, 12), 256, 0, st>>>(ca); }
  IISAN_LAUNCH_OK();
  { LaunchScope ls_(IISAN_K_MISC, st); lr_wu_kernel<<<dim3(LR_WU_BLOCKS, A, 3), 256, 0, st>>>(ca); }
  IISAN_LAUNCH_OK();
  { LaunchScope ls_(IISAN_K_MISC, st); lr_bias_kernel<<<dim3(d / 32, 3), 256, 0, st>>>(ca); }
  IISAN_LAUNCH_OK();
  // ---- heads:  dW_pre += dM W_fc^T ,  dW_fc += W_pre^T dM   (bf16 operands) ----
  {
    UmmaBatch hp{}; hp.n = 3;
    for (int t = 0; t < 3; ++t) {
      hp.p[t] = lr_linear(L.dMb[t], d, L.fcb[t], d, E, f, d);
      hp.p[t].epi.out_f32 = ca.G[t].w_pre; hp.p[t].epi.ld_f32 = f; hp.p[t].epi.atomic = 1;
    }
    IISAN_TRY(launch_umma_gemm(hp, st));
    UmmaBatch hf{}; hf.n = 3;
    for (int t = 0; t < 3; ++t) hf.p[t] = lr_wgrad(L.preb[t], f, f, L.dMb[t], d, d, E, ca.G[t].w_fc, 1);
    IISAN_TRY(launch_umma_gemm(hf, st));
  }
  {
    LrGateArgs ga{}; ga.tab = tab; ga.scal = L.scal; ga.A = A;
    for (int t = 0; t < 3; ++t) for (int s = 0; s < A; ++s) ga.g_gate[t][s] = ca.G[t].gate[s];
    LaunchScope ls_(IISAN_K_MISC, st);
    lr_gate_kernel<<<1, 32, 0, st>>>(ga);
  }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan
