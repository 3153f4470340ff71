// Side-adapter network, fast mode (IISAN_COMPUTE_BF16): every contraction runs on the tcgen05 GEMM primitive
// (umma_gemm.cu) with bf16 operands and fp32 accumulation in TMEM; elementwise stages emit bf16 operands.
// This is the general ("layered") fast path: any widths / bottlenecks / stage plans (IISAN-Versa included).
//
// Algorithm restated from the reference's PyTorch modules (nothing ported):
//   gated fusion       CC/model/model.py:319-326 ; inter-modal mix CC/model/model.py:335-337
//   AdapterBlock       CC/model/modules.py:113-116 ; heads CC/model/model.py:340-347
//   Versa extensions   CA/model/model.py:353-417 (solo stages, down_project dim alignment)
//
// HBM layout (workspace): bf16 copies of every weight in both orientations ([out,in] for forward, [in,out] for the data
// gradients), refreshed by one cast+transpose launch per forward; per stage and tower the stash x_s [N,d], z_s [N,r],
// last_s [N,d] (all bf16); backward scratch dy/dx (fp32 running gradient + bf16 operand twins), dz (bf16).
// bf16 cached states are consumed in place by TMA (row pitch layers*d) wherever they are a GEMM operand.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "launch.cuh"
#include "san_layout.cuh"
#include "san_lr.cuh"
#include "san_chain.cuh"
#include "san_mix.cuh"
#include "umma_gemm.cuh"

namespace iisan {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------------
// weight preparation: dst[r,c] = bf16(src[r,c]) and dstT[c,r] = bf16(src[r,c])
// ------------------------------------------------------------------------------------------------
struct CastJob { const float* src; bf16* dst; bf16* dstT; int rows, cols; };
constexpr int kCastJobs = 48;
struct CastBatch { CastJob j[kCastJobs]; int n; };

__global__ void __launch_bounds__(256) cast_transpose_kernel(const CastBatch batch) {
  const CastJob& J = batch.j[blockIdx.y];
  if (!J.dstT) {                               // plain cast: 128-bit loads, 64-bit stores
    const int64_t n = (int64_t)J.rows * J.cols;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(J.src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(J.dst) & 7) == 0)) {
      for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n / 4; i += (int64_t)gridDim.x * 256) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(J.src) + i);
        uint2 q;
        *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(J.dst)[i] = q;
      }
    } else {
      for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) J.dst[i] = __float2bfloat16_rn(J.src[i]);
    }
    return;
  }
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int tiles_c = (J.cols + 31) / 32, tiles_r = (J.rows + 31) / 32;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + k * 8, c = c0 + tx;
      float v = 0.f;
      if (r < J.rows && c < J.cols) {
        v = J.src[(int64_t)r * J.cols + c];
        if (J.dst) J.dst[(int64_t)r * J.cols + c] = __float2bfloat16_rn(v);
      }
      tile[ty + k * 8][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + k * 8, r = r0 + tx;
      if (J.dstT && r < J.rows && c < J.cols) J.dstT[(int64_t)c * J.rows + r] = __float2bfloat16_rn(tile[tx][ty + k * 8]);
    }
    __syncthreads();
  }
}

struct CastList {
  CastBatch b; cudaStream_t st; int status;
  explicit CastList(cudaStream_t s) : st(s), status(IISAN_OK) { b.n = 0; }
  void flush() {
    if (b.n == 0 || status != IISAN_OK) { b.n = 0; return; }
    int mx = 1;
    for (int i = 0; i < b.n; ++i) mx = max(mx, ((b.j[i].rows + 31) / 32) * ((b.j[i].cols + 31) / 32));
    { LaunchScope ls_(IISAN_K_MISC, st); cast_transpose_kernel<<<dim3(min(mx, 64), b.n), 256, 0, st>>>(b); }
    if (cudaPeekAtLastError() != cudaSuccess) status = cuda_fail(cudaGetLastError());
    b.n = 0;
  }
  void add(const float* src, bf16* dst, bf16* dstT, int rows, int cols) {
    if (!src) { status = IISAN_EINVAL; return; }
    b.j[b.n++] = CastJob{src, dst, dstT, rows, cols};
    if (b.n == kCastJobs) flush();
  }
};

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
// The fused chain kernel (san_chain.cu) covers the symmetric configurations: equal widths, r = 64, bf16 cached states, every
// tower active in every stage (no group layer-drop), towers starting from zero.
// d >= 640 (ten 64-column chunks): the kernels re-read their own stash one stage after writing it, ordered by the distance
// between the write and the re-read (san_chain.cu header); narrower states fall back to the layered path.
static bool san_chain_eligible(const iisan_san_desc& D) {
  if (D.d_text != D.d_img || D.d_text % 64 || D.d_text < 640 || D.r_text != 64 || D.r_img != 64 || D.r_mm != 64) return false;
  if (D.state_dtype != IISAN_BF16 || D.remove_first || D.n_stages > kChainMaxStages) return false;
  if (D.activation != IISAN_ACT_RELU) return false;          // the chain kernels and the low-rank adjoint are built on ReLU (mask = relu(z) > 0)
  for (int s = 0; s < D.n_stages; ++s)
    if (D.text_adapter[s] < 0 || D.img_adapter[s] < 0 || D.mm_index[s] < 0) return false;
  return true;
}

// Would this configuration take the fused chain if its states were stored in bf16?  (The host packs + casts fp32 / fp16 states
// once per step -- iisan_pack_states -- when the answer is yes: ops.SanFn.)
int san_chain_eligible_if_bf16(const iisan_san_desc& D) {
  iisan_san_desc B = D;
  B.state_dtype = IISAN_BF16;
  return san_chain_eligible(B) ? 1 : 0;
}

struct WCopy { bf16* w; bf16* wt; };

struct SanLayoutBf16 {
  WCopy t_down[IISAN_MAX_STAGES], t_up[IISAN_MAX_STAGES], i_down[IISAN_MAX_STAGES], i_up[IISAN_MAX_STAGES];
  WCopy m_down[IISAN_MAX_STAGES], m_up[IISAN_MAX_STAGES], dpw[IISAN_MAX_STAGES];
  WCopy fc_t, fc_i, fc_m, pre_t, pre_i, pre_m;
  bf16 *x_t[IISAN_MAX_STAGES], *z_t[IISAN_MAX_STAGES], *last_t[IISAN_MAX_STAGES];
  bf16 *x_i[IISAN_MAX_STAGES], *z_i[IISAN_MAX_STAGES], *last_i[IISAN_MAX_STAGES];
  bf16 *x_m[IISAN_MAX_STAGES], *z_m[IISAN_MAX_STAGES], *last_m[IISAN_MAX_STAGES];
  float* dpo[IISAN_MAX_STAGES];          // down_project output [N, d_mm] (Versa)
  float *a_t[IISAN_MAX_STAGES], *a_i[IISAN_MAX_STAGES], *a_m[IISAN_MAX_STAGES];      // GELU: fp32 pre-activations [N, r] (null for ReLU)
  bf16 *head_t, *head_i, *head_m;
  bf16* wide_b;                          // [N, max d] bf16 copy of one wide layer when the states are not bf16
  // backward scratch
  bf16* doutb;
  bf16 *dhead_t, *dhead_i, *dhead_m;
  float *dy_t, *dx_t, *dy_i, *dx_i, *dy_m, *dx_m;
  bf16 *dyb_t, *dxb_t, *dzb_t, *dyb_i, *dxb_i, *dzb_i, *dyb_m, *dxb_m, *dzb_m;
  bf16* ddpb;
  bf16 *wd_pack[3], *wu_pack[3];          // fused chain: per tower (text, img, mm) all stages' weights, contiguous
  bf16* dys[3];                           // fused chain backward: d last_s of all stages [A, N, d] per tower
  bf16* dzs[3][IISAN_MAX_STAGES];         //                        dz_s [N, r]
  void* lr_ws;                            // third generation (san_lr.cu): its own block of the workspace
  size_t bytes;

  static WCopy takew(Arena& a, size_t n) { WCopy c; c.w = a.take<bf16>(n); c.wt = a.take<bf16>(n); return c; }

  SanLayoutBf16(const iisan_san_desc& D, void* ws) {
    Arena a(ws);
    const size_t N = (size_t)D.n_items;
    const int E = D.emb;
    const bool dimdiff = D.d_text != D.d_img;
    const int dwide = D.d_text > D.d_img ? D.d_text : D.d_img;
    const int ft = D.asym ? E : D.d_text, fi = D.asym ? E : D.d_img, fm = D.d_mm;
    for (int s = 0; s < IISAN_MAX_STAGES; ++s) {
      t_down[s] = t_up[s] = i_down[s] = i_up[s] = m_down[s] = m_up[s] = dpw[s] = WCopy{nullptr, nullptr};
      x_t[s] = z_t[s] = last_t[s] = x_i[s] = z_i[s] = last_i[s] = x_m[s] = z_m[s] = last_m[s] = nullptr;
      dpo[s] = nullptr;
      a_t[s] = a_i[s] = a_m[s] = nullptr;
    }
    for (int s = 0; s < D.n_stages; ++s) {
      const int ta = D.text_adapter[s], ia = D.img_adapter[s], mi = D.mm_index[s];
      if (D.activation == IISAN_ACT_GELU) {
        if (ta >= 0) a_t[s] = a.take<float>(N * D.r_text);
        if (ia >= 0) a_i[s] = a.take<float>(N * D.r_img);
        if (mi >= 0) a_m[s] = a.take<float>(N * D.r_mm);
      }
      if (ta >= 0) {
        t_down[ta] = takew(a, (size_t)D.r_text * D.d_text); t_up[ta] = takew(a, (size_t)D.r_text * D.d_text);
      }
      if (ia >= 0) {
        i_down[ia] = takew(a, (size_t)D.r_img * D.d_img); i_up[ia] = takew(a, (size_t)D.r_img * D.d_img);
      }
      if (mi >= 0) {
        m_down[mi] = takew(a, (size_t)D.r_mm * D.d_mm); m_up[mi] = takew(a, (size_t)D.r_mm * D.d_mm);
        if (dimdiff) { dpw[mi] = takew(a, (size_t)D.d_mm * dwide); dpo[s] = a.take<float>(N * D.d_mm); }
      }
    }
    // x_s / z_s / last_s of one tower are contiguous over the stages, each stage padded to the 128-row tile of the fused chain
    // kernels (which address them as one [A * NP, .] matrix; no 256-byte arena padding in between)
    {
      const size_t NP = (size_t)chain_n_pad(D.n_items);
      bf16* xt = a.take<bf16>((size_t)D.n_stages * NP * D.d_text); bf16* xi = a.take<bf16>((size_t)D.n_stages * NP * D.d_img);
      bf16* xm = a.take<bf16>((size_t)D.n_stages * NP * D.d_mm);
      bf16* zt = a.take<bf16>((size_t)D.n_stages * NP * D.r_text); bf16* zi = a.take<bf16>((size_t)D.n_stages * NP * D.r_img);
      bf16* zm = a.take<bf16>((size_t)D.n_stages * NP * D.r_mm);
      bf16* bt = a.take<bf16>((size_t)D.n_stages * NP * D.d_text); bf16* bi = a.take<bf16>((size_t)D.n_stages * NP * D.d_img);
      bf16* bm = a.take<bf16>((size_t)D.n_stages * NP * D.d_mm);
      for (int s = 0; s < D.n_stages; ++s) {
        if (D.text_adapter[s] >= 0) { x_t[s] = xt + (size_t)s * NP * D.d_text; z_t[s] = zt + (size_t)s * NP * D.r_text; last_t[s] = bt + (size_t)s * NP * D.d_text; }
        if (D.img_adapter[s] >= 0) { x_i[s] = xi + (size_t)s * NP * D.d_img; z_i[s] = zi + (size_t)s * NP * D.r_img; last_i[s] = bi + (size_t)s * NP * D.d_img; }
        if (D.mm_index[s] >= 0) { x_m[s] = xm + (size_t)s * NP * D.d_mm; z_m[s] = zm + (size_t)s * NP * D.r_mm; last_m[s] = bm + (size_t)s * NP * D.d_mm; }
      }
    }
    fc_t = takew(a, (size_t)ft * D.d_text); fc_i = takew(a, (size_t)fi * D.d_img); fc_m = takew(a, (size_t)fm * D.d_mm);
    pre_t = takew(a, (size_t)E * ft); pre_i = takew(a, (size_t)E * fi); pre_m = takew(a, (size_t)E * fm);
    head_t = a.take<bf16>(N * ft); head_i = a.take<bf16>(N * fi); head_m = a.take<bf16>(N * fm);
    wide_b = (dimdiff && D.state_dtype != IISAN_BF16) ? a.take<bf16>(N * dwide) : nullptr;
    doutb = a.take<bf16>(N * 3 * E);
    dhead_t = a.take<bf16>(N * ft); dhead_i = a.take<bf16>(N * fi); dhead_m = a.take<bf16>(N * fm);
    dy_t = a.take<float>(N * D.d_text); dx_t = a.take<float>(N * D.d_text);
    dy_i = a.take<float>(N * D.d_img); dx_i = a.take<float>(N * D.d_img);
    dy_m = a.take<float>(N * D.d_mm); dx_m = a.take<float>(N * D.d_mm);
    dyb_t = a.take<bf16>(N * D.d_text); dxb_t = a.take<bf16>(N * D.d_text); dzb_t = a.take<bf16>(N * D.r_text);
    dyb_i = a.take<bf16>(N * D.d_img); dxb_i = a.take<bf16>(N * D.d_img); dzb_i = a.take<bf16>(N * D.r_img);
    dyb_m = a.take<bf16>(N * D.d_mm); dxb_m = a.take<bf16>(N * D.d_mm); dzb_m = a.take<bf16>(N * D.r_mm);
    ddpb = dimdiff ? a.take<bf16>(N * D.d_mm) : nullptr;
    for (int t = 0; t < 3; ++t) wd_pack[t] = wu_pack[t] = dys[t] = nullptr;
    if (san_chain_eligible(D)) {
      for (int t = 0; t < 3; ++t) {
        wd_pack[t] = a.take<bf16>((size_t)D.n_stages * D.r_mm * D.d_mm);
        wu_pack[t] = a.take<bf16>((size_t)D.n_stages * D.r_mm * D.d_mm);
        const size_t NP = (size_t)chain_n_pad(D.n_items);
        dys[t] = a.take<bf16>((size_t)D.n_stages * NP * D.d_mm);
        bf16* dz = a.take<bf16>((size_t)D.n_stages * NP * D.r_mm);
        for (int s = 0; s < D.n_stages; ++s) dzs[t][s] = dz + (size_t)s * NP * D.r_mm;
      }
    }
    lr_ws = nullptr;
    if (san_chain_eligible(D) && san_lr_eligible(D)) {
      lr_ws = a.base + a.off;
      a.off += align_up(san_lr_workspace_bytes(D), 256);
    }
    bytes = a.off;
  }
};

// test / profiling switches: IISAN_B200_NO_CHAIN=1 forces the layered path, IISAN_B200_CHAIN_GEN=1 | 2 an older generation of the
// fused chain (A/B measurements; the default is the newest generation that covers the shape: 3 = resident-state forward +
// low-rank adjoint backward (san_lr.cu) for d <= 768, 2 for the other multiples of 128)
static const bool g_disable_chain = [] { const char* e = getenv("IISAN_B200_NO_CHAIN"); return e && e[0] == '1'; }();
static std::atomic<int> g_chain_gen{[] { const char* e = getenv("IISAN_B200_CHAIN_GEN"); return (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 3; }()};
int set_chain_generation(int gen) { return gen >= 1 ? g_chain_gen.exchange(gen) : g_chain_gen.load(); }      // gen < 1: query only
// L2 prefetch distance (chunks) of the second-generation chain kernels; IISAN_B200_CHAIN_PF overrides (measurement switch)
static const int g_chain_pf_fwd = [] { const char* e = getenv("IISAN_B200_CHAIN_PF"); return e ? atoi(e) : 0; }();
static const int g_chain_pf_bwd = [] { const char* e = getenv("IISAN_B200_CHAIN_PF_BWD"); return e ? atoi(e) : 0; }();

size_t san_bf16_workspace_bytes(const iisan_san_desc& D) {
  SanLayoutBf16 L(D, nullptr);
  return L.bytes;
}

// TMA needs 16-byte row pitches and the UMMA N extent a multiple of 8.
int san_bf16_supported(const iisan_san_desc& D) {
  return (D.d_text % 8 == 0 && D.d_img % 8 == 0 && D.r_text % 8 == 0 && D.r_img % 8 == 0 && D.r_mm % 8 == 0 && D.emb % 8 == 0 &&
          D.out_ld % 4 == 0)
             ? 1
             : 0;
}

// ------------------------------------------------------------------------------------------------
// problem builders
// ------------------------------------------------------------------------------------------------
// y[M,N] = x[M,K] W[N,K]^T
static UmmaProblem mk_linear(const bf16* x, int64_t ldx, const bf16* w, int M, int N, int K) {
  UmmaProblem p{};
  p.A = UmmaOperand{x, M, K, ldx};
  p.B = UmmaOperand{w, N, K, K};
  p.a_mn_major = p.b_mn_major = 0;
  p.M = M; p.N = N; p.K = K; p.splitk = 1;
  return p;
}

// dx[M,N] = dy[M,K] W[K,N]   (data gradient: the nn.Linear weight [out = K, in = N] is read in place as an MN-major B operand)
static UmmaProblem mk_dgrad(const bf16* dy, int64_t lddy, const bf16* w, int M, int N, int K) {
  UmmaProblem p{};
  p.A = UmmaOperand{dy, M, K, lddy};
  p.B = UmmaOperand{w, K, N, N};
  p.a_mn_major = 0; p.b_mn_major = 1;
  p.M = M; p.N = N; p.K = K; p.splitk = 1;
  return p;
}

static int pick_split(int M, int N, int K, int nprob) {
  const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn) * (nprob < 1 ? 1 : nprob);
  const int kb = (K + 63) / 64;
  int s = (2 * 148 + tiles - 1) / tiles;
  if (s > kb / 2) s = kb / 2;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

// out[m,n] += y[rows,m]^T x[rows,n]   (weight gradient: reduction over the item rows, split-K + red.add)
// The larger of (m, n) is mapped to the UMMA M dimension; if that is n the tile is written transposed.
static UmmaProblem mk_wgrad(const bf16* y, int64_t ldy, int m, const bf16* x, int64_t ldx, int n, int rows, float* out, int nprob) {
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
  p.splitk = pick_split(p.M, p.N, p.K, nprob);
  return p;
}

template <typename T>
static const bf16* wide_operand(const iisan_san_desc* D, const void* base, int layers, int d, int layer, SanLayoutBf16& L,
                                int64_t* pitch, cudaStream_t st, int* status) {
  if (std::is_same<T, bf16>::value) {
    *pitch = (int64_t)layers * d;
    return reinterpret_cast<const bf16*>(base) + (int64_t)layer * d;
  }
  MixBatch gb{}; gb.n = 1;
  MixProb& g = gb.p[0];
  g.P = state_src<T>(base, layers, d, layer);
  g.mode = 2; g.X = nullptr; g.Xb = L.wide_b; g.N = D->n_items; g.d = d;
  *status = launch_mix<T>(gb, st);
  *pitch = d;
  return L.wide_b;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T>
static int san_forward_bf16_t(const iisan_san_desc* D, const iisan_san_params* P, const void* image, const void* text, void* ws,
                              float* out, cudaStream_t st) {
  SanLayoutBf16 L(*D, ws);
  const int N = D->n_items, E = D->emb;
  const bool dimdiff = D->d_text != D->d_img;
  const bool text_wide = D->d_text > D->d_img;
  const int dwide = text_wide ? D->d_text : D->d_img;
  const int ft = D->asym ? E : D->d_text, fi = D->asym ? E : D->d_img, fm = D->d_mm;
  const bool chain = san_chain_eligible(*D) && !g_disable_chain;
  if (chain && L.lr_ws && g_chain_gen.load(std::memory_order_relaxed) >= 3 && san_lr_usable(*D, *P))
    return san_lr_forward(D, P, image, text, L.lr_ws, out, st);
  // ---- bf16 weight copies (the fp32 nn.Parameters stay the source of truth) ----
  {
    CastList c(st);
    for (int s = 0; s < (chain ? 0 : D->n_stages); ++s) {      // layered path only: the chain kernels read the packed copies below
      const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
      if (ta >= 0) { c.add(P->text[ta].w_down, L.t_down[ta].w, L.t_down[ta].wt, D->r_text, D->d_text); c.add(P->text[ta].w_up, L.t_up[ta].w, L.t_up[ta].wt, D->d_text, D->r_text); }
      if (ia >= 0) { c.add(P->img[ia].w_down, L.i_down[ia].w, L.i_down[ia].wt, D->r_img, D->d_img); c.add(P->img[ia].w_up, L.i_up[ia].w, L.i_up[ia].wt, D->d_img, D->r_img); }
      if (mi >= 0) {
        c.add(P->mm[mi].w_down, L.m_down[mi].w, L.m_down[mi].wt, D->r_mm, D->d_mm); c.add(P->mm[mi].w_up, L.m_up[mi].w, L.m_up[mi].wt, D->d_mm, D->r_mm);
        if (dimdiff) c.add(P->down_project[mi].w, L.dpw[mi].w, L.dpw[mi].wt, D->d_mm, dwide);
      }
    }
    if (chain) {
      for (int s = 0; s < D->n_stages; ++s) {
        const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
        const size_t off = (size_t)s * D->r_mm * D->d_mm;
        c.add(P->text[ta].w_down, L.wd_pack[0] + off, nullptr, D->r_text, D->d_text); c.add(P->text[ta].w_up, L.wu_pack[0] + off, nullptr, D->d_text, D->r_text);
        c.add(P->img[ia].w_down, L.wd_pack[1] + off, nullptr, D->r_img, D->d_img); c.add(P->img[ia].w_up, L.wu_pack[1] + off, nullptr, D->d_img, D->r_img);
        c.add(P->mm[mi].w_down, L.wd_pack[2] + off, nullptr, D->r_mm, D->d_mm); c.add(P->mm[mi].w_up, L.wu_pack[2] + off, nullptr, D->d_mm, D->r_mm);
      }
    }
    // heads: plain casts (their data gradients read the [out, in] copies in place as MN-major operands)
    c.add(P->fc_text.w, L.fc_t.w, nullptr, ft, D->d_text); c.add(P->fc_img.w, L.fc_i.w, nullptr, fi, D->d_img);
    c.add(P->fc_mm.w, L.fc_m.w, nullptr, fm, D->d_mm);
    c.add(P->pre_text.w, L.pre_t.w, nullptr, E, ft); c.add(P->pre_img.w, L.pre_i.w, nullptr, E, fi);
    c.add(P->mm_down.w, L.pre_m.w, nullptr, E, fm);
    c.flush();
    IISAN_TRY(c.status);
  }
  const bf16* last_t = nullptr; const bf16* last_i = nullptr; const bf16* last_m = nullptr;
  bool first_t = true, first_i = true;
  if (chain) {
    // ---- all stages of all three towers in one launch (san_chain.cu) ----
    ChainArgs ca{};
    ca.n_items = N; ca.n_pad = chain_n_pad(N); ca.d = D->d_mm; ca.n_stages = D->n_stages; ca.pf = g_chain_pf_fwd;
    bool gen2 = g_chain_gen.load(std::memory_order_relaxed) >= 2 && chain2_shape_supported(D->d_mm);
    for (int s = 0; gen2 && s < D->n_stages; ++s) {      // the bias vectors are staged by 16-byte bulk copies
      const uintptr_t al = reinterpret_cast<uintptr_t>(P->text[D->text_adapter[s]].b_up) | reinterpret_cast<uintptr_t>(P->text[D->text_adapter[s]].b_down) |
                           reinterpret_cast<uintptr_t>(P->img[D->img_adapter[s]].b_up) | reinterpret_cast<uintptr_t>(P->img[D->img_adapter[s]].b_down) |
                           reinterpret_cast<uintptr_t>(P->mm[D->mm_index[s]].b_up) | reinterpret_cast<uintptr_t>(P->mm[D->mm_index[s]].b_down);
      if (al & 15) gen2 = false;
    }
    ca.rows = gen2 ? chain_tile_rows(N) : 128;
    IISAN_TRY(chain_fill_tower(&ca.tower[0], 0, text, N, (int64_t)D->layers_text * D->d_text, nullptr, 0, L.wd_pack[0], L.wu_pack[0], L.x_t[0], L.last_t[0], L.z_t[0], D->n_stages, D->d_text, ca.rows));
    IISAN_TRY(chain_fill_tower(&ca.tower[1], 0, image, N, (int64_t)D->layers_img * D->d_img, nullptr, 0, L.wd_pack[1], L.wu_pack[1], L.x_i[0], L.last_i[0], L.z_i[0], D->n_stages, D->d_img, ca.rows));
    IISAN_TRY(chain_fill_tower(&ca.tower[2], 1, image, N, (int64_t)D->layers_img * D->d_img, text, (int64_t)D->layers_text * D->d_text, L.wd_pack[2], L.wu_pack[2], L.x_m[0], L.last_m[0], L.z_m[0], D->n_stages, D->d_mm, ca.rows));
    // only last_{A-1} (the operand of the heads) is written: the backward recovers h_s - last_{s-1} of the intra-modal towers
    // from the x_s stash, and the inter-modal gate gradient only needs the raw states
    ca.tower[0].store_last = 0; ca.tower[1].store_last = 0; ca.tower[2].store_last = 0;
    for (int s = 0; s < D->n_stages; ++s) {
      const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
      ChainTower& t0 = ca.tower[0]; ChainTower& t1 = ca.tower[1]; ChainTower& t2 = ca.tower[2];
      t0.layer[s] = D->text_layer[s]; t0.gate[s] = P->gate_text[ta]; t0.b_down[s] = P->text[ta].b_down; t0.b_up[s] = P->text[ta].b_up;
      t1.layer[s] = D->img_layer[s]; t1.gate[s] = P->gate_img[ia]; t1.b_down[s] = P->img[ia].b_down; t1.b_up[s] = P->img[ia].b_up;
      t2.layer[s] = D->img_layer[s]; t2.layer2[s] = D->text_layer[s]; t2.gate[s] = P->gate_mm[mi]; t2.b_down[s] = P->mm[mi].b_down; t2.b_up[s] = P->mm[mi].b_up;
    }
    if (gen2) IISAN_TRY(launch_san_chain2_fwd(ca, 3, st));
    else IISAN_TRY(launch_san_chain_fwd(ca, 3, st));
    last_t = L.last_t[D->n_stages - 1]; last_i = L.last_i[D->n_stages - 1]; last_m = L.last_m[D->n_stages - 1];
  }
  for (int s = 0; s < (chain ? 0 : D->n_stages); ++s) {
    const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
    // ---- dim alignment GEMM on the raw wide state (CA/model/model.py:406-411) ----
    const float* dp = nullptr;
    if (mi >= 0 && dimdiff) {
      int status = IISAN_OK; int64_t pitch = 0;
      const bf16* a = text_wide ? wide_operand<T>(D, text, D->layers_text, D->d_text, D->text_layer[s], L, &pitch, st, &status)
                                : wide_operand<T>(D, image, D->layers_img, D->d_img, D->img_layer[s], L, &pitch, st, &status);
      IISAN_TRY(status);
      UmmaBatch b{}; b.n = 1;
      b.p[0] = mk_linear(a, pitch, L.dpw[mi].w, N, D->d_mm, dwide);
      b.p[0].epi.bias = P->down_project[mi].b; b.p[0].epi.out_f32 = L.dpo[s]; b.p[0].epi.ld_f32 = D->d_mm;
      IISAN_TRY(launch_umma_gemm(b, st));
      dp = L.dpo[s];
    }
    MixBatch mb{}; UmmaBatch down{}, up{};
    auto tower = [&](const iisan_adapter_ptrs& p, const WCopy& wd, const WCopy& wu, bf16* x, bf16* z, float* pre, bf16* last, int d, int r) {
      UmmaProblem& dn = down.p[down.n++];
      dn = mk_linear(x, d, wd.w, N, r, d);
      dn.epi.bias = p.b_down; dn.epi.out_bf16 = z; dn.epi.ld_bf16 = r;
      if (pre) { dn.epi.gelu = 1; dn.epi.out_f32 = pre; dn.epi.ld_f32 = r; }      // GELU: z = gelu(pre), pre kept for the backward
      else dn.epi.relu = 1;
      UmmaProblem& u = up.p[up.n++];
      u = mk_linear(z, r, wu.w, N, d, r);
      u.epi.bias = p.b_up; u.epi.resid_bf16 = x; u.epi.ld_resid_bf16 = d; u.epi.out_bf16 = last; u.epi.ld_bf16 = d;
    };
    if (ta >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      if (first_t && D->remove_first) m.R = state_src<T>(text, D->layers_text, D->d_text, 0);
      else if (last_t) m.R = dense_bf16_src(last_t, D->d_text);
      else m.R = MixSrc{nullptr, 0, 0};
      m.gate = P->gate_text[ta]; m.mode = 0; m.X = nullptr; m.Xb = L.x_t[s]; m.N = N; m.d = D->d_text;
      tower(P->text[ta], L.t_down[ta], L.t_up[ta], L.x_t[s], L.z_t[s], L.a_t[s], L.last_t[s], D->d_text, D->r_text);
    }
    if (ia >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      if (first_i && D->remove_first) m.R = state_src<T>(image, D->layers_img, D->d_img, 0);
      else if (last_i) m.R = dense_bf16_src(last_i, D->d_img);
      else m.R = MixSrc{nullptr, 0, 0};
      m.gate = P->gate_img[ia]; m.mode = 0; m.X = nullptr; m.Xb = L.x_i[s]; m.N = N; m.d = D->d_img;
      tower(P->img[ia], L.i_down[ia], L.i_up[ia], L.x_i[s], L.z_i[s], L.a_i[s], L.last_i[s], D->d_img, D->r_img);
    }
    if (mi >= 0) {
      MixProb& m = mb.p[mb.n++];
      m.P = (dp && !text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      m.Q = (dp && text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      m.R = last_m ? dense_bf16_src(last_m, D->d_mm) : MixSrc{nullptr, 0, 0};
      m.gate = P->gate_mm[mi]; m.mode = 1; m.X = nullptr; m.Xb = L.x_m[s]; m.N = N; m.d = D->d_mm;
      tower(P->mm[mi], L.m_down[mi], L.m_up[mi], L.x_m[s], L.z_m[s], L.a_m[s], L.last_m[s], D->d_mm, D->r_mm);
    }
    IISAN_TRY(launch_mix<T>(mb, st));
    IISAN_TRY(launch_umma_gemm(down, st));
    IISAN_TRY(launch_umma_gemm(up, st));
    if (ta >= 0) { last_t = L.last_t[s]; first_t = false; }
    if (ia >= 0) { last_i = L.last_i[s]; first_i = false; }
    if (mi >= 0) last_m = L.last_m[s];
  }
  if (!last_t || !last_i || !last_m) return IISAN_EINVAL;
  // ---- heads: e = pre(fc(last)) ----
  UmmaBatch fc{}; fc.n = 3;
  fc.p[0] = mk_linear(last_i, D->d_img, L.fc_i.w, N, fi, D->d_img);
  fc.p[0].epi.bias = P->fc_img.b; fc.p[0].epi.out_bf16 = L.head_i; fc.p[0].epi.ld_bf16 = fi;
  fc.p[1] = mk_linear(last_t, D->d_text, L.fc_t.w, N, ft, D->d_text);
  fc.p[1].epi.bias = P->fc_text.b; fc.p[1].epi.out_bf16 = L.head_t; fc.p[1].epi.ld_bf16 = ft;
  fc.p[2] = mk_linear(last_m, D->d_mm, L.fc_m.w, N, fm, D->d_mm);
  fc.p[2].epi.bias = P->fc_mm.b; fc.p[2].epi.out_bf16 = L.head_m; fc.p[2].epi.ld_bf16 = fm;
  IISAN_TRY(launch_umma_gemm(fc, st));
  UmmaBatch pre{}; pre.n = 3;
  pre.p[0] = mk_linear(L.head_i, fi, L.pre_i.w, N, E, fi);
  pre.p[0].epi.bias = P->pre_img.b; pre.p[0].epi.out_f32 = out; pre.p[0].epi.ld_f32 = D->out_ld;
  pre.p[1] = mk_linear(L.head_t, ft, L.pre_t.w, N, E, ft);
  pre.p[1].epi.bias = P->pre_text.b; pre.p[1].epi.out_f32 = out + E; pre.p[1].epi.ld_f32 = D->out_ld;
  pre.p[2] = mk_linear(L.head_m, fm, L.pre_m.w, N, E, fm);
  pre.p[2].epi.bias = P->mm_down.b; pre.p[2].epi.out_f32 = out + 2 * E; pre.p[2].epi.ld_f32 = D->out_ld;
  IISAN_TRY(launch_umma_gemm(pre, st));
  return IISAN_OK;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename T>
static int san_backward_bf16_t(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G, const void* image,
                               const void* text, void* ws, const float* d_out, cudaStream_t st) {
  SanLayoutBf16 L(*D, ws);
  const int N = D->n_items, E = D->emb;
  const bool dimdiff = D->d_text != D->d_img;
  const bool text_wide = D->d_text > D->d_img;
  const int dwide = text_wide ? D->d_text : D->d_img;
  const int ft = D->asym ? E : D->d_text, fi = D->asym ? E : D->d_img, fm = D->d_mm;
  int ls_t = -1, ls_i = -1, ls_m = -1;
  for (int s = 0; s < D->n_stages; ++s) {
    if (D->text_adapter[s] >= 0) ls_t = s;
    if (D->img_adapter[s] >= 0) ls_i = s;
    if (D->mm_index[s] >= 0) ls_m = s;
  }
  if (ls_t < 0 || ls_i < 0 || ls_m < 0) return IISAN_EINVAL;
  const bool chain = san_chain_eligible(*D) && !g_disable_chain;
  if (chain && L.lr_ws && g_chain_gen.load(std::memory_order_relaxed) >= 3 && san_lr_usable(*D, *P))
    return san_lr_backward(D, P, G, image, text, L.lr_ws, d_out, st);
  // ---- bf16 operand copy of d_out ----
  {
    MixBatch cb{}; cb.n = 1;
    MixProb& c = cb.p[0];
    c.P = dense_src(d_out, D->out_ld); c.mode = 2; c.X = nullptr; c.Xb = L.doutb; c.N = N; c.d = 3 * E;
    IISAN_TRY(launch_mix<T>(cb, st));
  }
  const bf16* do_i = L.doutb; const bf16* do_t = L.doutb + E; const bf16* do_m = L.doutb + 2 * E;
  const int ldo = 3 * E;
  // ---- heads backward ----
  {
    UmmaBatch w{}; w.n = 3;   // d pre weights [E, f] += dout^T head
    w.p[0] = mk_wgrad(do_i, ldo, E, L.head_i, fi, fi, N, G->pre_img.w, 3);
    w.p[1] = mk_wgrad(do_t, ldo, E, L.head_t, ft, ft, N, G->pre_text.w, 3);
    w.p[2] = mk_wgrad(do_m, ldo, E, L.head_m, fm, fm, N, G->mm_down.w, 3);
    IISAN_TRY(launch_umma_gemm(w, st));
    UmmaBatch dh{}; dh.n = 3;  // d head = d_out W_pre
    dh.p[0] = mk_dgrad(do_i, ldo, L.pre_i.w, N, fi, E); dh.p[0].epi.out_bf16 = L.dhead_i; dh.p[0].epi.ld_bf16 = fi;
    dh.p[1] = mk_dgrad(do_t, ldo, L.pre_t.w, N, ft, E); dh.p[1].epi.out_bf16 = L.dhead_t; dh.p[1].epi.ld_bf16 = ft;
    dh.p[2] = mk_dgrad(do_m, ldo, L.pre_m.w, N, fm, E); dh.p[2].epi.out_bf16 = L.dhead_m; dh.p[2].epi.ld_bf16 = fm;
    IISAN_TRY(launch_umma_gemm(dh, st));
    UmmaBatch wf{}; wf.n = 3;  // d fc weights [f, d] += dhead^T last
    wf.p[0] = mk_wgrad(L.dhead_i, fi, fi, L.last_i[ls_i], D->d_img, D->d_img, N, G->fc_img.w, 3);
    wf.p[1] = mk_wgrad(L.dhead_t, ft, ft, L.last_t[ls_t], D->d_text, D->d_text, N, G->fc_text.w, 3);
    wf.p[2] = mk_wgrad(L.dhead_m, fm, fm, L.last_m[ls_m], D->d_mm, D->d_mm, N, G->fc_mm.w, 3);
    IISAN_TRY(launch_umma_gemm(wf, st));
    ColsumBatch cf{}; cf.n = 6;   // bias gradients of both head layers in one launch
    cf.p[0] = {nullptr, fi, N, fi, G->fc_img.b, L.dhead_i};
    cf.p[1] = {nullptr, ft, N, ft, G->fc_text.b, L.dhead_t};
    cf.p[2] = {nullptr, fm, N, fm, G->fc_mm.b, L.dhead_m};
    cf.p[3] = {d_out, D->out_ld, N, E, G->pre_img.b, nullptr};
    cf.p[4] = {d_out + E, D->out_ld, N, E, G->pre_text.b, nullptr};
    cf.p[5] = {d_out + 2 * E, D->out_ld, N, E, G->mm_down.b, nullptr};
    IISAN_TRY(launch_colsum(cf, st));
    UmmaBatch dl{}; dl.n = 3;  // d last = d head W_fc
    const size_t lastoff = (size_t)(D->n_stages - 1) * chain_n_pad(N) * D->d_mm;
    dl.p[0] = mk_dgrad(L.dhead_i, fi, L.fc_i.w, N, D->d_img, fi);
    dl.p[1] = mk_dgrad(L.dhead_t, ft, L.fc_t.w, N, D->d_text, ft);
    dl.p[2] = mk_dgrad(L.dhead_m, fm, L.fc_m.w, N, D->d_mm, fm);
    if (chain) {      // the fused backward takes d last_{A-1} as bf16 from the per-stage gradient stash
      dl.p[0].epi.out_bf16 = L.dys[1] + lastoff; dl.p[0].epi.ld_bf16 = D->d_img;
      dl.p[1].epi.out_bf16 = L.dys[0] + lastoff; dl.p[1].epi.ld_bf16 = D->d_text;
      dl.p[2].epi.out_bf16 = L.dys[2] + lastoff; dl.p[2].epi.ld_bf16 = D->d_mm;
    } else {
      dl.p[0].epi.out_f32 = L.dy_i; dl.p[0].epi.ld_f32 = D->d_img; dl.p[0].epi.out_bf16 = L.dyb_i; dl.p[0].epi.ld_bf16 = D->d_img;
      dl.p[1].epi.out_f32 = L.dy_t; dl.p[1].epi.ld_f32 = D->d_text; dl.p[1].epi.out_bf16 = L.dyb_t; dl.p[1].epi.ld_bf16 = D->d_text;
      dl.p[2].epi.out_f32 = L.dy_m; dl.p[2].epi.ld_f32 = D->d_mm; dl.p[2].epi.out_bf16 = L.dyb_m; dl.p[2].epi.ld_bf16 = D->d_mm;
    }
    IISAN_TRY(launch_umma_gemm(dl, st));
  }
  if (chain) {
    // ---- data / gate / bias gradients of all stages and towers in one launch (san_chain.cu) ----
    ChainBwdArgs ca{};
    ca.n_items = N; ca.n_pad = chain_n_pad(N); ca.d = D->d_mm; ca.n_stages = D->n_stages; ca.pf = g_chain_pf_bwd;
    const bool gen2 = g_chain_gen.load(std::memory_order_relaxed) >= 2 && chain2_shape_supported(D->d_mm);
    ca.rows = gen2 ? chain_tile_rows(N) : 128;
    IISAN_TRY(chain_fill_bwd_tower(&ca.tower[0], 0, text, N, (int64_t)D->layers_text * D->d_text, nullptr, 0, L.wd_pack[0], L.wu_pack[0], L.dys[0], L.x_t[0], L.dzs[0][0], D->n_stages, D->d_text, ca.rows));
    IISAN_TRY(chain_fill_bwd_tower(&ca.tower[1], 0, image, N, (int64_t)D->layers_img * D->d_img, nullptr, 0, L.wd_pack[1], L.wu_pack[1], L.dys[1], L.x_i[0], L.dzs[1][0], D->n_stages, D->d_img, ca.rows));
    IISAN_TRY(chain_fill_bwd_tower(&ca.tower[2], 1, image, N, (int64_t)D->layers_img * D->d_img, text, (int64_t)D->layers_text * D->d_text, L.wd_pack[2], L.wu_pack[2], L.dys[2], L.x_m[0], L.dzs[2][0], D->n_stages, D->d_mm, ca.rows));
    ca.tower[0].z_stash = L.z_t[0]; ca.tower[1].z_stash = L.z_i[0]; ca.tower[2].z_stash = L.z_m[0];
    for (int s = 0; s < D->n_stages; ++s) {
      const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
      ChainBwdTower& t0 = ca.tower[0]; ChainBwdTower& t1 = ca.tower[1]; ChainBwdTower& t2 = ca.tower[2];
      t0.layer[s] = D->text_layer[s]; t0.gate[s] = P->gate_text[ta]; t0.g_gate[s] = G->gate_text[ta]; t0.g_b_down[s] = G->text[ta].b_down; t0.g_b_up[s] = G->text[ta].b_up;
      t1.layer[s] = D->img_layer[s]; t1.gate[s] = P->gate_img[ia]; t1.g_gate[s] = G->gate_img[ia]; t1.g_b_down[s] = G->img[ia].b_down; t1.g_b_up[s] = G->img[ia].b_up;
      t2.layer[s] = D->img_layer[s]; t2.layer2[s] = D->text_layer[s]; t2.gate[s] = P->gate_mm[mi]; t2.g_gate[s] = G->gate_mm[mi]; t2.g_b_down[s] = G->mm[mi].b_down; t2.g_b_up[s] = G->mm[mi].b_up;
    }
    if (gen2) IISAN_TRY(launch_san_chain2_bwd(ca, 3, st));
    else IISAN_TRY(launch_san_chain_bwd(ca, 3, st));
    // ---- weight gradients: reductions over all items; the 6 x A split-K GEMMs over the stashes share ONE launch ----
    {
      static thread_local UmmaBatchBig wg;
      wg.n = 0;
      const int np = 6 * D->n_stages;
      for (int s = 0; s < D->n_stages; ++s) {       // the chain backward ends with stage 0: its stashes are the hottest in L2
        const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
        const size_t off = (size_t)s * chain_n_pad(N) * D->d_mm;
        wg.p[wg.n++] = mk_wgrad(L.dys[0] + off, D->d_text, D->d_text, L.z_t[s], D->r_text, D->r_text, N, G->text[ta].w_up, np);
        wg.p[wg.n++] = mk_wgrad(L.dys[1] + off, D->d_img, D->d_img, L.z_i[s], D->r_img, D->r_img, N, G->img[ia].w_up, np);
        wg.p[wg.n++] = mk_wgrad(L.dys[2] + off, D->d_mm, D->d_mm, L.z_m[s], D->r_mm, D->r_mm, N, G->mm[mi].w_up, np);
        wg.p[wg.n++] = mk_wgrad(L.dzs[0][s], D->r_text, D->r_text, L.x_t[s], D->d_text, D->d_text, N, G->text[ta].w_down, np);
        wg.p[wg.n++] = mk_wgrad(L.dzs[1][s], D->r_img, D->r_img, L.x_i[s], D->d_img, D->d_img, N, G->img[ia].w_down, np);
        wg.p[wg.n++] = mk_wgrad(L.dzs[2][s], D->r_mm, D->r_mm, L.x_m[s], D->d_mm, D->d_mm, N, G->mm[mi].w_down, np);
      }
      IISAN_TRY(launch_umma_gemm_big(wg, st));
    }
    return IISAN_OK;
  }
  // ---- stages in reverse: dy_* holds d last_s on entry to stage s and d last_{s-1} on exit ----
  float* dy_t = L.dy_t; float* dx_t = L.dx_t; bf16* dyb_t = L.dyb_t; bf16* dxb_t = L.dxb_t;
  float* dy_i = L.dy_i; float* dx_i = L.dx_i; bf16* dyb_i = L.dyb_i; bf16* dxb_i = L.dxb_i;
  float* dy_m = L.dy_m; float* dx_m = L.dx_m; bf16* dyb_m = L.dyb_m; bf16* dxb_m = L.dxb_m;
  for (int s = D->n_stages - 1; s >= 0; --s) {
    const int ta = D->text_adapter[s], ia = D->img_adapter[s], mi = D->mm_index[s];
    int ps_t = -1, ps_i = -1;
    for (int q = 0; q < s; ++q) {
      if (D->text_adapter[q] >= 0) ps_t = q;
      if (D->img_adapter[q] >= 0) ps_i = q;
    }
    const int nt = (ta >= 0) + (ia >= 0) + (mi >= 0);
    UmmaBatch wu{}, dz{}, wd{}, dx{}; ColsumBatch cu{}, cd{};
    auto add = [&](const iisan_adapter_ptrs& g, const WCopy& wdn, const WCopy& wup, const bf16* x, const bf16* z, const float* pre, bf16* dzb,
                   const float* dy, const bf16* dyb, float* dxo, bf16* dxob, int d, int r) {
      wu.p[wu.n++] = mk_wgrad(dyb, d, d, z, r, r, N, g.w_up, nt);               // dWu[d,r] += dy^T z
      cu.p[cu.n++] = {dy, d, N, d, g.b_up};
      UmmaProblem& q = dz.p[dz.n++];                                             // dz = (dy Wu) * (z > 0)
      q = mk_linear(dyb, d, wup.wt, N, r, d);
      if (pre) { q.epi.gelu_pre = pre; q.epi.ld_gelu_pre = r; }                  // GELU: dz = (dy Wu) * gelu'(pre)
      else { q.epi.mask = z; q.epi.ld_mask = r; }
      q.epi.out_bf16 = dzb; q.epi.ld_bf16 = r;
      wd.p[wd.n++] = mk_wgrad(dzb, r, r, x, d, d, N, g.w_down, nt);             // dWd[r,d] += dz^T x
      cd.p[cd.n++] = {nullptr, r, N, r, g.b_down, dzb};
      UmmaProblem& u = dx.p[dx.n++];                                             // dx = dy + dz Wd
      u = mk_linear(dzb, r, wdn.wt, N, d, r);
      u.epi.resid_f32 = dy; u.epi.ld_resid_f32 = d; u.epi.out_f32 = dxo; u.epi.ld_f32 = d; u.epi.out_bf16 = dxob; u.epi.ld_bf16 = d;
    };
    if (ta >= 0) add(G->text[ta], L.t_down[ta], L.t_up[ta], L.x_t[s], L.z_t[s], L.a_t[s], L.dzb_t, dy_t, dyb_t, dx_t, dxb_t, D->d_text, D->r_text);
    if (ia >= 0) add(G->img[ia], L.i_down[ia], L.i_up[ia], L.x_i[s], L.z_i[s], L.a_i[s], L.dzb_i, dy_i, dyb_i, dx_i, dxb_i, D->d_img, D->r_img);
    if (mi >= 0) add(G->mm[mi], L.m_down[mi], L.m_up[mi], L.x_m[s], L.z_m[s], L.a_m[s], L.dzb_m, dy_m, dyb_m, dx_m, dxb_m, D->d_mm, D->r_mm);
    IISAN_TRY(launch_umma_gemm(wu, st));
    IISAN_TRY(launch_colsum(cu, st));
    IISAN_TRY(launch_umma_gemm(dz, st));
    IISAN_TRY(launch_umma_gemm(wd, st));
    IISAN_TRY(launch_colsum(cd, st));
    IISAN_TRY(launch_umma_gemm(dx, st));
    // ---- fusion backward ----
    MixBwdBatch mb{};
    const bool has_dp = (mi >= 0 && dimdiff);
    if (ta >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      m.P = state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      if (ps_t < 0) { if (D->remove_first) m.R = state_src<T>(text, D->layers_text, D->d_text, 0); else m.R = MixSrc{nullptr, 0, 0}; }
      else m.R = dense_bf16_src(L.last_t[ps_t], D->d_text);
      m.gate = P->gate_text[ta]; m.dgate = G->gate_text[ta]; m.mode = 0; m.dX = dx_t;
      m.dPrev = (ps_t >= 0) ? dy_t : nullptr; m.dPrevb = (ps_t >= 0) ? dyb_t : nullptr; m.N = N; m.d = D->d_text;
    }
    if (ia >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      m.P = state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      if (ps_i < 0) { if (D->remove_first) m.R = state_src<T>(image, D->layers_img, D->d_img, 0); else m.R = MixSrc{nullptr, 0, 0}; }
      else m.R = dense_bf16_src(L.last_i[ps_i], D->d_img);
      m.gate = P->gate_img[ia]; m.dgate = G->gate_img[ia]; m.mode = 0; m.dX = dx_i;
      m.dPrev = (ps_i >= 0) ? dy_i : nullptr; m.dPrevb = (ps_i >= 0) ? dyb_i : nullptr; m.N = N; m.d = D->d_img;
    }
    if (mi >= 0) {
      MixBwdProb& m = mb.p[mb.n++];
      const float* dp = has_dp ? L.dpo[s] : nullptr;
      m.P = (dp && !text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(image, D->layers_img, D->d_img, D->img_layer[s]);
      m.Q = (dp && text_wide) ? dense_src(dp, D->d_mm) : state_src<T>(text, D->layers_text, D->d_text, D->text_layer[s]);
      m.gate = P->gate_mm[mi]; m.dgate = G->gate_mm[mi]; m.mode = 1; m.dX = dx_m;
      m.dP_outb = (has_dp && !text_wide) ? L.ddpb : nullptr;
      m.dQ_outb = (has_dp && text_wide) ? L.ddpb : nullptr;
      m.N = N; m.d = D->d_mm;
    }
    IISAN_TRY(launch_mix_bwd<T>(mb, st));
    if (mi >= 0) {   // d last_mm_{s-1} = dx_mm
      float* t = dy_m; dy_m = dx_m; dx_m = t;
      bf16* tb = dyb_m; dyb_m = dxb_m; dxb_m = tb;
    }
    if (has_dp) {
      // down_project gradients: dW[d_mm, dwide] += ddp^T h_wide ; db += colsum(ddp)
      int status = IISAN_OK; int64_t pitch = 0;
      const bf16* hw = text_wide ? wide_operand<T>(D, text, D->layers_text, D->d_text, D->text_layer[s], L, &pitch, st, &status)
                                 : wide_operand<T>(D, image, D->layers_img, D->d_img, D->img_layer[s], L, &pitch, st, &status);
      IISAN_TRY(status);
      UmmaBatch w{}; w.n = 1;
      w.p[0] = mk_wgrad(L.ddpb, D->d_mm, D->d_mm, hw, pitch, dwide, N, G->down_project[mi].w, 1);
      IISAN_TRY(launch_umma_gemm(w, st));
      ColsumBatch c{}; c.n = 1;
      c.p[0] = {nullptr, D->d_mm, N, D->d_mm, G->down_project[mi].b, L.ddpb};
      IISAN_TRY(launch_colsum(c, st));
    }
  }
  return IISAN_OK;
}

int san_forward_bf16(const iisan_san_desc* D, const iisan_san_params* P, const void* image, const void* text, void* ws, float* out,
                     cudaStream_t st) {
  switch (D->state_dtype) {
    case IISAN_F32: return san_forward_bf16_t<float>(D, P, image, text, ws, out, st);
    case IISAN_BF16: return san_forward_bf16_t<bf16>(D, P, image, text, ws, out, st);
    case IISAN_F16: return san_forward_bf16_t<__half>(D, P, image, text, ws, out, st);
  }
  return IISAN_EINVAL;
}

int san_backward_bf16(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G, const void* image,
                      const void* text, void* ws, const float* d_out, cudaStream_t st) {
  switch (D->state_dtype) {
    case IISAN_F32: return san_backward_bf16_t<float>(D, P, G, image, text, ws, d_out, st);
    case IISAN_BF16: return san_backward_bf16_t<bf16>(D, P, G, image, text, ws, d_out, st);
    case IISAN_F16: return san_backward_bf16_t<__half>(D, P, G, image, text, ws, d_out, st);
  }
  return IISAN_EINVAL;
}

}  // namespace iisan
