// In-batch softmax cross-entropy, fast mode (IISAN_COMPUTE_BF16): logits tiles on tcgen05/TMEM, never materialised.
//
// Algorithm restated from CC/model/model.py:63-64 and :81-105 (nothing ported; see inbatch_ce.cu for the exact-mode twin):
//   logit(r, c) = <prec[r], score[c]> - log(pop[ids_cols[c]])  ;  masked entries are overwritten with -1e4:
//   column (u, p) with p < L and log_mask_cols[u, p] == 0 (:88-89)  or  ids_cols[c] among the 11 ids of the row's user and
//   c != label(r) (:91-100) ;  label(i, j) = (user_offset + i) * (L+1) + j + 1 (:82-85) ;  mean CE over valid rows (:102-104).
//
// Structure
//   prepass   : bf16 copies of prec / score (the MMA operands), debias[c] = log(pop[id_c]) and ONE mask bit per
//               (row-user, column): col-pad OR id membership, by exact int64 compares (bit-exact with the reference's
//               masks; 10x fewer compares than per (row, column) because the 10 rows of a user share their reject set).
//   tile pass : one CTA owns 128 loss rows (TMEM lanes) and streams 128-wide tiles of item columns: forward (online
//               log-sum-exp) or backward (d_prec in the CTA's accumulator AND the per-tile d_score partial)
//               S = O x T^T (K = E = 64) by tcgen05.mma into a double-buffered TMEM accumulator; the 8 epilogue warps
//               turn S into masked logits / softmax weights; for the backward the bf16 weight tile goes back to shared
//               memory (128B-swizzled K-major A operand) and a second tcgen05.mma accumulates W x T (the streamed tile is
//               reused as an MN-major B operand) into a TMEM accumulator that is flushed once per CTA.
//   warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..17 epilogue (the epilogue is ALU / latency
//               bound -- K = 64 makes the MMAs negligible -- so every TMEM lane quadrant gets four warps, one per 32-column
//               quarter of the streamed tile).
#include "common.cuh"
#include "launch.cuh"
#include "umma.cuh"
#include "ce_umma.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace iisan {

using namespace umma;
using bf16 = __nv_bfloat16;

constexpr float kNegMaskF = -1e4f;
constexpr int CT = 128;            // tile extent on both sides
constexpr int CE_E = 64;           // embedding width handled by this kernel (one 128-byte swizzle atom of bf16)
constexpr int CE_NST = 3;          // streamed-tile ring
constexpr int CE_EPI_WARPS = 16;
constexpr int CE_THREADS = 64 + 32 * CE_EPI_WARPS;    // 18 warps
constexpr int CE_EPI_THREADS = 32 * CE_EPI_WARPS;
constexpr int CE_TILE_BYTES = CT * CE_E * 2;     // 16 KB
constexpr int CE_A2_BYTES = CT * CT * 2;         // 32 KB
constexpr int CE_TMEM_COLS = 512;
constexpr int CE_ACC_COL = 256;
constexpr int CE_DSC_COL = 320;    // MODE 3: two [128 x 64] accumulators of the per-tile d_score partial

struct CeTileArgs {
  CUtensorMap map_prec, map_score;
  int R, L, S, B;                 // rows = B*L
  int C, Cw;                      // columns = Bc*S ; mask words per user
  int64_t user_offset;
  const float* lm_rows;           // [B, L]
  const float* lm_cols;           // [Bc, L]
  const float* debias;            // [C]
  const uint32_t* maskbits;       // [B, Cw]
  const float* lse;               // [R] (backward)
  const float* g_sum; const float* g_mean; const int32_t* n_valid;
  float* part_m; float* part_s; float* part_lab;   // forward partials [splits, R]
  float* d_out;                   // backward: d_prec [R, E] or d_score [C, E], accumulated atomically
  float* d_out2;                  // MODE 3: d_score [C, E] (d_out = d_prec)
  int tiles_stream;               // number of streamed tiles in total
  int tiles_per_split;
};

struct CeSmem {
  static constexpr int kO = 0;
  static constexpr int kT = kO + CE_TILE_BYTES;
  static constexpr int kA2 = kT + CE_NST * CE_TILE_BYTES;
  static constexpr int kBar = kA2 + 2 * CE_A2_BYTES;
  static constexpr int kAttr = kBar + 256;
  static constexpr int kTotal = kAttr + 2 * CT * 40 + 1024;   // attributes (double buffered) / forward combine scratch (9 x 128 floats) + align slack
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CE_EPI_THREADS) : "memory"); }

__device__ __forceinline__ float ce_scale_dev(const float* g_sum, const float* g_mean, const int32_t* n_valid) {
  float s = 0.f;
  if (g_sum) s += __ldg(g_sum);
  if (g_mean) s += __ldg(g_mean) / (float)__ldg(n_valid);
  return s;
}

constexpr float kLog2e = 1.4426950408889634f;
// 2^x, flush-to-zero approximation (one MUFU; exp(x) = 2^(x log2 e) folds the scale and the max / lse subtraction into one FMA)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0: forward partial log-sum-exp
// MODE 3: d_prec AND d_score in one pass: the bf16 weight tile W [128 rows x 128 columns] that feeds
//         acc_prec += W x T is read a second time as an MN-major A operand, dsc = W^T x O (K = the CTA's 128 rows), and the
//         [128 columns x 64] partial is added to d_score by vector reductions once per streamed tile.
template <int MODE>
__global__ void __launch_bounds__(CE_THREADS, 1) ce_tile_kernel(const __grid_constant__ CeTileArgs a) {
  static_assert(MODE == 0 || MODE == 3, "forward or fused backward");
  constexpr bool BWD = (MODE != 0);
  constexpr bool FUSED = (MODE == 3);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CeSmem::kBar);
  uint64_t* o_full = bars;                 // 1
  uint64_t* t_full = bars + 1;             // CE_NST
  uint64_t* t_empty = t_full + CE_NST;     // CE_NST
  uint64_t* s_full = t_empty + CE_NST;     // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* a2_full = s_empty + 2;         // 2
  uint64_t* a2_empty = a2_full + 2;        // 2
  uint64_t* acc_full = a2_empty + 2;       // 1
  uint64_t* dsc_full = acc_full + 1;       // 2
  uint64_t* dsc_empty = dsc_full + 2;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dsc_empty + 2);
  float* attr_f = reinterpret_cast<float*>(smem + CeSmem::kAttr);   // [2][CT][4] floats / ints

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * CT;                                     // first owner entity
  const int t_beg = blockIdx.y * a.tiles_per_split;
  const int t_end = min(a.tiles_stream, t_beg + a.tiles_per_split);
  const int n_tiles = t_end - t_beg;
  if (n_tiles <= 0) return;
  const CUtensorMap* map_o = &a.map_prec;
  const CUtensorMap* map_t = &a.map_score;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(map_o); tma_prefetch_desc(map_t);
    mbar_init(o_full, 1);
    for (int s = 0; s < CE_NST; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], CE_EPI_WARPS); mbar_init(&a2_full[b], CE_EPI_WARPS); mbar_init(&a2_empty[b], 1); }
    mbar_init(acc_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&dsc_full[b], 1); mbar_init(&dsc_empty[b], CE_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, CE_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(o_full, CE_TILE_BYTES);
      tma_load_2d(smem + CeSmem::kO, map_o, o_full, 0, o0);
      for (int t = 0; t < n_tiles; ++t) {
        const int st = t % CE_NST; const uint32_t ph = (uint32_t)(t / CE_NST) & 1u;
        mbar_wait(&t_empty[st], ph ^ 1u);
        mbar_expect_tx(&t_full[st], CE_TILE_BYTES);
        tma_load_2d(smem + CeSmem::kT + st * CE_TILE_BYTES, map_t, &t_full[st], 0, (t_beg + t) * CT);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc1 = instr_desc_bf16(CT, CT, 0, 0);       // S[128,128] = O (K-major) x T^T (K-major)
      constexpr uint32_t idesc2 = instr_desc_bf16(CT, CE_E, 0, 1);     // acc[128,64] += W (K-major) x T (MN-major)
      constexpr uint32_t idesc3 = instr_desc_bf16(CT, CE_E, 1, 1);     // dsc[128,64]  = W^T (MN-major) x O (MN-major)
      const uint32_t so = smem_u32(smem + CeSmem::kO);
      auto issue_s = [&](int t) {
        const int st = t % CE_NST; const uint32_t ph = (uint32_t)(t / CE_NST) & 1u;
        const int b = t & 1; const uint32_t bph = (uint32_t)(t >> 1) & 1u;
        mbar_wait(&t_full[st], ph);
        mbar_wait(&s_empty[b], bph ^ 1u);
        tc_fence_after();
        const uint32_t stile = smem_u32(smem + CeSmem::kT + st * CE_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < CE_E / 16; ++k)
          mma_bf16_ss(tmem_base + b * CT, smem_desc_sw128(so + k * 32, 16, 1024), smem_desc_sw128(stile + k * 32, 16, 1024), idesc1, k > 0 ? 1u : 0u);
        mma_commit(&s_full[b]);
        if (!BWD) mma_commit(&t_empty[st]);      // forward: the tile is free once S is computed
      };
      mbar_wait(o_full, 0);
      issue_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_s(t + 1);
        if (BWD) {
          const int st = t % CE_NST;
          const int b = t & 1; const uint32_t bph = (uint32_t)(t >> 1) & 1u;
          mbar_wait(&a2_full[b], bph);
          tc_fence_after();
          const uint32_t sa2 = smem_u32(smem + CeSmem::kA2 + b * CE_A2_BYTES);
          const uint32_t stile = smem_u32(smem + CeSmem::kT + st * CE_TILE_BYTES);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_bf16_ss(tmem_base + CE_ACC_COL, smem_desc_sw128(sa2 + kb * (CT * 64 * 2) + k * 32, 16, 1024),
                          smem_desc_sw128(stile + (kb * 4 + k) * 2048, 64 * 64 * 2, 1024), idesc2, (t > 0 || kb > 0 || k > 0) ? 1u : 0u);
          if (FUSED) {
            // d_score partial of this tile: rows of the CTA are the reduction dimension.  The W tile is stored as two 64-column
            // blocks of [128 rows x 128 B]; read MN-major, a k-step of 16 rows advances by two 8-row swizzle atoms (2048 B) and
            // the second 64-column atom lies one block (128 x 128 B) further (LBO).
            mbar_wait(&dsc_empty[b], bph ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < CT / 16; ++k)
              mma_bf16_ss(tmem_base + CE_DSC_COL + b * CE_E, smem_desc_sw128(sa2 + k * 2048, CT * 64 * 2, 1024),
                          smem_desc_sw128(so + k * 2048, 64 * 64 * 2, 1024), idesc3, k > 0 ? 1u : 0u);
            mma_commit(&dsc_full[b]);
          }
          mma_commit(&a2_empty[b]);
          mma_commit(&t_empty[st]);
        }
      }
      if (BWD) mma_commit(acc_full);
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;                  // 0..15
    const int quad = warp & 3;                // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;                 // which 32 of the 128 streamed entities (0..3)
    const int et = threadIdx.x - 64;          // 0..511
    const int o = o0 + quad * 32 + lane;      // owner entity of this thread
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float scale = BWD ? ce_scale_dev(a.g_sum, a.g_mean, a.n_valid) : 0.f;

    // ---- attributes of this thread's loss row ----
    float o_lse_l2 = 0.f;
    const bool o_ok = (o < a.R) && (a.lm_rows[o] != 0.f);
    const int oc = o < a.R ? o : 0;
    const int o_i = oc / a.L, o_j = oc % a.L;
    const int o_label = (int)((a.user_offset + o_i) * a.S + o_j + 1);
    const bool o_lab_masked = (o_j + 1 < a.L) && (a.lm_cols[(a.user_offset + o_i) * a.L + o_j + 1] == 0.f);
    const uint32_t* o_mask = a.maskbits + (int64_t)o_i * a.Cw;       // mask words of the row's user
    if (BWD) o_lse_l2 = -(o_ok ? a.lse[oc] : INFINITY) * kLog2e;
    float run_m = -INFINITY, run_s = 0.f, lab_val = 0.f;

    // MODE 3: add the d_score partial of streamed tile t (TMEM lanes = its 128 columns) to global memory
    auto flush_dsc = [&](int t) {
      const int b = t & 1; const uint32_t bph = (uint32_t)(t >> 1) & 1u;
      mbar_wait(&dsc_full[b], bph);
      tc_fence_after();
      uint32_t raw[16];
      tmem_ld_32x16(tmem_base + lane_addr + (uint32_t)(CE_DSC_COL + b * CE_E + half * 16), raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dsc_empty[b]);
      const int col = (t_beg + t) * CT + quad * 32 + lane;
      if (col < a.C) {
        float* op = a.d_out2 + (int64_t)col * CE_E + half * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          red_add_v4(op + 4 * q, __uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                     __uint_as_float(raw[4 * q + 3]));
      }
    };

    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1; const uint32_t bph = (uint32_t)(t >> 1) & 1u;
      const int tt0 = (t_beg + t) * CT;                                  // first streamed entity of this tile
      // ---- debias of this warp's 32 columns: straight from global memory (the 32 lanes read the same addresses: one L1
      //      wavefront per load), issued before the wait for S so that its latency is hidden; no CTA-wide barrier per tile ----
      const int k0 = half * 32;                                           // offset of this warp's 32 streamed entities inside the tile
      const int c0 = tt0 + k0;
      float4 deb[8];
      if (c0 + 32 <= a.C) {
        const float4* dp = reinterpret_cast<const float4*>(a.debias + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) deb[q] = __ldg(dp + q);
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          deb[q].x = (c0 + 4 * q < a.C) ? __ldg(a.debias + c0 + 4 * q) : 0.f;
          deb[q].y = (c0 + 4 * q + 1 < a.C) ? __ldg(a.debias + c0 + 4 * q + 1) : 0.f;
          deb[q].z = (c0 + 4 * q + 2 < a.C) ? __ldg(a.debias + c0 + 4 * q + 2) : 0.f;
          deb[q].w = (c0 + 4 * q + 3 < a.C) ? __ldg(a.debias + c0 + 4 * q + 3) : 0.f;
        }
      }
      // 32 consecutive columns, c0 % 32 == 0.  Columns >= C carry mask bits (ce_maskbits_kernel) and a zero debias: they
      // become -1e4 like every masked entry, whose softmax weight underflows to exactly 0 against any real logit of the row.
      const uint32_t mw = (c0 < a.C) ? __ldg(o_mask + (c0 >> 5)) : 0xffffffffu;
      mbar_wait(&s_full[b], bph);
      tc_fence_after();
      if (BWD) mbar_wait(&a2_empty[b], bph ^ 1u);
      uint8_t* a2 = smem + CeSmem::kA2 + b * CE_A2_BYTES;
      {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(b * CT + k0), raw);
        tmem_ld_wait();
        float v[32];
        {
          // masked entries are rare (the 11 items of the row's user, padded slots): when no lane of the warp has one in this
          // block of 32 columns the per-logit selects are skipped altogether
          if (!__any_sync(0xffffffffu, mw != 0u)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              v[4 * q] = __uint_as_float(raw[4 * q]) - deb[q].x; v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) - deb[q].y;
              v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) - deb[q].z; v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) - deb[q].w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 d = deb[q];
              v[4 * q] = ((mw >> (4 * q)) & 1u) ? kNegMaskF : __uint_as_float(raw[4 * q]) - d.x;
              v[4 * q + 1] = ((mw >> (4 * q + 1)) & 1u) ? kNegMaskF : __uint_as_float(raw[4 * q + 1]) - d.y;
              v[4 * q + 2] = ((mw >> (4 * q + 2)) & 1u) ? kNegMaskF : __uint_as_float(raw[4 * q + 2]) - d.z;
              v[4 * q + 3] = ((mw >> (4 * q + 3)) & 1u) ? kNegMaskF : __uint_as_float(raw[4 * q + 3]) - d.w;
            }
          }
          const bool has_label = o_label >= c0 && o_label < c0 + 32;
          if (has_label) {                                                // the label column escapes the reject mask
            const int kl = o_label - c0;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k == kl) v[k] = o_lab_masked ? kNegMaskF : __uint_as_float(raw[k]) - reinterpret_cast<const float*>(deb)[k];
          }
          if (!BWD) {
            float cm = v[0];
#pragma unroll
            for (int k = 1; k < 32; ++k) cm = fmaxf(cm, v[k]);
            const float nm = fmaxf(run_m, cm);
            const float nml = -nm * kLog2e;
            float s = run_s * ex2_ftz(fmaf(run_m, kLog2e, nml));           // run_m = -inf at the start: 0 * 2^-inf = 0
#pragma unroll
            for (int k = 0; k < 32; ++k) s += ex2_ftz(fmaf(v[k], kLog2e, nml));
            run_m = nm; run_s = s;
            if (has_label) {
              const int kl = o_label - c0;
#pragma unroll
              for (int k = 0; k < 32; ++k) if (k == kl) lab_val = v[k];
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = ex2_ftz(fmaf(v[k], kLog2e, o_lse_l2)) * scale;   // o_lse = +inf for invalid rows -> 0
            if (o_ok && has_label) {
              const int kl = o_label - c0;
#pragma unroll
              for (int k = 0; k < 32; ++k) if (k == kl) v[k] -= scale;
            }
          }
        }
        if (BWD) {
          // ---- bf16 weight tile -> shared memory, K-major with the 128-byte swizzle the MMA descriptor expects ----
          const int m = quad * 32 + lane;                                  // row of the A operand
          const int kb = k0 >> 6;                                          // 64-wide k block
          const int kk0 = k0 & 63;                                         // 0 or 32
          uint8_t* rowp = a2 + kb * (CT * 64 * 2) + (m >> 3) * 1024 + (m & 7) * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int z = 0; z < 4; ++z) h2[z] = __floats2bfloat162_rn(v[q * 8 + 2 * z], v[q * 8 + 2 * z + 1]);
            const int chunk = (kk0 >> 3) + q;                              // 16-byte chunk index inside the 128-byte row
            *reinterpret_cast<uint4*>(rowp + ((chunk ^ (m & 7)) << 4)) = pk;
          }
        }
      }
      // S buffer consumed
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      if (BWD) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[b]);
        if (FUSED && t >= 1) flush_dsc(t - 1);
      }
    }
    if (FUSED) flush_dsc(n_tiles - 1);

    if (!BWD) {
      // combine the four column quarters of each row, then write the split's partial
      float* xm = attr_f;                 // reuse: [3 quarters][m | s | lab][128]  (tile attributes are dead)
      epi_bar_sync();
      if (half > 0) {
        float* q = xm + (half - 1) * 3 * CT + quad * 32 + lane;
        q[0] = run_m; q[CT] = run_s; q[2 * CT] = lab_val;
      }
      epi_bar_sync();
      if (half == 0 && o < a.R) {
        float nm = run_m;
#pragma unroll
        for (int h = 0; h < 3; ++h) nm = fmaxf(nm, xm[h * 3 * CT + quad * 32 + lane]);
        float s = 0.f, lab = lab_val;
        if (run_m > -INFINITY) s += run_s * __expf(run_m - nm);
#pragma unroll
        for (int h = 0; h < 3; ++h) {
          const float* q = xm + h * 3 * CT + quad * 32 + lane;
          if (q[0] > -INFINITY) s += q[CT] * __expf(q[0] - nm);
          lab += q[2 * CT];
        }
        const int64_t idx = (int64_t)blockIdx.y * a.R + o;
        a.part_m[idx] = nm; a.part_s[idx] = s; a.part_lab[idx] = lab;
      }
    } else {
      // flush the accumulator: thread owns entity o, 16 of the 64 output features
      mbar_wait(acc_full, 0);
      tc_fence_after();
      uint32_t raw[16];
      tmem_ld_32x16(tmem_base + lane_addr + (uint32_t)(CE_ACC_COL + half * 16), raw);
      tmem_ld_wait();
      if (o < a.R) {
        float* op = a.d_out + (int64_t)o * CE_E + half * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          red_add_v4(op + 4 * q, __uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                     __uint_as_float(raw[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, CE_TMEM_COLS);
}

// ---- prepass ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_prepass_cast_kernel(const float* __restrict__ prec, const float* __restrict__ score, int64_t n_prec,
                                                              int64_t n_score, bf16* __restrict__ prec_b, bf16* __restrict__ score_b,
                                                              const int64_t* __restrict__ ids_cols, const float* __restrict__ pop, int C,
                                                              float* __restrict__ debias) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < n_prec / 4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(prec)[i];
    uint2 q;
    *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(prec_b)[i] = q;
  }
  for (int64_t i = i0; i < n_score / 4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(score)[i];
    uint2 q;
    *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(score_b)[i] = q;
  }
  for (int64_t c = i0; c < C; c += stride) debias[c] = logf(pop[ids_cols[c]]);
}

// one CTA per row-user, one warp per 32-column word (lane k = column 32 w + k: coalesced id loads, the word is a ballot):
// bit k = col-pad masked OR id of column in the user's S ids
__global__ void __launch_bounds__(256) ce_maskbits_kernel(const int64_t* __restrict__ ids_rows, const int64_t* __restrict__ ids_cols,
                                                          const float* __restrict__ lm_cols, int B, int S, int L, int C, int Cw,
                                                          uint32_t* __restrict__ bits) {
  const int i = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t rid[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) rid[k] = (k < S) ? __ldg(ids_rows + (int64_t)i * S + k) : __ldg(ids_rows + (int64_t)i * S);
  // Item ids are small non-negative numbers: when all ids of this row-user fit 32 bits, a column whose id has a zero high word is
  // compared on the low words only (12 instead of 34 compare instructions per column); anything else takes the exact 64-bit
  // compare, so the masks stay bit-exact for arbitrary int64 ids.
  uint32_t rlo[17];
  bool rows_narrow = true;
#pragma unroll
  for (int k = 0; k < 17; ++k) { rlo[k] = (uint32_t)rid[k]; rows_narrow = rows_narrow && ((uint64_t)rid[k] >> 32) == 0ull; }
  // four independent words per iteration: their loads overlap; blockIdx.y strides over the iterations (the kernel is latency
  // bound: one CTA per user walking all of its columns left most of the chip idle)
  for (int w0 = warp + 32 * (int)blockIdx.y; w0 < Cw; w0 += 32 * (int)gridDim.y) {
    int64_t id[4]; float lmv[4]; int pp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = (w0 + 8 * q) * 32 + lane;
      id[q] = 0; lmv[q] = 1.f; pp[q] = L;
      if (w0 + 8 * q < Cw && c < C) {
        const int u = c / S, p = c - u * S;
        pp[q] = p;
        id[q] = ids_cols[c];
        if (p < L) lmv[q] = lm_cols[(int64_t)u * L + p];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int w = w0 + 8 * q;
      if (w >= Cw) break;                                  // warp-uniform
      const int c = w * 32 + lane;
      bool m = true;                                       // columns beyond C: masked (never read as real columns)
      if (c < C) {
        m = (pp[q] < L) && (lmv[q] == 0.f);
        bool hit = false;
        if (rows_narrow && ((uint64_t)id[q] >> 32) == 0ull) {
          const uint32_t lo = (uint32_t)id[q];
          if (S <= 11) {                                   // the shipped max_seq_len = 10 (slots >= S repeat slot 0)
#pragma unroll
            for (int k = 0; k < 11; ++k) hit |= (rlo[k] == lo);
          } else {
#pragma unroll
            for (int k = 0; k < 17; ++k) hit |= (rlo[k] == lo);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 17; ++k) hit |= (rid[k] == id[q]);
        }
        m = m || hit;
      }
      const uint32_t word = __ballot_sync(0xffffffffu, m);
      if (lane == 0) bits[(int64_t)i * Cw + w] = word;
    }
  }
}

__global__ void __launch_bounds__(256) ce_combine_kernel(const float* __restrict__ part_m, const float* __restrict__ part_s,
                                                         const float* __restrict__ part_lab, int splits, int R,
                                                         const float* __restrict__ lm_rows, float* __restrict__ lse_out,
                                                         double* __restrict__ acc_sum, int* __restrict__ acc_cnt, float* loss_sum,
                                                         int32_t* n_valid, float* loss) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f; int cnt = 0;
  if (r < R) {
    if (lm_rows[r] != 0.f) {
      float m = -INFINITY;
      for (int k = 0; k < splits; ++k) m = fmaxf(m, part_m[(int64_t)k * R + r]);
      float s = 0.f, lab = 0.f;
      for (int k = 0; k < splits; ++k) {
        const float mk = part_m[(int64_t)k * R + r];
        if (mk > -INFINITY) s += part_s[(int64_t)k * R + r] * expf(mk - m);
        lab += part_lab[(int64_t)k * R + r];
      }
      const float lse = m + logf(s);
      lse_out[r] = lse;
      contrib = lse - lab; cnt = 1;
    } else {
      lse_out[r] = 0.f;
    }
  }
  __shared__ float rs[8]; __shared__ int rc[8];
  contrib = warp_sum(contrib);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = contrib; rc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0; int n = 0;
    for (int k = 0; k < 8; ++k) { t += rs[k]; n += rc[k]; }
    if (n) { atomicAdd(acc_sum, t); atomicAdd(acc_cnt, n); }
    // the CTA that takes the last ticket sees every partial: it writes the results (no separate finalize launch)
    __threadfence();
    int* done = acc_cnt + 1;
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      __threadfence();
      const double s = *reinterpret_cast<volatile double*>(acc_sum);
      const int nn = *reinterpret_cast<volatile int*>(acc_cnt);
      if (loss_sum) *loss_sum = (float)s;
      if (n_valid) *n_valid = nn;
      if (loss) *loss = (float)(s / (double)nn);
    }
  }
}

// expand the mask bits into the probe format of iisan_inbatch_ce_masks (bit0 masked, bit2 label, bit3 row valid)
__global__ void ce_maskprobe_kernel(const uint32_t* __restrict__ bits, int B, int L, int S, int C, int Cw, int64_t user_offset,
                                    const float* __restrict__ lm_rows, const float* __restrict__ lm_cols, uint8_t* __restrict__ out) {
  const int row = blockIdx.x;
  const int i = row / L, j = row % L;
  const int64_t label = (user_offset + i) * S + j + 1;
  const bool lab_masked = (j + 1 < L) && (lm_cols[(user_offset + i) * L + j + 1] == 0.f);
  const bool valid = lm_rows[row] != 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    bool m = (bits[(int64_t)i * Cw + (c >> 5)] >> (c & 31)) & 1u;
    uint8_t b = 0;
    if (c == label) { m = lab_masked; b |= 4; }
    if (m) b |= 1;
    if (valid) b |= 8;
    out[(int64_t)row * C + c] = b;
  }
}

// ---- host side --------------------------------------------------------------------------------------------
int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_inner, int box_outer);

struct CeFastLayout {
  float* lse; double* acc_sum; int* acc_cnt;
  bf16* prec_b; bf16* score_b; float* debias; uint32_t* maskbits;
  float *part_m, *part_s, *part_lab;
  int splits_fwd; size_t bytes;
  CeFastLayout(const iisan_ce_desc& d, void* ws) {
    Arena a(ws);
    const size_t R = (size_t)d.row_users * d.seq_len, S = d.seq_len + 1, C = (size_t)d.col_users * S, Cw = (C + 31) / 32;
    lse = a.take<float>(R);
    acc_sum = a.take<double>(1); acc_cnt = a.take<int>(2);      // acc_cnt[1]: ticket counter of ce_combine_kernel
    prec_b = a.take<bf16>(R * d.emb); score_b = a.take<bf16>(C * d.emb);
    debias = a.take<float>(C); maskbits = a.take<uint32_t>((size_t)d.row_users * Cw);
    splits_fwd = ce_fast_splits((int)((R + CT - 1) / CT), (int)((C + CT - 1) / CT));
    part_m = a.take<float>(R * splits_fwd); part_s = a.take<float>(R * splits_fwd); part_lab = a.take<float>(R * splits_fwd);
    bytes = a.off;
  }
};

// One CTA per SM (512 TMEM columns, 129 KB smem).  Short column pools: the split count keeps owner_tiles * splits within ONE
// wave of 148 CTAs (a 160-CTA grid would run a second, almost empty wave).  Long pools (the global negative pool of a
// data-parallel step: W times the columns): one wave of 40 x 3 = 120 CTAs leaves 28 SMs idle for the whole kernel, so the pool
// is cut into k waves' worth of splits when the modelled time -- waves x (tiles per CTA + one tile-equivalent of CTA prologue:
// barrier init, tensor-memory allocation, owner-tile load) -- drops by at least 8 %.
int ce_fast_splits(int owner_tiles, int stream_tiles) {
  auto shape = [&](int s, int* per_out) {          // effective split count for a requested one
    if (s > stream_tiles) s = stream_tiles;
    if (s < 1) s = 1;
    const int per = (stream_tiles + s - 1) / s;
    *per_out = per;
    return (stream_tiles + per - 1) / per;
  };
  int per1 = 0;
  const int s1 = shape(148 / (owner_tiles < 1 ? 1 : owner_tiles), &per1);
  int best = s1;
  const int waves1 = (owner_tiles * s1 + 147) / 148;
  double best_cost = (double)waves1 * (per1 + 1);
  const double bar = 0.92 * best_cost;
  const char* env = getenv("IISAN_B200_CE_ONE_WAVE");      // A/B switch, read per call (scripts/ce_gather_bench.py toggles it)
  const bool one_wave = env && env[0] == '1';
  for (int k = 2; k <= 8 && !one_wave; ++k) {
    int per = 0;
    const int s = shape(k * 148 / (owner_tiles < 1 ? 1 : owner_tiles), &per);
    // measured (scripts/ce_gather_bench.py; W = 2: 168 -> 174 us with 8 tiles per CTA, W = 4: 269 -> 262 us with 16, W = 8: 480 -> 447 us
    // with 32): short CTAs lose to their prologue.  IISAN_B200_CE_MIN_TILES lowers the bar (tests force the multi-wave grid at W = 2).
    const char* mt = getenv("IISAN_B200_CE_MIN_TILES");
    if (per < (mt ? atoi(mt) : 16)) break;
    const int waves = (owner_tiles * s + 147) / 148;
    const double cost = (double)waves * (per + 1);
    if (cost < bar && cost < best_cost) { best = s; best_cost = cost; }
  }
  return best;
}

int ce_fast_supported(const iisan_ce_desc& d) {
  const int64_t C = (int64_t)d.col_users * (d.seq_len + 1);
  return d.emb == CE_E && d.seq_len + 1 <= 17 && C < (int64_t)1 << 30;
}

size_t ce_fast_workspace_bytes(const iisan_ce_desc& d) {
  CeFastLayout L(d, nullptr);
  return L.bytes;
}

static int fill_args(const iisan_ce_desc& d, const CeFastLayout& W, CeTileArgs* A, const float* lm_rows, const float* lm_cols) {
  const int R = d.row_users * d.seq_len, S = d.seq_len + 1, C = d.col_users * S;
  IISAN_TRY(make_tensor_map_bf16(&A->map_prec, W.prec_b, R, d.emb, d.emb, CE_E, CT));
  IISAN_TRY(make_tensor_map_bf16(&A->map_score, W.score_b, C, d.emb, d.emb, CE_E, CT));
  A->R = R; A->L = d.seq_len; A->S = S; A->B = d.row_users; A->C = C; A->Cw = (C + 31) / 32;
  A->user_offset = d.user_offset; A->lm_rows = lm_rows; A->lm_cols = lm_cols; A->debias = W.debias; A->maskbits = W.maskbits;
  A->lse = W.lse; A->g_sum = nullptr; A->g_mean = nullptr; A->n_valid = nullptr;
  A->part_m = W.part_m; A->part_s = W.part_s; A->part_lab = W.part_lab; A->d_out = nullptr;
  return IISAN_OK;
}

template <int MODE>
static int launch_tile(const CeTileArgs& A, int owner_tiles, int splits, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(ce_tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CeSmem::kTotal));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  { LaunchScope ls_(IISAN_K_CE, st); ce_tile_kernel<MODE><<<dim3(owner_tiles, splits), CE_THREADS, CeSmem::kTotal, st>>>(A); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

static int run_prepass(const iisan_ce_desc& d, const CeFastLayout& W, const float* prec, const float* score, const int64_t* ids_rows,
                       const int64_t* ids_cols, const float* lm_cols, const float* pop, cudaStream_t st) {
  const int R = d.row_users * d.seq_len, S = d.seq_len + 1, C = d.col_users * S, Cw = (C + 31) / 32;
  if (prec) {
    LaunchScope ls_(IISAN_K_CE, st);
    ce_prepass_cast_kernel<<<148 * 2, 256, 0, st>>>(prec, score, (int64_t)R * d.emb, (int64_t)C * d.emb, W.prec_b, W.score_b, ids_cols, pop, C, W.debias);
  }
  IISAN_LAUNCH_OK();
  {
    LaunchScope ls_(IISAN_K_CE, st);
    ce_maskbits_kernel<<<dim3((unsigned)d.row_users, (unsigned)std::max(1, std::min(8, (Cw + 31) / 32))), 256, 0, st>>>(ids_rows, ids_cols, lm_cols, d.row_users, S, d.seq_len, C, Cw, W.maskbits);
  }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int ce_fast_forward(const iisan_ce_desc& d, const float* prec, const float* score, const int64_t* ids_rows, const int64_t* ids_cols,
                    const float* lm_rows, const float* lm_cols, const float* pop, void* ws, float* loss_sum, int32_t* n_valid,
                    float* loss, cudaStream_t st) {
  CeFastLayout W(d, ws);
  const int R = d.row_users * d.seq_len, S = d.seq_len + 1, C = d.col_users * S;
  IISAN_CUDA_OK(cudaMemsetAsync(W.acc_sum, 0, 256 + 2 * sizeof(int), st));
  IISAN_TRY(run_prepass(d, W, prec, score, ids_rows, ids_cols, lm_cols, pop, st));
  CeTileArgs A;
  IISAN_TRY(fill_args(d, W, &A, lm_rows, lm_cols));
  const int owner_tiles = (R + CT - 1) / CT;
  A.tiles_stream = (C + CT - 1) / CT;
  A.tiles_per_split = (A.tiles_stream + W.splits_fwd - 1) / W.splits_fwd;
  IISAN_TRY(launch_tile<0>(A, owner_tiles, W.splits_fwd, st));
  { LaunchScope ls_(IISAN_K_CE, st); ce_combine_kernel<<<(R + 255) / 256, 256, 0, st>>>(W.part_m, W.part_s, W.part_lab, W.splits_fwd, R, lm_rows, W.lse, W.acc_sum, W.acc_cnt, loss_sum, n_valid, loss); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int ce_fast_backward(const iisan_ce_desc& d, const float* lm_rows, const float* lm_cols, void* ws, const float* g_sum,
                     const float* g_mean, const int32_t* n_valid, float* d_prec, float* d_score, cudaStream_t st) {
  CeFastLayout W(d, ws);
  const int R = d.row_users * d.seq_len, S = d.seq_len + 1, C = d.col_users * S;
  IISAN_CUDA_OK(cudaMemsetAsync(d_prec, 0, (size_t)R * d.emb * sizeof(float), st));
  IISAN_CUDA_OK(cudaMemsetAsync(d_score, 0, (size_t)C * d.emb * sizeof(float), st));
  CeTileArgs A;
  IISAN_TRY(fill_args(d, W, &A, lm_rows, lm_cols));
  A.g_sum = g_sum; A.g_mean = g_mean; A.n_valid = n_valid;
  const int row_tiles = (R + CT - 1) / CT, col_tiles = (C + CT - 1) / CT;
  {      // one pass: d_prec in the CTA's accumulator, d_score by per-tile reductions
    const int splits = ce_fast_splits(row_tiles, col_tiles);
    A.tiles_stream = col_tiles; A.tiles_per_split = (col_tiles + splits - 1) / splits; A.d_out = d_prec; A.d_out2 = d_score;
    IISAN_TRY(launch_tile<3>(A, row_tiles, splits, st));
  }
  return IISAN_OK;
}

int ce_fast_masks(const iisan_ce_desc& d, const int64_t* ids_rows, const int64_t* ids_cols, const float* lm_rows, const float* lm_cols,
                  void* ws, uint8_t* out, cudaStream_t st) {
  CeFastLayout W(d, ws);
  const int R = d.row_users * d.seq_len, S = d.seq_len + 1, C = d.col_users * S, Cw = (C + 31) / 32;
  IISAN_TRY(run_prepass(d, W, nullptr, nullptr, ids_rows, ids_cols, lm_cols, nullptr, st));
  { LaunchScope ls_(IISAN_K_CE, st); ce_maskprobe_kernel<<<R, 256, 0, st>>>(W.maskbits, d.row_users, d.seq_len, S, C, Cw, d.user_offset, lm_rows, lm_cols, out); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan
