// Fused side-adapter chain (fast mode, symmetric configurations: d_text == d_img, r == 64, every tower active in every
// stage, bf16 cached states).  One CTA owns 128 items of ONE tower and walks all A stages without leaving the SM.
//
// forward                                                          reference
//   x_0       = fuse(h_0, 0)                                       gated fusion   CC/model/model.py:319-326 (mm: :335-337)
//   z_s       = relu(x_s Wd_s^T + bd_s)                            AdapterBlock   CC/model/modules.py:113-116
//   last_s    = z_s Wu_s^T + bu_s + x_s
//   x_{s+1}   = fuse(h_{s+1}, last_s)                              (mm tower: last_s + g h_cv + (1-g) h_text)
// backward (data / gate / bias gradients; weight gradients are reductions over ALL items and stay split-K GEMMs over the
// stashes written here)
//   dz_s      = (dy_s Wu_s) * (z_s > 0)                            dy_s = d last_s
//   dx_s      = dy_s + dz_s Wd_s
//   dgate_s  += sum dx_s * (h_s - last_{s-1})                      (mm: h_cv - h_text), times g(1-g)/0.1 ; the intra-modal towers
//                                                                  use (h_s - x_s) g/0.1 instead (x_s is stashed, last_{s-1} is not)
//   dy_{s-1}  = (1 - g_s) dx_s                                     (mm: dx_s)
//
// The running state never exists as a whole on chip (128 x 768 fp32 exceeds TMEM); it streams through 64-column chunks.
// For chunk c the per-chunk product (U_c = z_s Wu_s[c]^T, resp. dz_s Wd_s[:, c]) is one tcgen05.mma group into a
// double-buffered TMEM accumulator; the 8 epilogue warps add bias / residual, fuse the next hidden-state chunk and write
// the bf16 result 128B-swizzled into shared memory, where it is at once (i) the A operand of the next stage's accumulation
// over the chunks (z_{s+1}, resp. dz_{s-1}, accumulates in TMEM) and (ii) the source of a TMA store into the stash.
// All global traffic is TMA: the cached states are read in place from the caller's [N, layers, d] tensors (row pitch
// layers*d: only selected layers are ever touched), the stashes are written by bulk tensor stores and read back one stage
// later through the same ring (ordered by cp.async.bulk.wait_group on the issuing thread + the ring's mbarrier chain,
// whose depth of ~2 chunks is far below the 12 chunks between a write and its re-read).
//
// warp roles: 0 weight TMA producer | 1 TMEM allocator + MMA issuer + TMA stores | 2 data TMA producer | 3..10 epilogue
#include <cstdlib>

#include "common.cuh"
#include "launch.cuh"
#include "san_chain.cuh"
#include "umma.cuh"

namespace iisan {

using namespace umma;
using bf16 = __nv_bfloat16;

constexpr int CH_ROWS = 128;
constexpr int CH_CW = 64;                    // chunk width (columns) == one 128-byte swizzle atom
constexpr int CH_R = 64;                     // adapter bottleneck handled by these kernels
constexpr int CH_NW = 3;                     // weight ring (units of two 8 KB chunks)
constexpr int CH_NH = 5;                     // data ring (16 KB tiles)
constexpr int CH_THREADS = 352;              // 11 warps
constexpr int CH_TILE_BYTES = CH_ROWS * CH_CW * 2;   // 16 KB
constexpr int CH_W_BYTES = CH_CW * CH_R * 2;         // 8 KB
constexpr int CH_TMEM_COLS = 256;
constexpr int CH_ZACC = 0, CH_UACC = 64;     // TMEM columns: z (dz) accumulator, two chunk accumulators

// ---- optional wait-time accounting (build variant "trace": -DIISAN_CHAIN_TRACE, see iisan_b200/build.py and
// scripts/chain_trace.py).  The middle CTA of every tower adds, per role and wait site, the cycles spent in the wait and the
// number of waits.  Roles: 0 weight producer, 1 MMA / store thread, 2 data producer, 3 one epilogue thread.  In the default
// build the macros vanish and the kernels are unchanged. ----
#ifdef IISAN_CHAIN_TRACE
constexpr int CH_TR_ROLES = 4, CH_TR_SITES = 16;
__device__ unsigned long long g_chain_trace[2][3][CH_TR_ROLES][CH_TR_SITES][2];   // [fwd|bwd][tower][role][site]{cycles, count}
#define CH_T0() const long long ch_t0_ = clock64()
#define CH_T1(pass, role, site)                                                                   \
  do {                                                                                            \
    if (blockIdx.x == gridDim.x / 2) {                                                            \
      g_chain_trace[pass][blockIdx.y][role][site][0] += (unsigned long long)(clock64() - ch_t0_); \
      g_chain_trace[pass][blockIdx.y][role][site][1] += 1ull;                                     \
    }                                                                                             \
  } while (0)
#define CH_T1E(pass, site) do { if (threadIdx.x == 96) CH_T1(pass, 3, site); } while (0)        /* first epilogue thread */
#define CH_TT0() const long long ch_tt0_ = clock64()
#define CH_TT1(pass, role)                                                                                       \
  do {                                                                                                           \
    if (blockIdx.x == gridDim.x / 2) {                                                                           \
      g_chain_trace[pass][blockIdx.y][role][CH_TR_SITES - 1][0] += (unsigned long long)(clock64() - ch_tt0_);    \
      g_chain_trace[pass][blockIdx.y][role][CH_TR_SITES - 1][1] += 1ull;                                         \
    }                                                                                                            \
  } while (0)
#else
#define CH_TT0() do {} while (0)
#define CH_TT1(pass, role) do {} while (0)
#define CH_T0() do {} while (0)
#define CH_T1(pass, role, site) do {} while (0)
#define CH_T1E(pass, site) do {} while (0)
#endif

struct ChainSmem {
  static constexpr int kZ = 0;                                   // z / dz operand [128 x 64]
  static constexpr int kXk = kZ + CH_TILE_BYTES;                 // 2 chunk operands (x_{s+1}[c], resp. dy_{s-1}[c])
  static constexpr int kLk = kXk + 2 * CH_TILE_BYTES;            // 2 staging tiles of last_s[c] (forward)
  static constexpr int kW = kLk + 2 * CH_TILE_BYTES;             // CH_NW x (chunk | chunk)
  static constexpr int kH = kW + CH_NW * 2 * CH_W_BYTES;         // CH_NH data tiles
  static constexpr int kBar = kH + CH_NH * CH_TILE_BYTES;
  static constexpr int kTotal = kBar + 512 + 1024;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_3() { asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

// sum over the 32 lanes of v[k] for each k: lane l ends with the total of column l (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? v[k] : v[k + off];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[k] = (upper ? v[k + off] : v[k]) + recv;
    }
  }
  return v[0];
}

struct ChainBars {
  uint64_t *w_full, *w_empty, *h_full, *h_empty, *xk_full, *xk_empty, *u_full, *u_empty, *z_full, *z_ready;
  uint32_t* tmem_slot;
  __device__ explicit ChainBars(uint8_t* smem) {
    uint64_t* b = reinterpret_cast<uint64_t*>(smem + ChainSmem::kBar);
    w_full = b; w_empty = w_full + CH_NW; h_full = w_empty + CH_NW; h_empty = h_full + CH_NH;
    xk_full = h_empty + CH_NH; xk_empty = xk_full + 2; u_full = xk_empty + 2; u_empty = u_full + 2;
    z_full = u_empty + 2; z_ready = z_full + 1;
    tmem_slot = reinterpret_cast<uint32_t*>(z_ready + 1);
  }
  __device__ void init(int epi_warps = 8) {
    for (int i = 0; i < CH_NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < CH_NH; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], epi_warps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&xk_full[i], epi_warps); mbar_init(&xk_empty[i], 1); mbar_init(&u_full[i], 1); mbar_init(&u_empty[i], epi_warps); }
    mbar_init(z_full, 1); mbar_init(z_ready, epi_warps);
    fence_barrier_init();
  }
};

// ================================================================================================================
// forward
// ================================================================================================================
// 16 epilogue warps: four per TMEM lane quadrant, each thread owns one row and 16 of the chunk's 64 columns (the epilogue is
// latency bound: twice the warps in flight hide the mbarrier / TMEM / shared-memory round trips of each other)
constexpr int CF_EPI_WARPS = 16;
constexpr int CF_THREADS = 96 + 32 * CF_EPI_WARPS;
constexpr int CF_NCOL = 64 / (CF_EPI_WARPS / 4);     // columns per thread

__global__ void __launch_bounds__(CF_THREADS, 1) san_chain_fwd_kernel(const __grid_constant__ ChainArgs a) {
  const ChainTower& T = a.tower[blockIdx.y];
  const bool is_mm = (T.mode == 1);
  const int NC = a.d / CH_CW;
  const int A = a.n_stages;
  const int m0 = blockIdx.x * CH_ROWS;
  const int NP = a.n_pad;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  ChainBars B(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&T.map_wd); tma_prefetch_desc(&T.map_wu); tma_prefetch_desc(&T.map_h); tma_prefetch_desc(&T.map_x);
    tma_prefetch_desc(&T.map_last); tma_prefetch_desc(&T.map_z);
    if (is_mm) tma_prefetch_desc(&T.map_h2);
    B.init(CF_EPI_WARPS);
  }
  if (warp == 1) tmem_alloc(B.tmem_slot, CH_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;

  // unit u = (sv + 1) * NC + c, sv = -1 .. A-1 : first half Wu_sv[c] (sv >= 0), second half Wd_{sv+1}[c] (sv+1 < A)
  const int n_units = (A + 1) * NC;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (elect_one()) {
      for (int u = 0; u < n_units; ++u) {
        const int sv = u / NC - 1, c = u % NC;
        const int slot = u % CH_NW; const uint32_t ph = (uint32_t)(u / CH_NW) & 1u;
        const bool has_wu = sv >= 0, has_wd = sv + 1 < A;
        { CH_T0(); mbar_wait(&B.w_empty[slot], ph ^ 1u); CH_T1(0, 0, 0); }
        uint8_t* dst = smem + ChainSmem::kW + slot * 2 * CH_W_BYTES;
        mbar_expect_tx(&B.w_full[slot], (has_wu ? CH_W_BYTES : 0) + (has_wd ? CH_W_BYTES : 0));
        if (has_wu) tma_load_2d(dst, &T.map_wu, &B.w_full[slot], 0, sv * a.d + c * CH_CW);                 // Wu_sv rows [c*64, +64), all r
        if (has_wd) tma_load_2d(dst + CH_W_BYTES, &T.map_wd, &B.w_full[slot], c * CH_CW, (sv + 1) * CH_R);  // Wd_{sv+1}: all r rows, chunk cols
      }
    }
  } else if (warp == 2) {
    // ===================== data producer =====================
    if (elect_one()) {
      int slot = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* m, int col, int row) {
        { CH_T0(); mbar_wait(&B.h_empty[slot], ph ^ 1u); CH_T1(0, 2, 0); }
        mbar_expect_tx(&B.h_full[slot], CH_TILE_BYTES);
        tma_load_2d(smem + ChainSmem::kH + slot * CH_TILE_BYTES, m, &B.h_full[slot], col, row);
        if (++slot == CH_NH) { slot = 0; ph ^= 1u; }
      };
      // consumption order of the epilogue: x_0[c] needs h_0[c] ; chunk c of stage s needs the residual x_s[c] (stored by this
      // CTA one stage earlier) and, unless s is the last stage, h_{s+1}[c]
      for (int c = 0; c < NC; ++c) {
        load(&T.map_h, T.layer[0] * a.d + c * CH_CW, m0);
        if (is_mm) load(&T.map_h2, T.layer2[0] * a.d + c * CH_CW, m0);
      }
      for (int s = 0; s < A; ++s) {
        for (int c = 0; c < NC; ++c) {
          load(&T.map_x, c * CH_CW, s * NP + m0);
          if (s + 1 < A) {
            load(&T.map_h, T.layer[s + 1] * a.d + c * CH_CW, m0);
            if (is_mm) load(&T.map_h2, T.layer2[s + 1] * a.d + c * CH_CW, m0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer + TMA stores =====================
    if (elect_one()) {
      CH_TT0();
      constexpr uint32_t idesc = instr_desc_bf16(CH_ROWS, 64, 0, 0);   // [128 x 64] (+)= A (K-major) x B^T (K-major), K = 64
      const uint32_t sz = smem_u32(smem + ChainSmem::kZ);
      int n_x = 0, n_u = 0;
      // chunk c of x_{sx} (and last_{sx-1} when staged) is in xk[b] / lk[b]: store it, feed z_acc (+)= xk x Wd_{sx}[c]^T
      auto down = [&](int unit, int c, int sx, bool with_last) {
        const int b = n_x & 1; const uint32_t ph = (uint32_t)(n_x >> 1) & 1u;
        const int slot = unit % CH_NW;
        { CH_T0(); mbar_wait(&B.xk_full[b], ph); CH_T1(0, 1, 0); }
        tc_fence_after();
        const uint8_t* xk = smem + ChainSmem::kXk + b * CH_TILE_BYTES;
        tma_store_2d(&T.map_x, xk, c * CH_CW, sx * NP + m0);
        if (with_last) tma_store_2d(&T.map_last, smem + ChainSmem::kLk + b * CH_TILE_BYTES, c * CH_CW, (sx - 1) * NP + m0);
        bulk_commit();
        const uint32_t sx_a = smem_u32(xk);
        const uint32_t sw = smem_u32(smem + ChainSmem::kW + slot * 2 * CH_W_BYTES + CH_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss(tmem_base + CH_ZACC, smem_desc_sw128(sx_a + k * 32, 16, 1024), smem_desc_sw128(sw + k * 32, 16, 1024), idesc, (c > 0 || k > 0) ? 1u : 0u);
        { CH_T0(); bulk_wait_read0(); CH_T1(0, 1, 1); }   // the stores have read their tiles: the buffers are free once the MMAs retire
        { CH_T0(); bulk_wait_3(); CH_T1(0, 1, 2); }       // stores older than 3 chunks are complete in memory (re-read 9+ chunks later)
        { CH_T0(); mma_commit(&B.xk_empty[b]); mma_commit(&B.w_empty[slot]); CH_T1(0, 1, 3); }
        ++n_x;
      };
      // final stage: only last_{A-1}[c] sits in lk[b]
      auto flush_last = [&](int c) {
        const int b = n_x & 1; const uint32_t ph = (uint32_t)(n_x >> 1) & 1u;
        { CH_T0(); mbar_wait(&B.xk_full[b], ph); CH_T1(0, 1, 4); }
        tma_store_2d(&T.map_last, smem + ChainSmem::kLk + b * CH_TILE_BYTES, c * CH_CW, (A - 1) * NP + m0);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&B.xk_empty[b]);
        ++n_x;
      };
      for (int c = 0; c < NC; ++c) {                 // x_0 chunks arrive from the epilogue warps
        const int unit = c; const uint32_t wph = (uint32_t)(unit / CH_NW) & 1u;
        { CH_T0(); mbar_wait(&B.w_full[unit % CH_NW], wph); CH_T1(0, 1, 5); }
        down(unit, c, 0, false);
      }
      mma_commit(B.z_full);
      for (int s = 0; s < A; ++s) {
        const bool more = s + 1 < A;
        { CH_T0(); mbar_wait(B.z_ready, (uint32_t)s & 1u); CH_T1(0, 1, 6); }
        tc_fence_after();
        tma_store_2d(&T.map_z, smem + ChainSmem::kZ, 0, s * NP + m0);   // z_s stash
        bulk_commit();
        for (int c = 0; c < NC; ++c) {
          const int unit = (s + 1) * NC + c; const int slot = unit % CH_NW; const uint32_t wph = (uint32_t)(unit / CH_NW) & 1u;
          const int b = n_u & 1; const uint32_t uph = (uint32_t)(n_u >> 1) & 1u;
          { CH_T0(); mbar_wait(&B.w_full[slot], wph); CH_T1(0, 1, 5); }
          { CH_T0(); mbar_wait(&B.u_empty[b], uph ^ 1u); CH_T1(0, 1, 7); }
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + ChainSmem::kW + slot * 2 * CH_W_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_bf16_ss(tmem_base + CH_UACC + b * 64, smem_desc_sw128(sz + k * 32, 16, 1024), smem_desc_sw128(sw + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
          mma_commit(&B.u_full[b]);
          ++n_u;
          if (!more) {
            mma_commit(&B.w_empty[slot]);                        // last stage: only Wu lives in the slot
            if (c >= 1) flush_last(c - 1);
          } else if (c >= 1) {
            down(unit - 1, c - 1, s + 1, T.store_last != 0);     // overlap: the previous chunk's down-projection
          }
        }
        if (more) { down((s + 1) * NC + NC - 1, NC - 1, s + 1, T.store_last != 0); mma_commit(B.z_full); }
        else flush_last(NC - 1);
      }
      bulk_wait_0();
      CH_TT1(0, 1);
    }
  } else {
    // ===================== epilogue warps =====================
    constexpr int NCOL = CF_NCOL, NQ = NCOL / 8;   // NQ 16-byte groups of 8 bf16 per thread and tile
    CH_TT0();
    const int ew = warp - 3;                  // 0..15
    const int quad = warp & 3;                // TMEM lane quadrant
    const int cq = ew >> 2;                   // which NCOL of the chunk's 64 columns
    const int m = quad * 32 + lane;           // row inside the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int sw_row = (m >> 3) * 1024 + (m & 7) * 128;      // byte offset of row m inside a swizzled [128 x 64] bf16 tile
    int n_x = 0, n_u = 0;

    auto put_tile = [&](uint8_t* tile_base, const float* v) {
      uint8_t* tile = tile_base + sw_row;
#pragma unroll
      for (int q = 0; q < NQ; ++q) *reinterpret_cast<uint4*>(tile + (((cq * NQ + q) ^ (m & 7)) << 4)) = pack8(v + q * 8);
    };
    // x chunk (or, in the final stage, the last chunk) -> swizzled shared memory; signals the MMA / store thread
    auto emit = [&](const float (&xv)[NCOL], bool is_x) {
      const int b = n_x & 1; const uint32_t ph = (uint32_t)(n_x >> 1) & 1u;
      { CH_T0(); mbar_wait(&B.xk_empty[b], ph ^ 1u); CH_T1E(0, 0); }
      put_tile(smem + (is_x ? ChainSmem::kXk : ChainSmem::kLk) + b * CH_TILE_BYTES, xv);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.xk_full[b]);
      ++n_x;
    };
    int h_slot = 0; uint32_t h_ph = 0;
    auto read_h = [&](float* hv) {            // this thread's NCOL columns of the next ring tile
      { CH_T0(); mbar_wait(&B.h_full[h_slot], h_ph); CH_T1E(0, 1); }
      const uint8_t* tile = smem + ChainSmem::kH + h_slot * CH_TILE_BYTES + sw_row;
#pragma unroll
      for (int q = 0; q < NQ; ++q) unpack8(*reinterpret_cast<const uint4*>(tile + (((cq * NQ + q) ^ (m & 7)) << 4)), hv + q * 8);
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.h_empty[h_slot]);
      if (++h_slot == CH_NH) { h_slot = 0; h_ph ^= 1u; }
    };
    auto tmem_cols = [&](uint32_t col0, float* out) {      // NCOL accumulator columns of this thread's row
      uint32_t raw[NCOL];
      tmem_ld_32x16(tmem_base + lane_addr + col0 + (uint32_t)(cq * NCOL), raw);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < NCOL; ++k) out[k] = __uint_as_float(raw[k]);
    };
    static_assert(NCOL == 16, "tmem_ld_32x16");

    // ---- x_0 = fuse(h_0, 0) ----
    {
      const float g = gate_value(T.gate[0]);
      const float omg = 1.0f - g;
      for (int c = 0; c < NC; ++c) {
        float hv[NCOL], xv[NCOL];
        read_h(hv);
        if (is_mm) {
          float h2[NCOL];
          read_h(h2);
#pragma unroll
          for (int k = 0; k < NCOL; ++k) xv[k] = fmaf(g, hv[k], omg * h2[k]);
        } else {
#pragma unroll
          for (int k = 0; k < NCOL; ++k) xv[k] = g * hv[k];
        }
        emit(xv, true);
      }
    }
    for (int s = 0; s < A; ++s) {
      const bool more = s + 1 < A;
      // ---- z_s = relu(zacc + bd) -> A operand (the MMA thread also stores it to the stash) ----
      {
        { CH_T0(); mbar_wait(B.z_full, (uint32_t)s & 1u); CH_T1E(0, 2); }
        tc_fence_after();
        float zv[NCOL];
        tmem_cols(CH_ZACC, zv);
        const float4* bd = reinterpret_cast<const float4*>(T.b_down[s] + cq * NCOL);
#pragma unroll
        for (int q = 0; q < NCOL / 4; ++q) {
          const float4 bq = __ldg(bd + q);
          zv[4 * q] = fmaxf(zv[4 * q] + bq.x, 0.f); zv[4 * q + 1] = fmaxf(zv[4 * q + 1] + bq.y, 0.f);
          zv[4 * q + 2] = fmaxf(zv[4 * q + 2] + bq.z, 0.f); zv[4 * q + 3] = fmaxf(zv[4 * q + 3] + bq.w, 0.f);
        }
        put_tile(smem + ChainSmem::kZ, zv);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.z_ready);
      }
      const float g = more ? gate_value(T.gate[s + 1]) : 0.f;
      const float omg = 1.0f - g;
      for (int c = 0; c < NC; ++c) {
        float4 bq[NCOL / 4];                   // bias chunk first: its global-load latency hides behind the ring / MMA waits
        {
          const float4* bu = reinterpret_cast<const float4*>(T.b_up[s] + c * CH_CW + cq * NCOL);
#pragma unroll
          for (int q = 0; q < NCOL / 4; ++q) bq[q] = __ldg(bu + q);
        }
        float xr[NCOL];
        read_h(xr);                            // residual x_s[c]
        const int b = n_u & 1; const uint32_t uph = (uint32_t)(n_u >> 1) & 1u;
        { CH_T0(); mbar_wait(&B.u_full[b], uph); CH_T1E(0, 3); }
        tc_fence_after();
        float lv[NCOL];
        tmem_cols((uint32_t)(CH_UACC + b * 64), lv);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.u_empty[b]);
        ++n_u;
#pragma unroll
        for (int q = 0; q < NCOL / 4; ++q) {
          lv[4 * q] += bq[q].x + xr[4 * q]; lv[4 * q + 1] += bq[q].y + xr[4 * q + 1];
          lv[4 * q + 2] += bq[q].z + xr[4 * q + 2]; lv[4 * q + 3] += bq[q].w + xr[4 * q + 3];
        }
        if (more) {
          float hv[NCOL], xv[NCOL];
          read_h(hv);
          if (is_mm) {
            float h2[NCOL];
            read_h(h2);
#pragma unroll
            for (int k = 0; k < NCOL; ++k) xv[k] = fmaf(omg, h2[k], fmaf(g, hv[k], lv[k]));
          } else {
#pragma unroll
            for (int k = 0; k < NCOL; ++k) xv[k] = fmaf(g, hv[k], omg * lv[k]);
          }
          emit(xv, true);
        } else {
          emit(lv, false);
        }
      }
    }
    if (threadIdx.x == 96) CH_TT1(0, 3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, CH_TMEM_COLS);
}

// ================================================================================================================
// backward
// ================================================================================================================
__global__ void __launch_bounds__(CH_THREADS, 1) san_chain_bwd_kernel(const __grid_constant__ ChainBwdArgs a) {
  const ChainBwdTower& T = a.tower[blockIdx.y];
  const bool is_mm = (T.mode == 1);
  const int NC = a.d / CH_CW;
  const int A = a.n_stages;
  const int m0 = blockIdx.x * CH_ROWS;
  const int NP = a.n_pad;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  ChainBars B(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&T.map_wd); tma_prefetch_desc(&T.map_wu); tma_prefetch_desc(&T.map_h); tma_prefetch_desc(&T.map_dy);
    tma_prefetch_desc(&T.map_aux); tma_prefetch_desc(&T.map_dz);
    B.init();
  }
  if (warp == 1) tmem_alloc(B.tmem_slot, CH_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;

  // unit u = (j + 1) * NC + c, j = -1 .. A-1, stage sv = A-1-j : first half Wd_sv[c] (j >= 0), second half Wu_{sv-1}[c]
  // (sv-1 >= 0 ; for j = -1 that is Wu_{A-1}[c])
  const int n_units = (A + 1) * NC;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (elect_one()) {
      for (int u = 0; u < n_units; ++u) {
        const int j = u / NC - 1, c = u % NC;
        const int sv = A - 1 - j;                         // j = -1 -> A
        const int slot = u % CH_NW; const uint32_t ph = (uint32_t)(u / CH_NW) & 1u;
        const bool has_wd = j >= 0, has_wu = sv - 1 >= 0;
        { CH_T0(); mbar_wait(&B.w_empty[slot], ph ^ 1u); CH_T1(1, 0, 0); }
        uint8_t* dst = smem + ChainSmem::kW + slot * 2 * CH_W_BYTES;
        mbar_expect_tx(&B.w_full[slot], (has_wd ? CH_W_BYTES : 0) + (has_wu ? CH_W_BYTES : 0));
        if (has_wd) tma_load_2d(dst, &T.map_wd, &B.w_full[slot], c * CH_CW, sv * CH_R);                       // Wd_sv[:, chunk] : [r, 64]
        if (has_wu) tma_load_2d(dst + CH_W_BYTES, &T.map_wu, &B.w_full[slot], 0, (sv - 1) * a.d + c * CH_CW);  // Wu_{sv-1}[chunk, :] : [64, r]
      }
    }
  } else if (warp == 2) {
    // ===================== data producer =====================
    if (elect_one()) {
      int slot = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* m, int col, int row) {
        { CH_T0(); mbar_wait(&B.h_empty[slot], ph ^ 1u); CH_T1(1, 2, 0); }
        mbar_expect_tx(&B.h_full[slot], CH_TILE_BYTES);
        tma_load_2d(smem + ChainSmem::kH + slot * CH_TILE_BYTES, m, &B.h_full[slot], col, row);
        if (++slot == CH_NH) { slot = 0; ph ^= 1u; }
      };
      for (int c = 0; c < NC; ++c) load(&T.map_dy, c * CH_CW, (A - 1) * NP + m0);
      for (int s = A - 1; s >= 0; --s) {
        for (int c = 0; c < NC; ++c) {
          load(&T.map_dy, c * CH_CW, s * NP + m0);
          load(&T.map_h, T.layer[s] * a.d + c * CH_CW, m0);
          if (is_mm) load(&T.map_aux, T.layer2[s] * a.d + c * CH_CW, m0);
          else if (s > 0) load(&T.map_aux, c * CH_CW, s * NP + m0);          // x_s (the forward's stash)
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer + TMA stores =====================
    if (elect_one()) {
      CH_TT0();
      constexpr uint32_t idesc = instr_desc_bf16(CH_ROWS, 64, 0, 1);   // A K-major, B MN-major ([K, N] row-major weight tiles)
      const uint32_t sz = smem_u32(smem + ChainSmem::kZ);
      int n_x = 0, n_u = 0;
      // dy_{sp}[c] is in xk[b]: store it (sp >= 0), feed dz_acc (+)= dy chunk x Wu_{sp}[c]
      auto dzacc_step = [&](int unit, int c, int sp, bool store) {
        const int b = n_x & 1; const uint32_t ph = (uint32_t)(n_x >> 1) & 1u;
        const int slot = unit % CH_NW;
        { CH_T0(); mbar_wait(&B.xk_full[b], ph); CH_T1(1, 1, 0); }
        tc_fence_after();
        const uint8_t* xk = smem + ChainSmem::kXk + b * CH_TILE_BYTES;
        if (store) { tma_store_2d(&T.map_dy, xk, c * CH_CW, sp * NP + m0); bulk_commit(); }
        const uint32_t sx = smem_u32(xk);
        const uint32_t sw = smem_u32(smem + ChainSmem::kW + slot * 2 * CH_W_BYTES + CH_W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ss(tmem_base + CH_ZACC, smem_desc_sw128(sx + k * 32, 16, 1024), smem_desc_sw128(sw + k * 2048, 8192, 1024), idesc, (c > 0 || k > 0) ? 1u : 0u);
        if (store) { { CH_T0(); bulk_wait_read0(); CH_T1(1, 1, 1); } { CH_T0(); bulk_wait_3(); CH_T1(1, 1, 2); } }
        { CH_T0(); mma_commit(&B.xk_empty[b]); mma_commit(&B.w_empty[slot]); CH_T1(1, 1, 3); }
        ++n_x;
      };
      for (int c = 0; c < NC; ++c) {
        const int unit = c; const uint32_t wph = (uint32_t)(unit / CH_NW) & 1u;
        { CH_T0(); mbar_wait(&B.w_full[unit % CH_NW], wph); CH_T1(1, 1, 5); }
        dzacc_step(unit, c, A - 1, false);
      }
      mma_commit(B.z_full);
      for (int j = 0; j < A; ++j) {
        const int s = A - 1 - j;
        const bool more = s > 0;
        { CH_T0(); mbar_wait(B.z_ready, (uint32_t)j & 1u); CH_T1(1, 1, 6); }
        tc_fence_after();
        tma_store_2d(&T.map_dz, smem + ChainSmem::kZ, 0, s * NP + m0);   // dz_s stash (wgrad operand)
        bulk_commit();
        for (int c = 0; c < NC; ++c) {
          const int unit = (j + 1) * NC + c; const int slot = unit % CH_NW; const uint32_t wph = (uint32_t)(unit / CH_NW) & 1u;
          const int b = n_u & 1; const uint32_t uph = (uint32_t)(n_u >> 1) & 1u;
          { CH_T0(); mbar_wait(&B.w_full[slot], wph); CH_T1(1, 1, 5); }
          { CH_T0(); mbar_wait(&B.u_empty[b], uph ^ 1u); CH_T1(1, 1, 7); }
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + ChainSmem::kW + slot * 2 * CH_W_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)       // dx chunk = dz (K = r) x Wd[:, chunk]
            mma_bf16_ss(tmem_base + CH_UACC + b * 64, smem_desc_sw128(sz + k * 32, 16, 1024), smem_desc_sw128(sw + k * 2048, 8192, 1024), idesc, k > 0 ? 1u : 0u);
          mma_commit(&B.u_full[b]);
          ++n_u;
          if (!more) mma_commit(&B.w_empty[slot]);
          else if (c >= 1) dzacc_step(unit - 1, c - 1, s - 1, true);
        }
        if (more) { dzacc_step((j + 1) * NC + NC - 1, NC - 1, s - 1, true); mma_commit(B.z_full); }
      }
      bulk_wait_0();
      CH_TT1(1, 1);
    }
  } else {
    // ===================== epilogue warps =====================
    CH_TT0();
    const int ew = warp - 3;
    const int quad = warp & 3;
    const int hf = ew >> 2;
    const int m = quad * 32 + lane;
    const int64_t row = (int64_t)m0 + m;
    const bool row_ok = row < a.n_items;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int sw_row = (m >> 3) * 1024 + (m & 7) * 128;
    int n_x = 0, n_u = 0;
    int h_slot = 0; uint32_t h_ph = 0;

    auto put_tile = [&](uint8_t* tile_base, const float* v) {
      uint8_t* tile = tile_base + sw_row;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(tile + (((hf * 4 + q) ^ (m & 7)) << 4)) = pack8(v + q * 8);
    };
    auto read_tile = [&](float* hv) {                 // this thread's 32 columns of the next ring tile (zeros for rows past N)
      { CH_T0(); mbar_wait(&B.h_full[h_slot], h_ph); CH_T1E(1, 1); }
      const uint8_t* tile = smem + ChainSmem::kH + h_slot * CH_TILE_BYTES + sw_row;
#pragma unroll
      for (int q = 0; q < 4; ++q) unpack8(*reinterpret_cast<const uint4*>(tile + (((hf * 4 + q) ^ (m & 7)) << 4)), hv + q * 8);
      if (!row_ok) {
#pragma unroll
        for (int k = 0; k < 32; ++k) hv[k] = 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.h_empty[h_slot]);
      if (++h_slot == CH_NH) { h_slot = 0; h_ph ^= 1u; }
    };
    auto emit = [&](const float* xv) {                // bf16 chunk -> A operand (the MMA thread stores it to the stash)
      const int b = n_x & 1; const uint32_t ph = (uint32_t)(n_x >> 1) & 1u;
      { CH_T0(); mbar_wait(&B.xk_empty[b], ph ^ 1u); CH_T1E(1, 0); }
      put_tile(smem + ChainSmem::kXk + b * CH_TILE_BYTES, xv);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.xk_full[b]);
      ++n_x;
    };

    // ---- stage "A": feed dy_{A-1} into the first dz accumulation ----
    for (int c = 0; c < NC; ++c) {
      float dv[32];
      read_tile(dv);
      emit(dv);
    }
    for (int s = A - 1; s >= 0; --s) {
      const int j = A - 1 - s;
      const bool more = s > 0;
      // ---- dz_s = dz_acc * (z_s > 0) ----
      {
        { CH_T0(); mbar_wait(B.z_full, (uint32_t)j & 1u); CH_T1E(1, 2); }
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(CH_ZACC + hf * 32), raw);
        tmem_ld_wait();
        float zv[32], dzv[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok) v = *reinterpret_cast<const uint4*>(T.z_stash + ((int64_t)s * NP + row) * CH_R + hf * 32 + q * 8);
          unpack8(v, zv + q * 8);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) dzv[k] = (zv[k] > 0.f && row_ok) ? __uint_as_float(raw[k]) : 0.f;
        put_tile(smem + ChainSmem::kZ, dzv);
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.z_ready);
        const float cs = warp_colsum32(dzv, lane);                       // db_down
        atomicAdd(T.g_b_down[s] + hf * 32 + lane, cs);
      }
      const float g = gate_value(T.gate[s]);
      const float omg = 1.0f - g;
      float gpart = 0.f;
      for (int c = 0; c < NC; ++c) {
        float dy[32], hv[32], av[32];
        read_tile(dy);
        read_tile(hv);
        const bool has_aux = is_mm || more;
        if (has_aux) read_tile(av);
        const int b = n_u & 1; const uint32_t uph = (uint32_t)(n_u >> 1) & 1u;
        { CH_T0(); mbar_wait(&B.u_full[b], uph); CH_T1E(1, 3); }
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(CH_UACC + b * 64 + hf * 32), raw);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B.u_empty[b]);
        ++n_u;
        float dx[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          dx[k] = row_ok ? __uint_as_float(raw[k]) + dy[k] : 0.f;
          const float diff = has_aux ? hv[k] - av[k] : hv[k];
          gpart = fmaf(dx[k], diff, gpart);
        }
        if (more) {
          if (!is_mm) {
#pragma unroll
            for (int k = 0; k < 32; ++k) dx[k] *= omg;
          }
          emit(dx);
        }
        const float cs = warp_colsum32(dy, lane);                        // db_up
        atomicAdd(T.g_b_up[s] + c * CH_CW + hf * 32 + lane, cs);
      }
      // ---- gate gradient: d x_s / d g = h_s - last_{s-1} (mm: h_cv - h_text), d sigmoid(p/0.1)/dp = g(1-g)/0.1.  The intra-modal
      //      towers do not stash last_{s-1}: x_s = g h_s + (1-g) last_{s-1} gives h_s - last_{s-1} = (h_s - x_s) / (1-g), and the
      //      (1-g) cancels against the sigmoid derivative (no division: a saturated gate is harmless) ----
      gpart = warp_sum(gpart);
      const float gfac = (!is_mm && more) ? g / 0.1f : g * omg / 0.1f;
      if (lane == 0) atomicAdd(T.g_gate[s], gpart * gfac);
    }
    if (threadIdx.x == 96) CH_TT1(1, 3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, CH_TMEM_COLS);
}

// ================================================================================================================
// host side
// ================================================================================================================
int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_inner, int box_outer);

// Row tiles.  The first-generation kernels use 128-row tiles.  The second generation takes the tile height as a launch parameter
// (IISAN_B200_CHAIN_ROWS, e.g. 115 rows = 49 tiles per tower = one wave of 147 CTAs); the MMAs keep M = 128, surplus TMEM lanes
// idle.  The stash matrices are addressed by row (stage s starts at row s * n_pad), so both generations can read what either wrote.
int chain_tile_rows(int n_items) {
  static const int forced = [] { const char* e = getenv("IISAN_B200_CHAIN_ROWS"); return e ? atoi(e) : 0; }();      // measurement switch
  if (forced >= 16 && forced <= CH_ROWS) return forced;
  // Measured on B200 (B = 512): 147 CTAs x 115 rows run exactly as long as 132 CTAs x 128 rows -- the kernels are bound by the
  // per-chunk latency chain inside a CTA, not by the rows it owns -- so the default stays at full tiles.
  (void)n_items;
  return CH_ROWS;
}
int chain_n_pad(int n_items) {
  const int p128 = (n_items + CH_ROWS - 1) / CH_ROWS * CH_ROWS;
  const int r = chain_tile_rows(n_items);
  const int pr = (n_items + r - 1) / r * r;
  return p128 > pr ? p128 : pr;
}

static int fill_common(CUtensorMap* map_wd, CUtensorMap* map_wu, const bf16* wd_pack, const bf16* wu_pack, int n_stages, int d) {
  IISAN_TRY(make_tensor_map_bf16(map_wd, wd_pack, (int64_t)n_stages * CH_R, d, d, CH_CW, CH_R));
  IISAN_TRY(make_tensor_map_bf16(map_wu, wu_pack, (int64_t)n_stages * d, CH_R, CH_R, CH_R, CH_CW));
  return IISAN_OK;
}

int chain_fill_tower(ChainTower* T, int mode, const void* h, int64_t n_items, int64_t h_pitch_cols, const void* h2, int64_t h2_pitch_cols,
                     const bf16* wd_pack, const bf16* wu_pack, const bf16* x_all, const bf16* last_all, const bf16* z_all, int n_stages,
                     int d, int box_rows) {
  const int CH_ROWS = box_rows;          // rows of the TMA boxes (= the row tile of the generation that will run)
  const int64_t np = chain_n_pad((int)n_items);
  T->mode = mode;
  IISAN_TRY(make_tensor_map_bf16(&T->map_h, h, n_items, h_pitch_cols, h_pitch_cols, CH_CW, CH_ROWS));
  if (mode == 1) IISAN_TRY(make_tensor_map_bf16(&T->map_h2, h2, n_items, h2_pitch_cols, h2_pitch_cols, CH_CW, CH_ROWS));
  else T->map_h2 = T->map_h;
  IISAN_TRY(make_tensor_map_bf16(&T->map_x, x_all, (int64_t)n_stages * np, d, d, CH_CW, CH_ROWS));
  IISAN_TRY(make_tensor_map_bf16(&T->map_last, last_all, (int64_t)n_stages * np, d, d, CH_CW, CH_ROWS));
  IISAN_TRY(make_tensor_map_bf16(&T->map_z, z_all, (int64_t)n_stages * np, CH_R, CH_R, CH_R, CH_ROWS));
  T->z_out = const_cast<bf16*>(z_all);
  return fill_common(&T->map_wd, &T->map_wu, wd_pack, wu_pack, n_stages, d);
}

int chain_fill_bwd_tower(ChainBwdTower* T, int mode, const void* h, int64_t n_items, int64_t h_pitch_cols, const void* h2,
                         int64_t h2_pitch_cols, const bf16* wd_pack, const bf16* wu_pack, const bf16* dy_all, const bf16* x_all,
                         const bf16* dz_all, int n_stages, int d, int box_rows) {
  const int CH_ROWS = box_rows;
  const int64_t np = chain_n_pad((int)n_items);
  T->mode = mode;
  IISAN_TRY(make_tensor_map_bf16(&T->map_h, h, n_items, h_pitch_cols, h_pitch_cols, CH_CW, CH_ROWS));
  IISAN_TRY(make_tensor_map_bf16(&T->map_dy, dy_all, (int64_t)n_stages * np, d, d, CH_CW, CH_ROWS));
  if (mode == 1) IISAN_TRY(make_tensor_map_bf16(&T->map_aux, h2, n_items, h2_pitch_cols, h2_pitch_cols, CH_CW, CH_ROWS));
  else IISAN_TRY(make_tensor_map_bf16(&T->map_aux, x_all, (int64_t)n_stages * np, d, d, CH_CW, CH_ROWS));
  IISAN_TRY(make_tensor_map_bf16(&T->map_dz, dz_all, (int64_t)n_stages * np, CH_R, CH_R, CH_R, CH_ROWS));
  T->dz_out = const_cast<bf16*>(dz_all);
  return fill_common(&T->map_wd, &T->map_wu, wd_pack, wu_pack, n_stages, d);
}

int launch_san_chain_bwd(const ChainBwdArgs& args, int n_towers, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(san_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmem::kTotal));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  const int tiles = (args.n_items + CH_ROWS - 1) / CH_ROWS;
  { LaunchScope ls_(IISAN_K_CHAIN_BWD, st); san_chain_bwd_kernel<<<dim3(tiles, n_towers), CH_THREADS, ChainSmem::kTotal, st>>>(args); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int launch_san_chain_fwd(const ChainArgs& args, int n_towers, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(san_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmem::kTotal));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  const int tiles = (args.n_items + CH_ROWS - 1) / CH_ROWS;
  { LaunchScope ls_(IISAN_K_CHAIN, st); san_chain_fwd_kernel<<<dim3(tiles, n_towers), CF_THREADS, ChainSmem::kTotal, st>>>(args); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan

#ifdef IISAN_CHAIN_TRACE
// trace build only (not declared in include/iisan_b200.h): copy the wait-time counters to the host, optionally clearing them.
// layout: [fwd|bwd][tower text|img|mm][role][site]{cycles, count}, site 15 = lifetime of the role's thread
extern "C" int iisan_debug_chain_trace_read(unsigned long long* host_out, int reset) {
  using namespace iisan;
  if (!host_out) return IISAN_EINVAL;
  IISAN_CUDA_OK(cudaDeviceSynchronize());
  IISAN_CUDA_OK(cudaMemcpyFromSymbol(host_out, g_chain_trace, sizeof(g_chain_trace)));
  if (reset) {
    static unsigned long long zeros[sizeof(g_chain_trace) / sizeof(unsigned long long)];
    IISAN_CUDA_OK(cudaMemcpyToSymbol(g_chain_trace, zeros, sizeof(g_chain_trace)));
  }
  return IISAN_OK;
}
#endif
