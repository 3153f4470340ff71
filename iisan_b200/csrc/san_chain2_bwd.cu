// Fused side-adapter chain, second generation (backward: data / gate / bias gradients; the weight gradients stay split-K GEMMs
// over the stashes written here).  Same contract as san_chain_bwd_kernel (san_chain.cu); the dataflow is the forward's
// (san_chain2.cu): MMA A operands in tensor memory, per-parity rings, 16 epilogue warps in 4 groups, store warp, explicit
// store-completion ordering of the stash re-read.  New here: the bias gradient of the up-projection, a column sum over the 128
// rows of every dy tile, runs ON THE TENSOR CORE (ones[128 x 128] x dy tile, the tile read MN-major from the staging buffer it
// already sits in) and is read back by one warp -- the first generation spent a 31-shuffle butterfly per 32 columns in every
// epilogue warp on it.
//
//   dz_s      = (dy_s Wu_s) * (z_s > 0)                            dy_s = d last_s          CC/model/modules.py:113-116 (backward)
//   dx_s      = dy_s + dz_s Wd_s
//   dgate_s  += sum dx_s * (h_s - last_{s-1})                      (mm: h_cv - h_text), times g(1-g)/0.1 ; the intra-modal towers
//                                                                  use (h_s - x_s) g/0.1 instead (x_s is stashed, last_{s-1} is not)
//   dy_{s-1}  = (1 - g_s) dx_s                                     (mm: dx_s)               CC/model/model.py:319-326, 335-337
//   db_up_s   = colsum(dy_s) ; db_down_s = colsum(dz_s)
//
// warp roles: 0 weight TMA producer | 1 TMEM allocator + MMA issuer | 2 data TMA producer | 3 TMA stores + column-sum readout |
//             4..19 epilogue
#include "san_chain2.cuh"

namespace iisan {

using bf16 = __nv_bfloat16;
using namespace c2;

#ifdef IISAN_CHAIN_TRACE
__device__ unsigned int g_c2b_trace[3][8][8];      // [tower][role][site] ; site 7 = lifetime of the role
#undef C2_TRACE_BUF
#define C2_TRACE_BUF g_c2b_trace
#endif

namespace c2b {
// tensor memory columns
constexpr int T_ZACC = 0;                   // fp32 dz accumulator [128 x 64]
constexpr int T_ZOP = 64;                   // packed bf16 dz operand (32 columns)
constexpr int T_UACC = 96;                  // NU x 64 : dz_s Wd_s[:, c]
constexpr int T_XOP = T_UACC + NU * 64;     // NX x 32 : dy_{s-1} chunk operands
constexpr int T_CS = T_XOP + NX * 32;       // 64 : column sums of a dy tile (every row holds the same 64 sums)
constexpr int T_ONES = T_CS + 64;           // 8 : one K step of an all-ones bf16 A operand
constexpr int T_COLS = 512;
static_assert(T_ONES + 8 <= T_COLS, "tensor memory budget");

struct Smem {
  static constexpr int kW = 0;
  static constexpr int kD = kW + NW * W_BYTES;                // [parity][NDR] tiles
  static constexpr int kX = kD + 2 * NDR * TILE_BYTES;
  static constexpr int kBar = kX + NX * TILE_BYTES;
  static constexpr int kTotal = kBar + 1024 + 1024;           // barriers + alignment slack
  static constexpr int bWFull = 0, bWEmpty = bWFull + 8 * NW, bDFull = bWEmpty + 8 * NW, bDEmpty = bDFull + 16 * NDR;
  static constexpr int bXFull = bDEmpty + 16 * NDR, bXEmpty = bXFull + 8 * NX, bUFull = bXEmpty + 8 * NX, bUEmpty = bUFull + 8 * NU;
  static constexpr int bZFull = bUEmpty + 8 * NU, bZReady = bZFull + 8, bCsFull = bZReady + 8, bCsEmpty = bCsFull + 8;
  static constexpr int bTmem = bCsEmpty + 8, bStored = bTmem + 4, bGates = bStored + 4;
};
static_assert(Smem::bGates + 4 * kChainMaxStages <= 1024, "barrier block");
static_assert(Smem::kTotal <= 232448, "shared memory budget");
}  // namespace c2b

// Tile sequence of the data ring of one chunk parity (j = c >> 1).  Phase 0: dy_{A-1}[c].  Stage number q (s = A-1-q): per chunk
// dy_s[c], h_s[c] and -- inter-modal tower: the text states; intra-modal towers with s > 0: the x_s stash -- a third tile.
struct BwdTileSeq {
  int NCh, A, mm;
  __device__ int tiles(int s) const { return (mm || s > 0) ? 3 : 2; }
  __device__ int phase0(int j) const { return j; }
  __device__ int stage(int q, int j) const { return NCh * (1 + 3 * q) + j * tiles(A - 1 - q); }    // every earlier stage has s > 0
};

__global__ void __launch_bounds__(THREADS, 1) san_chain2_bwd_kernel(const __grid_constant__ ChainBwdArgs a) {
  using S = c2b::Smem;
  const ChainBwdTower& T = a.tower[blockIdx.y];
  const bool is_mm = (T.mode == 1);
  const int NC = a.d / CW;
  const int A = a.n_stages;
  const int m0 = blockIdx.x * a.rows;            // a.rows <= 128 rows per tile (see chain_tile_rows)
  const int NP = a.n_pad;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = sbase + S::kBar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&T.map_wd); tma_prefetch_desc(&T.map_wu); tma_prefetch_desc(&T.map_h); tma_prefetch_desc(&T.map_dy);
    tma_prefetch_desc(&T.map_aux);
    for (int i = 0; i < NW; ++i) { mbar_init_a(bar0 + S::bWFull + 8 * i, 1); mbar_init_a(bar0 + S::bWEmpty + 8 * i, 1); }
    for (int i = 0; i < 2 * NDR; ++i) { mbar_init_a(bar0 + S::bDFull + 8 * i, 1); mbar_init_a(bar0 + S::bDEmpty + 8 * i, 8); }
    for (int i = 0; i < NX; ++i) { mbar_init_a(bar0 + S::bXFull + 8 * i, 8); mbar_init_a(bar0 + S::bXEmpty + 8 * i, 2); }
    for (int i = 0; i < NU; ++i) { mbar_init_a(bar0 + S::bUFull + 8 * i, 1); mbar_init_a(bar0 + S::bUEmpty + 8 * i, 8); }
    mbar_init_a(bar0 + S::bZFull, 1); mbar_init_a(bar0 + S::bZReady, EPI_WARPS);
    mbar_init_a(bar0 + S::bCsFull, 1); mbar_init_a(bar0 + S::bCsEmpty, 1);
    st_release_shared(bar0 + S::bStored, 0u);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar0 + S::bTmem), "r"((uint32_t)c2b::T_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 2 && lane < A) {
    const float gv = gate_value(T.gate[lane]);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bar0 + S::bGates + 4 * lane), "f"(gv) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bar0 + S::bTmem) : "memory");
  if (warp >= 4 && warp < 8) {               // the all-ones A operand (bf16 1.0 pairs), one K step, all 128 lanes
    uint32_t ones[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ones[i] = 0x3F803F80u;
    tmem_st_32x8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c2b::T_ONES, ones);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const BwdTileSeq TS{NC / 2, A, is_mm ? 1 : 0};
  const int n_emit = A * NC;                 // x-slot uses: phase 0 (dy_{A-1}, not stored) + the dy_{s-1} chunks of stages A-1 .. 1

  if (warp == 0) {
    // ===================== weight producer: 8 KB units in the MMA thread's consumption order =====================
    if (elect_one()) {
      int n = 0;
      TR_DECL();
      auto put = [&](bool up, int s, int c) {
        const int slot = n % NW;
        TR(0, mbar_wait_park(bar0 + S::bWEmpty + 8 * slot, ((uint32_t)(n / NW) & 1u) ^ 1u));
        const uint32_t dst = sbase + S::kW + slot * W_BYTES, bar = bar0 + S::bWFull + 8 * slot;
        mbar_expect_tx_a(bar, W_BYTES);
        if (up) tma_load_2d_a(dst, &T.map_wu, bar, 0, s * a.d + c * CW);      // Wu_s[chunk rows, :] : [64 (k = column) x r]
        else tma_load_2d_a(dst, &T.map_wd, bar, c * CW, s * R);               // Wd_s[:, chunk] : [r (k) x 64]
        ++n;
      };
      for (int c = 0; c < NC; ++c) put(true, A - 1, c);
      for (int q = 0; q < A; ++q) {
        const int s = A - 1 - q;
        const bool more = s > 0;
        for (int c = 0; c < NC + LOOK; ++c) {
          if (c < NC) put(false, s, c);
          if (c >= LOOK && more) put(true, s - 1, c - LOOK);
        }
      }
      TR_FLUSH(0);
    }
  } else if (warp == 2) {
    // ===================== data producer =====================
    if (elect_one()) {
      int n0 = 0, n1 = 0;
      TR_DECL();
      auto load = [&](int par, const CUtensorMap* m, int col, int row) {
        int& n = par ? n1 : n0;
        const int slot = par * NDR + (n & (NDR - 1));
        TR(0, mbar_wait_park(bar0 + S::bDEmpty + 8 * slot, ((uint32_t)(n / NDR) & 1u) ^ 1u));
        const uint32_t bar = bar0 + S::bDFull + 8 * slot;
        mbar_expect_tx_a(bar, (uint32_t)(a.rows * CW * 2));
        tma_load_2d_a(sbase + S::kD + slot * TILE_BYTES, m, bar, col, row);
        ++n;
      };
      // with a.pf > 0 the tiles that come from HBM (hidden states, the forward's x stash) are requested into L2 a.pf chunks
      // before their ring load: three tiles per chunk leave the ring about one chunk of look-ahead
      const int PF = a.pf;
      auto prefetch = [&](int k) {                // k-th chunk of the stage loops: stage number k / NC, chunk k % NC
        const int q = k / NC, c = k % NC;
        if (q >= A) return;
        const int s = A - 1 - q;
        tma_prefetch_l2_2d(&T.map_h, T.layer[s] * a.d + c * CW, m0);
        if (is_mm) tma_prefetch_l2_2d(&T.map_aux, T.layer2[s] * a.d + c * CW, m0);
        else if (s > 0) tma_prefetch_l2_2d(&T.map_aux, c * CW, s * NP + m0);
      };
      if (PF > 0) for (int k = 0; k < PF; ++k) prefetch(k);
      for (int c = 0; c < NC; ++c) load(c & 1, &T.map_dy, c * CW, (A - 1) * NP + m0);
      for (int q = 0; q < A; ++q) {
        const int s = A - 1 - q;
        for (int c = 0; c < NC; ++c) {
          if (PF > 0) prefetch(q * NC + c + PF);
          if (q > 0) {      // dy_s[c] was stored by this CTA as x-slot use q*NC + c: wait until that store is complete
            const uint32_t need = (uint32_t)(q * NC + c + 1);
            TR(1, while (ld_acquire_shared(bar0 + S::bStored) < need) __nanosleep(64));
          }
          load(c & 1, &T.map_dy, c * CW, s * NP + m0);
          load(c & 1, &T.map_h, T.layer[s] * a.d + c * CW, m0);
          if (is_mm) load(c & 1, &T.map_aux, T.layer2[s] * a.d + c * CW, m0);
          else if (s > 0) load(c & 1, &T.map_aux, c * CW, s * NP + m0);          // x_s (the forward's stash)
        }
      }
      TR_FLUSH(2);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_bf16(ROWS, 64, 0, 1);   // A in TMEM (K-major), B MN-major ([K, N] row-major tiles)
      int nw = 0;
      TR_DECL();
      auto wait_w = [&]() -> uint32_t {
        const int slot = nw % NW;
        TR(0, mbar_wait_park(bar0 + S::bWFull + 8 * slot, (uint32_t)(nw / NW) & 1u));
        return sbase + S::kW + slot * W_BYTES;
      };
      auto free_w = [&]() { mma_commit_a(bar0 + S::bWEmpty + 8 * (nw % NW)); ++nw; };
      // x-slot use nx holds a dy chunk: dz_acc (+)= dy chunk x Wu[c] ; column sums of the chunk (staging tile as MN-major B, K = rows)
      auto down = [&](int nx, int c, bool last_chunk) {
        const uint32_t sw = wait_w();
        const int xb = nx & 1;
        TR(1, mbar_wait_park(bar0 + S::bXFull + 8 * xb, (uint32_t)(nx >> 1) & 1u));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_bf16_ts(tmem_base + c2b::T_ZACC, tmem_base + c2b::T_XOP + xb * 32 + k * 8, smem_desc_sw128(sw + k * 2048, 8192, 1024), idesc,
                      (c > 0 || k > 0) ? 1u : 0u);
        if (last_chunk) mma_commit_a(bar0 + S::bZFull);
        free_w();
        TR(5, mbar_wait_park(bar0 + S::bCsEmpty, ((uint32_t)nx & 1u) ^ 1u));
        tc_fence_after();
        const uint32_t stg = sbase + S::kX + xb * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_bf16_ts(tmem_base + c2b::T_CS, tmem_base + c2b::T_ONES, smem_desc_sw128(stg + k * 2048, 8192, 1024), idesc, k > 0 ? 1u : 0u);
        TR(4, mma_commit_a(bar0 + S::bCsFull); mma_commit_a(bar0 + S::bXEmpty + 8 * xb));
      };
      for (int c = 0; c < NC; ++c) down(c, c, c == NC - 1);
      for (int q = 0; q < A; ++q) {
        const bool more = q + 1 < A;
        TR(2, mbar_wait_park(bar0 + S::bZReady, (uint32_t)q & 1u));
        tc_fence_after();
        for (int c = 0; c < NC + LOOK; ++c) {
          if (c < NC) {
            const uint32_t sw = wait_w();
            const int nu = q * NC + c, ub = nu & 3;
            TR(3, mbar_wait_park(bar0 + S::bUEmpty + 8 * ub, ((uint32_t)(nu >> 2) & 1u) ^ 1u));
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)       // dx chunk = dz (K = r) x Wd[:, chunk]
              mma_bf16_ts(tmem_base + c2b::T_UACC + ub * 64, tmem_base + c2b::T_ZOP + k * 8, smem_desc_sw128(sw + k * 2048, 8192, 1024), idesc,
                          k > 0 ? 1u : 0u);
            TR(4, mma_commit_a(bar0 + S::bUFull + 8 * ub); free_w());
          }
          if (c >= LOOK && more) down((q + 1) * NC + (c - LOOK), c - LOOK, c - LOOK == NC - 1);
        }
      }
      TR_FLUSH(1);
    }
  } else if (warp == 3) {
    // ===================== store warp: dy_{s-1} stash + column-sum readout (db_up) =====================
    const uint32_t lane_addr = (uint32_t)(3 * 32) << 16;
    uint32_t pub = 0;
    TR_DECL();
    for (int nx = 0; nx < n_emit; ++nx) {
      const int q = nx / NC, c = nx % NC;               // the chunk dy_{A-1-q}[c]; q == 0 came from the heads and is not stored again
      const int xb = nx & 1;
      const int sdy = A - 1 - q;
      if (lane == 0) {
        TR(0, mbar_wait_park(bar0 + S::bXFull + 8 * xb, (uint32_t)(nx >> 1) & 1u));
        if (q > 0) { tma_store_2d_a(&T.map_dy, sbase + S::kX + xb * TILE_BYTES, c * CW, sdy * NP + m0); bulk_commit(); }
      }
      __syncwarp();
      // column sums: every TMEM row holds the same 64 sums; lane l keeps columns l and l + 32
      TR(3, mbar_wait_a(bar0 + S::bCsFull, (uint32_t)nx & 1u));
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)c2b::T_CS, v0);
      tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(c2b::T_CS + 32), v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + S::bCsEmpty);
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) {               // select tree: after level w, entry i < w holds the candidate of lanes with that bit pattern
        const bool hi = (lane & w) != 0;
#pragma unroll
        for (int i = 0; i < w; ++i) { v0[i] = hi ? v0[i + w] : v0[i]; v1[i] = hi ? v1[i + w] : v1[i]; }
      }
      float* gb = T.g_b_up[sdy] + c * CW;
      atomicAdd(gb + lane, __uint_as_float(v0[0]));
      atomicAdd(gb + 32 + lane, __uint_as_float(v1[0]));
      if (lane == 0) {
        if (q > 0) { TR(1, bulk_wait_read0()); }
        mbar_arrive_a(bar0 + S::bXEmpty + 8 * xb);
        uint32_t done;
        if (q == 0) done = (uint32_t)nx + 1u;
        else { TR(2, bulk_wait<2>()); done = (uint32_t)nx - 1u; }      // every store but the two youngest is complete
        if (done > pub) { pub = done; st_release_shared(bar0 + S::bStored, pub); }
      }
    }
    if (lane == 0) {
      bulk_wait<0>();
      st_release_shared(bar0 + S::bStored, (uint32_t)n_emit);
      TR_FLUSH(3);
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 4;                  // 0..15
    const int quad = warp & 3;                // TMEM lane quadrant (warp % 4)
    const int grp = ew >> 2;                  // 0..3
    const int half = grp & 1;                 // which 32 columns of the chunk
    const int par = grp >> 1;                 // chunks c == par (mod 2)
    const int m = quad * 32 + lane;           // row inside the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw_row = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
    uint32_t offq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) offq[q] = sw_row + (uint32_t)(((half * 4 + q) ^ (m & 7)) << 4);
    const int64_t grow = (int64_t)m0 + m;
    const bool row_ok = m < a.rows && grow < a.n_items;     // TMEM lanes beyond the tile's rows, rows beyond the batch
    const uint32_t rmask = row_ok ? 0xFFFFFFFFu : 0u;       // rows past the item count: dy is whatever the workspace held
    const uint32_t bar_d_full = bar0 + S::bDFull + par * NDR * 8, bar_d_empty = bar0 + S::bDEmpty + par * NDR * 8;
    const uint32_t bar_x_full = bar0 + S::bXFull + par * 8, bar_x_empty = bar0 + S::bXEmpty + par * 8;
    const uint32_t d_base = sbase + S::kD + par * NDR * TILE_BYTES;
    const uint32_t x_tile = sbase + S::kX + par * TILE_BYTES;
    const uint32_t x_tmem = tmem_base + lane_addr + (uint32_t)(c2b::T_XOP + par * 32 + half * 16);
    TR_DECL();
    static_assert(NDR == 4 && NX == 2 && NU == 4, "ring index arithmetic below");

    auto d_tile = [&](int t) -> uint32_t { return d_base + (uint32_t)(t & 3) * TILE_BYTES; };
    auto d_wait = [&](int t) { TR(0, mbar_wait_a(bar_d_full + (t & 3) * 8, (uint32_t)(t >> 2) & 1u)); };
    auto d_release = [&](int t) { if (lane == 0) mbar_arrive_a(bar_d_empty + (t & 3) * 8); };
    auto emit = [&](int ux, const uint32_t (&o)[16]) {
      TR(1, mbar_wait_a(bar_x_empty, ((uint32_t)ux & 1u) ^ 1u));
#pragma unroll
      for (int q = 0; q < 4; ++q) sts128(x_tile + offq[q], o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      TR(4, tmem_st_32x16(x_tmem, o); tmem_st_wait(); tc_fence_before());
      TR(5, fence_proxy_async_smem(); __syncwarp());
      if (lane == 0) mbar_arrive_a(bar_x_full);
    };

    // ---- phase 0: dy_{A-1} chunks -> operand of the first dz accumulation (and staging tile: column sums) ----
    for (int c = par; c < NC; c += 2) {
      const int t0 = TS.phase0(c >> 1);
      uint32_t o[16];
      d_wait(t0);
      const uint32_t tb0 = d_tile(t0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 w = lds128(tb0 + offq[q]);
        o[4 * q] = w.x & rmask; o[4 * q + 1] = w.y & rmask; o[4 * q + 2] = w.z & rmask; o[4 * q + 3] = w.w & rmask;
      }
      __syncwarp();
      d_release(t0);
      emit(c >> 1, o);
    }

    // ---- the chunks of one stage ----
    auto stage_chunks = [&](auto mm_tag, auto more_tag, int q, int s) -> float {
      constexpr bool MM = decltype(mm_tag)::value, MORE = decltype(more_tag)::value;
      constexpr bool AUX = MM || MORE;
      const float g = lds32f(bar0 + S::bGates + 4 * s);
      const uint64_t omg2 = f2pack(1.0f - g, 1.0f - g), neg1 = f2pack(-1.0f, -1.0f);
      uint64_t gacc = f2pack(0.f, 0.f);
      const int ux0 = (q + 1) * (NC >> 1);
      for (int c = par; c < NC; c += 2) {
        const int t0 = TS.stage(q, c >> 1);       // dy_s[c] ; t0 + 1: h_s[c] ; t0 + 2: aux
        const int nu = q * NC + c, ub = nu & 3;
        TR(3, mbar_wait_a(bar0 + S::bUFull + 8 * ub, (uint32_t)(nu >> 2) & 1u));
        tc_fence_after();
        uint32_t raw[32];
        TR(6, tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(c2b::T_UACC + ub * 64 + half * 32), raw); tmem_ld_wait());
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar0 + S::bUEmpty + 8 * ub);
        d_wait(t0); d_wait(t0 + 1);
        if (AUX) d_wait(t0 + 2);
        const uint32_t tb0 = d_tile(t0), tb1 = d_tile(t0 + 1), tb2 = d_tile(t0 + 2);
        uint32_t o[16];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const uint4 dq = lds128(tb0 + offq[qq]);
          const uint4 hq = lds128(tb1 + offq[qq]);
          const uint32_t dw[4] = {dq.x & rmask, dq.y & rmask, dq.z & rmask, dq.w & rmask};
          const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
          uint64_t dx[4], df[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            dx[k] = fadd2(u2pack(raw[8 * qq + 2 * k], raw[8 * qq + 2 * k + 1]), bf2(dw[k]));
            df[k] = bf2(hw[k]);
          }
          if (AUX) {
            const uint4 aq = lds128(tb2 + offq[qq]);
            const uint32_t aw[4] = {aq.x, aq.y, aq.z, aq.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) df[k] = ffma2(bf2(aw[k]), neg1, df[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) gacc = ffma2(dx[k], df[k], gacc);
          if (MORE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[4 * qq + k] = pack2x(MM ? dx[k] : fmul2(omg2, dx[k]));
          }
        }
        __syncwarp();
        d_release(t0); d_release(t0 + 1);
        if (AUX) d_release(t0 + 2);
        if (MORE) emit(ux0 + (c >> 1), o);
      }
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(gacc));
      return row_ok ? lo + hi : 0.f;
    };

    for (int q = 0; q < A; ++q) {
      const int s = A - 1 - q;
      const bool more = s > 0;
      // ---- dz_s = dz_acc * (z_s > 0): packed bf16 into the TMEM operand of the dx MMAs, to the stash (weight-gradient operand),
      //      db_down.  Warp (quad, grp) takes columns [grp*16, +16) of its 32 rows ----
      {
        const uint4* zs = reinterpret_cast<const uint4*>(T.z_stash + ((int64_t)s * NP + (row_ok ? grow : (int64_t)m0)) * R + grp * 16);
        const uint4 z0 = __ldg(zs), z1 = __ldg(zs + 1);              // issued before the wait
        TR(2, mbar_wait_a(bar0 + S::bZFull, (uint32_t)q & 1u));
        tc_fence_after();
        uint32_t raw[16];
        tmem_ld_32x16(tmem_base + lane_addr + (uint32_t)(c2b::T_ZACC + grp * 16), raw);
        tmem_ld_wait();
        const uint32_t zw[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float dzv[16];
        uint32_t zo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const bool p0 = row_ok && (zw[k] & 0x7FFFu) != 0u && (zw[k] & 0x8000u) == 0u;              // low element > 0
          const bool p1 = row_ok && (zw[k] & 0x7FFF0000u) != 0u && (zw[k] & 0x80000000u) == 0u;      // high element > 0
          dzv[2 * k] = p0 ? __uint_as_float(raw[2 * k]) : 0.f;
          dzv[2 * k + 1] = p1 ? __uint_as_float(raw[2 * k + 1]) : 0.f;
          zo[k] = pack2(dzv[2 * k], dzv[2 * k + 1]);
        }
        tmem_st_32x8(tmem_base + lane_addr + (uint32_t)(c2b::T_ZOP + grp * 8), zo);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar0 + S::bZReady);
        if (m < a.rows) {
          uint4* dzs = reinterpret_cast<uint4*>(T.dz_out + ((int64_t)s * NP + grow) * R + grp * 16);
          dzs[0] = make_uint4(zo[0], zo[1], zo[2], zo[3]);
          dzs[1] = make_uint4(zo[4], zo[5], zo[6], zo[7]);
        }
        // db_down: sum over the 32 rows of this warp for each of its 16 columns -> lanes 0..15 (15 + 16 shuffles)
#pragma unroll
        for (int k = 0; k < 16; ++k) dzv[k] += __shfl_xor_sync(0xffffffffu, dzv[k], 16);
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int k = 0; k < off; ++k) {
            const float send = upper ? dzv[k] : dzv[k + off];
            const float recv = __shfl_xor_sync(0xffffffffu, send, off);
            dzv[k] = (upper ? dzv[k + off] : dzv[k]) + recv;
          }
        }
        if (lane < 16) atomicAdd(T.g_b_down[s] + grp * 16 + lane, dzv[0]);
      }
      float gpart;
      if (is_mm) gpart = more ? stage_chunks(BoolTag<true>{}, BoolTag<true>{}, q, s) : stage_chunks(BoolTag<true>{}, BoolTag<false>{}, q, s);
      else gpart = more ? stage_chunks(BoolTag<false>{}, BoolTag<true>{}, q, s) : stage_chunks(BoolTag<false>{}, BoolTag<false>{}, q, s);
      // ---- gate gradient: d x_s / d g = h_s - last_{s-1} (mm: h_cv - h_text), d sigmoid(p/0.1)/dp = g(1-g)/0.1.  The intra-modal
      //      towers do not stash last_{s-1}: h_s - last_{s-1} = (h_s - x_s) / (1-g), and the (1-g) cancels ----
      gpart = warp_sum(gpart);
      const float g = lds32f(bar0 + S::bGates + 4 * s);
      const float gfac = (!is_mm && more) ? g / 0.1f : g * (1.0f - g) / 0.1f;
      if (lane == 0) atomicAdd(T.g_gate[s], gpart * gfac);
    }
    if (quad == 0 && lane == 0) TR_FLUSH(4 + grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, c2b::T_COLS);
}

int launch_san_chain2_bwd(const ChainBwdArgs& args, int n_towers, cudaStream_t st) {
  if (!chain2_shape_supported(args.d) || args.rows < 1 || args.rows > ROWS) return IISAN_EINVAL;
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(san_chain2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c2b::Smem::kTotal));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  const int tiles = (args.n_items + args.rows - 1) / args.rows;
  { LaunchScope ls_(IISAN_K_CHAIN_BWD, st); san_chain2_bwd_kernel<<<dim3(tiles, n_towers), THREADS, c2b::Smem::kTotal, st>>>(args); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan

#ifdef IISAN_CHAIN_TRACE
extern "C" int iisan_debug_chain2_bwd_trace_read(unsigned int* host_out) {
  using namespace iisan;
  if (!host_out) return IISAN_EINVAL;
  IISAN_CUDA_OK(cudaDeviceSynchronize());
  IISAN_CUDA_OK(cudaMemcpyFromSymbol(host_out, g_c2b_trace, sizeof(g_c2b_trace)));
  return IISAN_OK;
}
#endif
