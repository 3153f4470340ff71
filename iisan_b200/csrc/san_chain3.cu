// Fused side-adapter chain, third generation (forward).  One CTA = one 128-row tile of ONE tower, all A stages and the merged
// head on chip.  Same arithmetic as the second generation (san_chain2.cu) up to the head, which is applied as ONE [E, d]
// matrix; what changed is where the running state lives:
//
//   * x_s (128 x d bf16) is RESIDENT on the SM for the whole kernel: chunks 0..8 as packed bf16 in tensor memory (288 columns),
//     chunks 9.. in shared memory (128-byte swizzled tiles).  An epilogue thread owns the same cells in every stage and
//     overwrites x_s[c] with x_{s+1}[c], which is the A operand of the next down-projection (tcgen05.mma with A in TMEM, or an
//     SS-mode MMA for the shared-memory chunks) AND of the residual: the up-projection accumulator is started as x_s[c] I
//     (an MMA with a 64 x 64 identity tile: bf16 values times 1.0, exact), so the epilogue reads  x_s + relu(z_s) Wu_s^T  in ONE
//     tcgen05.ld and never unpacks the residual.  No stash store, no residual re-read, no store warp, no x-slot hand-back;
//   * the up-projection bias arrives pre-multiplied by the gate factor of the next fusion ((1 - g_{s+1}) b_up, prepared with the
//     weight casts), so a column costs the epilogue two fused multiply-adds, one bf16 unpack and one pack;
//   * the only global traffic of width d is the hidden-state stream itself (the algorithmic bytes); the backward needs
//     relu(z_s) [N, 64] per stage and nothing else (san_lr.cu);
//   * stage A is the merged head: y = last_{A-1} M^T + c with M = W_pre W_fc, one more down-projection of the same loop.
//
// forward                                                          reference
//   x_0       = fuse(h_0, 0)                                       gated fusion   CC/model/model.py:319-326 (mm: :335-337)
//   z_s       = relu(x_s Wd_s^T + bd_s)                            AdapterBlock   CC/model/modules.py:113-116
//   last_s    = z_s Wu_s^T + bu_s + x_s
//   x_{s+1}   = fuse(h_{s+1}, last_s)                              (mm tower: last_s + g h_cv + (1-g) h_text)
//   y         = pre_fc(fc(last_{A-1}))                             CC/model/model.py:340-347
//
// warp roles: 0 weight TMA producer | 1 TMEM allocator + MMA issuer (down-projections) | 2 hidden-state TMA producer |
//             3 / 20 MMA issuers (up-projections of the even / odd chunks) | 4..19 epilogue
#include "san_chain3.cuh"

#include "san_chain2.cuh"

namespace iisan {

using bf16 = __nv_bfloat16;

namespace c3 {
// Ring sizes per tower kind (same 160 KB in total): an intra-modal tower reads ONE hidden-state tile per chunk and is bound by
// its weight stream, the inter-modal tower reads two and starves on a shallow hidden-state ring (measured: 600 cycles per chunk
// at 3 tiles per parity).
#ifndef C3_NW_INTRA
#define C3_NW_INTRA 8
#endif
#ifndef C3_NDR_INTRA
#define C3_NDR_INTRA 3
#endif
#ifndef C3_NW_MM
#define C3_NW_MM 4
#endif
#ifndef C3_NDR_MM
#define C3_NDR_MM 4
#endif
#ifndef C3_NT
#define C3_NT 9
#endif
constexpr int NW3_MAX = C3_NW_INTRA > C3_NW_MM ? C3_NW_INTRA : C3_NW_MM;         // weight ring (8 KB units in consumption order)
constexpr int NDR3_MAX = C3_NDR_INTRA > C3_NDR_MM ? C3_NDR_INTRA : C3_NDR_MM;   // hidden-state ring depth PER CHUNK PARITY (16 KB tiles)
constexpr int RING3_BYTES = (C3_NW_INTRA * W_BYTES + 2 * C3_NDR_INTRA * TILE_BYTES) > (C3_NW_MM * W_BYTES + 2 * C3_NDR_MM * TILE_BYTES)
                                ? (C3_NW_INTRA * W_BYTES + 2 * C3_NDR_INTRA * TILE_BYTES) : (C3_NW_MM * W_BYTES + 2 * C3_NDR_MM * TILE_BYTES);
constexpr int NT3 = C3_NT;                  // x chunks resident in tensor memory
constexpr int NS3 = 12 - C3_NT;             // x chunks resident in shared memory
constexpr int NU3 = 2;                      // one U accumulator per chunk parity
constexpr int LOOK3 = 2;                    // U(i) is issued before the down-projection of chunk i - 2
constexpr int MAX_D3 = (NT3 + NS3) * CW;    // 768
constexpr int BIAS3 = (MAX_D3 + R) * 4;     // one stage: b_up [d] | b_down [64]
constexpr int T3_ZACC = 0, T3_ZOP = 64, T3_UACC = 96, T3_X = T3_UACC + NU3 * 64;
static_assert(T3_X + NT3 * 32 <= 512, "tensor memory budget");

struct Smem3 {
  static constexpr int kW = 0;                                  // weight ring, then the hidden-state ring [parity][NDR] (runtime split)
  static constexpr int kXs = kW + RING3_BYTES;                  // resident x chunks NT3..
  static constexpr int kIdent = kXs + NS3 * TILE_BYTES;         // [64 x 64] bf16 identity (K-major, swizzled): B operand of the residual MMAs
  static constexpr int kBias = kIdent + W_BYTES;                // two stages
  static constexpr int kBar = kBias + 2 * BIAS3 + 256;          // (a gap between the bulk-copy destination and the barriers: see lr_rank_kernel)
  static constexpr int kTotal = kBar + 1024 + 1024;             // barriers + alignment slack
  static constexpr int bWFull = 0, bWEmpty = bWFull + 8 * NW3_MAX, bDFull = bWEmpty + 8 * NW3_MAX, bDEmpty = bDFull + 16 * NDR3_MAX;
  static constexpr int bXFull = bDEmpty + 16 * NDR3_MAX, bUFull = bXFull + 8 * (NT3 + NS3), bUEmpty = bUFull + 8 * NU3;
  static constexpr int bZFull = bUEmpty + 8 * NU3, bZReady = bZFull + 8, bBias = bZReady + 8, bTmem = bBias + 16;
  static constexpr int bGates = bTmem + 8;                      // kChainMaxStages floats
};
static_assert(Smem3::bGates + 4 * kChainMaxStages <= 1024, "barrier block");
static_assert(Smem3::kTotal <= 232448, "shared memory budget");
}  // namespace c3
using namespace c3;

// optional wait-time accounting (build variant "trace"; scripts/chain3_trace.py): lap timers per role, middle CTA of each tower
#ifdef IISAN_CHAIN_TRACE
__device__ unsigned int g_c3_trace[3][8][8];      // [tower][role][site] ; site 7 = lifetime of the role
#undef C2_TRACE_BUF
#define C2_TRACE_BUF g_c3_trace
#define T3_MARK() unsigned int tr_lap = clock()
#define T3_LAP(site) do { const unsigned int now_ = clock(); tr_acc[site] += now_ - tr_lap; tr_lap = now_; } while (0)
#else
#define T3_MARK() do {} while (0)
#define T3_LAP(site) do {} while (0)
#endif

constexpr int THREADS3 = THREADS + 32;        // + warp 20: the second up-projection issuer

__global__ void __launch_bounds__(THREADS3, 1) san_chain3_fwd_kernel(const __grid_constant__ Chain3Args a) {
  const Chain3Tower& T = a.tower[blockIdx.y];
  const bool is_mm = (T.mode == 1);
  const int NW3 = is_mm ? C3_NW_MM : C3_NW_INTRA, NDR3 = is_mm ? C3_NDR_MM : C3_NDR_INTRA;
  const int kD = Smem3::kW + NW3 * W_BYTES;      // start of the hidden-state ring
  const int NC = a.d / CW;                       // even, <= 12
  const int NCh = NC >> 1;
  const int A = a.n_stages;
  const int m0 = blockIdx.x * ROWS;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = sbase + Smem3::kBar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&T.map_wd); tma_prefetch_desc(&T.map_wu); tma_prefetch_desc(&T.map_h);
    if (is_mm) tma_prefetch_desc(&T.map_h2);
    for (int i = 0; i < NW3_MAX; ++i) { mbar_init_a(bar0 + Smem3::bWFull + 8 * i, 1); mbar_init_a(bar0 + Smem3::bWEmpty + 8 * i, 1); }
    for (int i = 0; i < 2 * NDR3_MAX; ++i) { mbar_init_a(bar0 + Smem3::bDFull + 8 * i, 1); mbar_init_a(bar0 + Smem3::bDEmpty + 8 * i, 8); }   // 8 warps read a tile
    for (int i = 0; i < NT3 + NS3; ++i) mbar_init_a(bar0 + Smem3::bXFull + 8 * i, 8);                                                   // 8 warps write a chunk
    for (int i = 0; i < NU3; ++i) { mbar_init_a(bar0 + Smem3::bUFull + 8 * i, 1); mbar_init_a(bar0 + Smem3::bUEmpty + 8 * i, 8); }
    mbar_init_a(bar0 + Smem3::bZFull, 1); mbar_init_a(bar0 + Smem3::bZReady, EPI_WARPS);
    mbar_init_a(bar0 + Smem3::bBias, 1); mbar_init_a(bar0 + Smem3::bBias + 8, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar0 + Smem3::bTmem), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 3) {           // identity tile: row n = 128 bytes, 16-byte groups swizzled by (n & 7)
    for (int n = lane; n < 64; n += 32) {
      const uint32_t row = sbase + Smem3::kIdent + (uint32_t)((n >> 3) * 1024 + (n & 7) * 128);
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (gq == (n >> 3)) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;      // bf16 1.0 at column n
        sts128(row + (uint32_t)((gq ^ (n & 7)) << 4), w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async_smem();
  }
  if (warp == 2 && lane < A) {
    const float gv = gate_value(T.gate[lane]);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bar0 + Smem3::bGates + 4 * lane), "f"(gv) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(bar0 + Smem3::bTmem) : "memory");
  const int nh = is_mm ? 2 : 1;                  // hidden-state tiles per chunk

  if (warp == 0) {
    // ===================== weight + bias producer: 8 KB units in the MMA thread's consumption order =====================
    if (elect_one()) {
      int n = 0;
      TR_DECL();
      auto put = [&](bool up, int s, int c) {
        const int slot = n % NW3;
        TR(0, mbar_wait_park(bar0 + Smem3::bWEmpty + 8 * slot, ((uint32_t)(n / NW3) & 1u) ^ 1u));
        const uint32_t dst = sbase + Smem3::kW + slot * W_BYTES, bar = bar0 + Smem3::bWFull + 8 * slot;
        mbar_expect_tx_a(bar, W_BYTES);
        if (up) tma_load_2d_a(dst, &T.map_wu, bar, 0, s * a.d + c * CW);      // Wu_s rows [c*64, +64), all r : [64 x r]
        else tma_load_2d_a(dst, &T.map_wd, bar, c * CW, s * R);               // Wd_s[:, chunk] : [r x 64]   (s == A: merged head)
        ++n;
      };
      // biases of stage s -> buffer s & 1 (b_up [d] | b_down [64]); stage A: only the merged head bias
      auto put_bias = [&](int s) {
        const uint32_t dst = sbase + Smem3::kBias + (s & 1) * BIAS3, bar = bar0 + Smem3::bBias + 8 * (s & 1);
        if (s < A) {
          mbar_expect_tx_a(bar, (uint32_t)(a.d * 4 + R * 4));
          bulk_load_a(dst, T.b_up[s], (uint32_t)(a.d * 4), bar);
        } else {
          mbar_expect_tx_a(bar, (uint32_t)(R * 4));
        }
        bulk_load_a(dst + a.d * 4, T.b_down[s], R * 4, bar);
      };
      put_bias(0);
      for (int c = 0; c < NC; ++c) put(false, 0, c);
      for (int s = 0; s < A; ++s) {
        const int n_u0 = n;                       // index of the unit U(s, 0)
        bool bias_sent = false;
        for (int i = 0; i < NC + LOOK3; ++i) {
          if (i < NC) put(true, s, i);
          if (i >= LOOK3) put(false, s + 1, i - LOOK3);
          // the last put waited for the release of unit n - 1 - NW3: once that is U(s, 0) or younger the MMA thread has passed
          // z_ready(s), i.e. every epilogue warp has left stage s - 1 and its bias buffer may be overwritten
          if (!bias_sent && n - 1 - NW3 >= n_u0) { put_bias(s + 1); bias_sent = true; }
        }
        if (!bias_sent) {
          mbar_wait_park(bar0 + Smem3::bWEmpty + 8 * (n_u0 % NW3), (uint32_t)(n_u0 / NW3) & 1u);
          put_bias(s + 1);
        }
      }
      TR_FLUSH(0);
    }
  } else if (warp == 2) {
    // ===================== hidden-state producer: per chunk parity, tiles in the epilogue's consumption order =====================
    if (elect_one()) {
      int n0 = 0, n1 = 0;
      TR_DECL();
      auto load = [&](int par, const CUtensorMap* m, int col) {
        int& n = par ? n1 : n0;
        const int slot = par * NDR3 + (n % NDR3);
        TR(0, mbar_wait_park(bar0 + Smem3::bDEmpty + 8 * slot, ((uint32_t)(n / NDR3) & 1u) ^ 1u));
        const uint32_t bar = bar0 + Smem3::bDFull + 8 * slot;
        mbar_expect_tx_a(bar, (uint32_t)TILE_BYTES);
        tma_load_2d_a(sbase + kD + slot * TILE_BYTES, m, bar, col, m0);
        ++n;
      };
      for (int p = 0; p < A; ++p)
        for (int c = 0; c < NC; ++c) {
          load(c & 1, &T.map_h, T.layer[p] * a.d + c * CW);
          if (is_mm) load(c & 1, &T.map_h2, T.layer2[p] * a.d + c * CW);
        }
      TR_FLUSH(2);
    }
  } else if (warp == 1 || warp == 3 || warp == 20) {
    // ===================== MMA issuers: warp 1 the down-projections, warps 3 / 20 the up-projections of the even / odd chunks =====================
    // Three threads, because a chunk costs an issuing thread two barrier waits, its MMAs and two or three tcgen05.commit of ~100
    // cycles each even when nothing has to be waited for (measured: ~570 cycles per up-projection): one thread paces the whole
    // CTA at ~1000 cycles per chunk.  All consume the ONE
    // weight ring; unit indices in the producer's order: the NC units Wd(0, c), then per stage s the interleaving
    // i = 0 .. NC + LOOK3 - 1 : [i < NC] Wu(s, i) , [i >= LOOK3] Wd(s + 1, i - LOOK3).
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_bf16(ROWS, 64, 0, 0);   // [128 x 64] (+)= A x B^T (K-major), K = 64
      auto unit_u = [&](int s, int i) { return NC + s * 2 * NC + i + (i > LOOK3 ? i - LOOK3 : 0); };
      auto unit_d = [&](int k, int c) {          // k == 0: the prologue units ; else stage s = k - 1, loop index i = c + LOOK3
        if (k == 0) return c;
        const int i = c + LOOK3;
        return NC + (k - 1) * 2 * NC + (i < NC ? i + 1 : NC) + c;
      };
      TR_DECL();
      auto wait_w = [&](int n) -> uint32_t {
        const int slot = n % NW3;
        TR(0, mbar_wait_park(bar0 + Smem3::bWFull + 8 * slot, (uint32_t)(n / NW3) & 1u));
        return sbase + Smem3::kW + slot * W_BYTES;
      };
      auto free_w = [&](int n) { mma_commit_a(bar0 + Smem3::bWEmpty + 8 * (n % NW3)); };
      // z_acc (+)= x_k[c] Wd_k[c]^T : the chunk is resident in tensor memory (c < NT3) or in shared memory
      auto down = [&](int k, int c) {
        const int n = unit_d(k, c);
        const uint32_t sw = wait_w(n);
        TR(1, mbar_wait_park(bar0 + Smem3::bXFull + 8 * c, (uint32_t)k & 1u));
        T3_MARK();
        tc_fence_after();
        if (c < NT3) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ts(tmem_base + T3_ZACC, tmem_base + T3_X + c * 32 + kk * 8, smem_desc_sw128(sw + kk * 32, 16, 1024), idesc,
                        (c > 0 || kk > 0) ? 1u : 0u);
        } else {
          const uint32_t xs = sbase + Smem3::kXs + (c - NT3) * TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ss(tmem_base + T3_ZACC, smem_desc_sw128(xs + kk * 32, 16, 1024), smem_desc_sw128(sw + kk * 32, 16, 1024), idesc,
                        (c > 0 || kk > 0) ? 1u : 0u);
        }
        if (c == NC - 1) mma_commit_a(bar0 + Smem3::bZFull);
        free_w(n);
        T3_LAP(2);
      };
      // U accumulator of chunk i of stage s: x_s[i] I (exact) + relu(z_s) Wu_s[i]^T ; the residual part is issued ahead of z_ready
      auto up = [&](int s, int i, bool& z_seen) {
        const int n = unit_u(s, i);
        const int nu = s * NC + i, ub = nu & 1;
        TR(1, mbar_wait_park(bar0 + Smem3::bUEmpty + 8 * ub, ((uint32_t)(nu >> 1) & 1u) ^ 1u));
        if (!z_seen) mbar_wait_park(bar0 + Smem3::bXFull + 8 * i, (uint32_t)s & 1u);      // ahead of z_ready: x_s[i] itself must be final
        T3_MARK();
        tc_fence_after();
        const uint32_t ident = sbase + Smem3::kIdent;
        if (i < NT3) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ts(tmem_base + T3_UACC + ub * 64, tmem_base + T3_X + i * 32 + kk * 8, smem_desc_sw128(ident + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
        } else {
          const uint32_t xs = sbase + Smem3::kXs + (i - NT3) * TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_bf16_ss(tmem_base + T3_UACC + ub * 64, smem_desc_sw128(xs + kk * 32, 16, 1024), smem_desc_sw128(ident + kk * 32, 16, 1024), idesc, kk > 0 ? 1u : 0u);
        }
        if (!z_seen) { TR(3, mbar_wait_park(bar0 + Smem3::bZReady, (uint32_t)s & 1u)); tc_fence_after(); z_seen = true; }
        const uint32_t sw = wait_w(n);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_bf16_ts(tmem_base + T3_UACC + ub * 64, tmem_base + T3_ZOP + kk * 8, smem_desc_sw128(sw + kk * 32, 16, 1024), idesc, 1u);
        mma_commit_a(bar0 + Smem3::bUFull + 8 * ub);
        free_w(n);
        T3_LAP(2);
      };
      if (a.one_issuer) {
        // every MMA of the CTA from ONE thread, in the weight ring's order (U(i) ahead of the down-projection of chunk i - LOOK3)
        if (warp == 1) {
          for (int c = 0; c < NC; ++c) down(0, c);
          for (int s = 0; s < A; ++s) {
            bool z_seen = false;
            for (int i = 0; i < NC + LOOK3; ++i) {
              if (i < NC) up(s, i, z_seen);
              if (i >= LOOK3) down(s + 1, i - LOOK3);
            }
          }
          TR_FLUSH(1);
        }
      } else if (warp == 1) {
        for (int k = 0; k <= A; ++k)
          for (int c = 0; c < NC; ++c) down(k, c);
        TR_FLUSH(1);
      } else {
        for (int s = 0; s < A; ++s) {
          bool z_seen = false;
          for (int i = (warp == 3 ? 0 : 1); i < NC; i += 2) up(s, i, z_seen);
        }
        if (warp == 3) TR_FLUSH(3);
      }
    }
  } else if (warp >= 4 && warp < 20) {
    // ===================== epilogue warps =====================
    const int ew = warp - 4;                  // 0..15
    const int quad = warp & 3;                // TMEM lane quadrant (warp % 4)
    const int grp = ew >> 2;                  // 0..3
    const int half = grp & 1;                 // which 32 columns of the chunk
    const int par = grp >> 1;                 // chunks c == par (mod 2)
    const int m = quad * 32 + lane;           // row inside the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw_row = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);      // row m inside a swizzled [128 x 64] bf16 tile
    uint32_t offq[4];                         // this thread's four 16-byte groups of a tile
#pragma unroll
    for (int q = 0; q < 4; ++q) offq[q] = sw_row + (uint32_t)(((half * 4 + q) ^ (m & 7)) << 4);
    const int64_t grow = (int64_t)m0 + m;
    const uint32_t bar_d_full = bar0 + Smem3::bDFull + par * NDR3 * 8, bar_d_empty = bar0 + Smem3::bDEmpty + par * NDR3 * 8;
    const uint32_t d_base = sbase + kD + par * NDR3 * TILE_BYTES;
    TR_DECL();

    auto d_slot = [&](int t) -> int { return is_mm ? t % C3_NDR_MM : t % C3_NDR_INTRA; };
    auto d_phase = [&](int t) -> uint32_t { return (uint32_t)(is_mm ? t / C3_NDR_MM : t / C3_NDR_INTRA) & 1u; };
    auto d_tile = [&](int t) -> uint32_t { return d_base + (uint32_t)d_slot(t) * TILE_BYTES; };
    auto d_wait = [&](int t) { mbar_wait_a(bar_d_full + d_slot(t) * 8, d_phase(t)); };
    auto d_release = [&](int t) { if (lane == 0) mbar_arrive_a(bar_d_empty + d_slot(t) * 8); };
    // overwrite this thread's 32 columns of the resident chunk c with the next state, then publish the chunk
    auto x_write = [&](int c, const uint32_t (&o)[16]) {
      if (c < NT3) {
        tmem_st_32x16(tmem_base + lane_addr + (uint32_t)(T3_X + c * 32 + half * 16), o);
        tmem_st_wait();
        tc_fence_before();
      } else {
        const uint32_t xs = sbase + Smem3::kXs + (uint32_t)(c - NT3) * TILE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) sts128(xs + offq[q], o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + Smem3::bXFull + 8 * c);
    };

    // ---- x_0 = fuse(h_0, 0) ----
    auto x0_phase = [&](auto mm_tag) {
      constexpr bool MM = decltype(mm_tag)::value;
      const float g = lds32f(bar0 + Smem3::bGates);
      const uint64_t g2 = f2pack(g, g), omg2 = f2pack(1.0f - g, 1.0f - g);
      for (int c = par; c < NC; c += 2) {
        const int t0 = (c >> 1) * (MM ? 2 : 1);
        uint32_t o[16];
        d_wait(t0);
        if (MM) d_wait(t0 + 1);
        const uint32_t tb0 = d_tile(t0), tb1 = d_tile(t0 + 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 hq = lds128(tb0 + offq[q]);
          const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
          if (MM) {
            const uint4 h2q = lds128(tb1 + offq[q]);
            const uint32_t h2w[4] = {h2q.x, h2q.y, h2q.z, h2q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) o[4 * q + k] = pack2x(ffma2(g2, bf2(hw[k]), fmul2(omg2, bf2(h2w[k]))));
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[4 * q + k] = pack2x(fmul2(g2, bf2(hw[k])));
          }
        }
        __syncwarp();
        d_release(t0);
        if (MM) d_release(t0 + 1);
        x_write(c, o);
      }
    };
    if (is_mm) x0_phase(BoolTag<true>{}); else x0_phase(BoolTag<false>{});

    // ---- the chunks of one stage: x_{s+1}[c] from U_s[c], the resident x_s[c] and h_{s+1}[c] ----
    auto stage_chunks = [&](auto mm_tag, auto more_tag, int s) {
      constexpr bool MM = decltype(mm_tag)::value, MORE = decltype(more_tag)::value;
      float g = 0.f;
      if (MORE) g = lds32f(bar0 + Smem3::bGates + 4 * (s + 1));
      const uint64_t g2 = f2pack(g, g), omg2 = f2pack(1.0f - g, 1.0f - g);
      const uint32_t bias = sbase + Smem3::kBias + (s & 1) * BIAS3 + half * 128;
      for (int c = par; c < NC; c += 2) {
        const int t0 = ((s + 1) * NCh + (c >> 1)) * (MM ? 2 : 1);       // hidden states of stage s + 1 (MORE only)
        const int nu = s * NC + c;
        T3_MARK();
        mbar_wait_a(bar0 + Smem3::bUFull + 8 * par, (uint32_t)(nu >> 1) & 1u);
        T3_LAP(0);
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + lane_addr + (uint32_t)(T3_UACC + par * 64 + half * 32), raw);      // x_s[c] + relu(z_s) Wu_s[c]^T
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar0 + Smem3::bUEmpty + 8 * par);
        T3_LAP(1);
        const uint32_t bc = bias + c * (CW * 4);        // intra-modal, MORE: (1 - g_{s+1}) b_up ; else b_up  (prepared by lr_prep_fwd_kernel)
        T3_LAP(2);
        if (MORE) { d_wait(t0); if (MM) d_wait(t0 + 1); }
        T3_LAP(3);
        const uint32_t tb1 = d_tile(t0), tb2 = d_tile(t0 + 1);
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 b0 = lds128(bc + q * 32), b1 = lds128(bc + q * 32 + 16);
          const uint64_t bb[4] = {u2pack(b0.x, b0.y), u2pack(b0.z, b0.w), u2pack(b1.x, b1.y), u2pack(b1.z, b1.w)};
          uint64_t lv[4];
          if (MORE) {
            const uint4 hq = lds128(tb1 + offq[q]);
            const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
            if (MM) {
              const uint4 h2q = lds128(tb2 + offq[q]);
              const uint32_t h2w[4] = {h2q.x, h2q.y, h2q.z, h2q.w};
#pragma unroll
              for (int k = 0; k < 4; ++k)
                lv[k] = ffma2(g2, bf2(hw[k]), ffma2(omg2, bf2(h2w[k]), fadd2(u2pack(raw[8 * q + 2 * k], raw[8 * q + 2 * k + 1]), bb[k])));
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                lv[k] = ffma2(g2, bf2(hw[k]), ffma2(omg2, u2pack(raw[8 * q + 2 * k], raw[8 * q + 2 * k + 1]), bb[k]));
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) lv[k] = fadd2(u2pack(raw[8 * q + 2 * k], raw[8 * q + 2 * k + 1]), bb[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) o[4 * q + k] = pack2x(lv[k]);
        }
        if (MORE) {
          __syncwarp();
          d_release(t0);
          if (MM) d_release(t0 + 1);
        }
        T3_LAP(4);
        x_write(c, o);
        T3_LAP(5);
      }
    };

    for (int s = 0; s <= A; ++s) {
      T3_MARK();
      // ---- z_s = relu(zacc + bd): packed bf16 into the TMEM operand of the U MMAs and into the backward's stash; stage A:
      //      y = zacc + c, fp32, straight to the output.  All 16 warps: warp (quad, grp) takes columns [grp*16, +16) ----
      mbar_wait_a(bar0 + Smem3::bBias + 8 * (s & 1), (uint32_t)(s >> 1) & 1u);      // this stage's biases have landed
      const uint32_t bd = sbase + Smem3::kBias + (s & 1) * BIAS3 + a.d * 4 + grp * 64;
      mbar_wait_a(bar0 + Smem3::bZFull, (uint32_t)s & 1u);
      tc_fence_after();
      uint32_t raw[16];
      tmem_ld_32x16(tmem_base + lane_addr + (uint32_t)(T3_ZACC + grp * 16), raw);
      tmem_ld_wait();
      if (s == A) {
        if (grow < a.n_items) {
          float4* yo = reinterpret_cast<float4*>(a.out + grow * a.out_ld + T.out_col + grp * 16);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 b = lds128(bd + q * 16);
            yo[q] = make_float4(__uint_as_float(raw[4 * q]) + __uint_as_float(b.x), __uint_as_float(raw[4 * q + 1]) + __uint_as_float(b.y),
                                __uint_as_float(raw[4 * q + 2]) + __uint_as_float(b.z), __uint_as_float(raw[4 * q + 3]) + __uint_as_float(b.w));
          }
        }
        break;
      }
      uint32_t zo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 b = lds128(bd + q * 16);
        const float z0 = fmaxf(__uint_as_float(raw[4 * q]) + __uint_as_float(b.x), 0.f), z1 = fmaxf(__uint_as_float(raw[4 * q + 1]) + __uint_as_float(b.y), 0.f);
        const float z2 = fmaxf(__uint_as_float(raw[4 * q + 2]) + __uint_as_float(b.z), 0.f), z3 = fmaxf(__uint_as_float(raw[4 * q + 3]) + __uint_as_float(b.w), 0.f);
        zo[2 * q] = pack2(z0, z1); zo[2 * q + 1] = pack2(z2, z3);
      }
      tmem_st_32x8(tmem_base + lane_addr + (uint32_t)(T3_ZOP + grp * 8), zo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + Smem3::bZReady);
      if (grow < a.n_items) {
        uint4* zs = reinterpret_cast<uint4*>(T.r_out + (grow * A + s) * R + grp * 16);
        zs[0] = make_uint4(zo[0], zo[1], zo[2], zo[3]);
        zs[1] = make_uint4(zo[4], zo[5], zo[6], zo[7]);
      }
      T3_LAP(6);
      const bool more = s + 1 < A;
      if (is_mm) { if (more) stage_chunks(BoolTag<true>{}, BoolTag<true>{}, s); else stage_chunks(BoolTag<true>{}, BoolTag<false>{}, s); }
      else { if (more) stage_chunks(BoolTag<false>{}, BoolTag<true>{}, s); else stage_chunks(BoolTag<false>{}, BoolTag<false>{}, s); }
    }
    if (quad == 0 && lane == 0) TR_FLUSH(4 + grp);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Widths: an even number of 64-column chunks, at most 12 (the resident state: 9 chunks in tensor memory, 3 in shared memory);
// the merged head is one more 64-row down-projection, i.e. E == 64.
bool chain3_shape_supported(int d, int emb) { return d % 128 == 0 && d >= 256 && d <= MAX_D3 && emb == R; }

int launch_san_chain3_fwd(const Chain3Args& args, int n_towers, cudaStream_t st) {
  if (!chain3_shape_supported(args.d, R) || args.n_stages < 1 || args.n_stages > kChainMaxStages || (args.out_ld & 3)) return IISAN_EINVAL;
  for (int t = 0; t < n_towers; ++t) {
    uintptr_t al = reinterpret_cast<uintptr_t>(args.tower[t].b_down[args.n_stages]);
    for (int s = 0; s < args.n_stages; ++s)
      al |= reinterpret_cast<uintptr_t>(args.tower[t].b_up[s]) | reinterpret_cast<uintptr_t>(args.tower[t].b_down[s]);
    if (al & 15) return IISAN_EINVAL;
  }
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(san_chain3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem3::kTotal));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  const int tiles = (args.n_items + ROWS - 1) / ROWS;
  { LaunchScope ls_(IISAN_K_CHAIN, st); san_chain3_fwd_kernel<<<dim3(tiles, n_towers), THREADS3, Smem3::kTotal, st>>>(args); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan

#ifdef IISAN_CHAIN_TRACE
// trace build only: [tower][role][site] cycle sums of the last chain3 forward launch (scripts/chain3_trace.py)
extern "C" int iisan_debug_chain3_trace_read(unsigned int* host_out) {
  using namespace iisan;
  if (!host_out) return IISAN_EINVAL;
  IISAN_CUDA_OK(cudaDeviceSynchronize());
  IISAN_CUDA_OK(cudaMemcpyFromSymbol(host_out, g_c3_trace, sizeof(g_c3_trace)));
  return IISAN_OK;
}
#endif
