// Shared pieces of the second-generation fused chain kernels (san_chain2.cu forward, san_chain2_bwd.cu backward): ring / TMEM
// constants, explicit shared-window primitives, packed f32x2 arithmetic, optional wait-time accounting.
#pragma once
#include "common.cuh"
#include "launch.cuh"
#include "san_chain.cuh"
#include "umma.cuh"

namespace iisan {

using namespace umma;

namespace c2 {
constexpr int ROWS = 128;
constexpr int CW = 64;                      // chunk width == one 128-byte swizzle atom
constexpr int R = 64;                       // adapter bottleneck
constexpr int TILE_BYTES = ROWS * CW * 2;   // 16 KB
constexpr int W_BYTES = CW * R * 2;         // 8 KB weight chunk
constexpr int NW = 6;                       // weight ring, 8 KB units in the MMA thread's consumption order
constexpr int NDR = 4;                      // data ring depth (h / h2 / residual tiles, 16 KB); ONE RING PER CHUNK PARITY
constexpr int NX = 2;                       // x slots (one per chunk parity): TMEM operand (32 columns) + staging tile of the stash store
constexpr int NU = 4;                       // U accumulators (64 TMEM columns each), two per chunk parity
// Every mbarrier below is waited on by ONE agent in strict use order, or by the two epilogue groups that own the chunks of one
// parity: x slots, U accumulators (index = chunk counter mod 2 / mod 4, NC even) and data rings are therefore per parity, so a
// waiter is never more than one phase ahead of its barrier (a parity wait cannot tell phase k from phase k-2).
constexpr int LOOK = NU - 1;                // the U MMAs run this many chunks ahead of the down-projections
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 128 + 32 * EPI_WARPS;   // 640
constexpr int MAX_D = 1024;
constexpr int BIAS_BYTES = (MAX_D + R) * 4;     // one stage: b_up [d] | b_down [64]
// tensor memory columns
constexpr int T_ZACC = 0;                   // fp32 z accumulator [128 x 64]
constexpr int T_ZOP = 64;                   // packed bf16 z operand (32 columns)
constexpr int T_UACC = 96;                  // NU x 64
constexpr int T_XOP = T_UACC + NU * 64;     // NX x 32
constexpr int T_COLS = 512;
static_assert(T_XOP + NX * 32 <= T_COLS, "tensor memory budget");

struct Smem {
  static constexpr int kW = 0;
  static constexpr int kD = kW + NW * W_BYTES;                // [parity][NDR] tiles
  static constexpr int kX = kD + 2 * NDR * TILE_BYTES;
  static constexpr int kBias = kX + NX * TILE_BYTES;          // two stages
  static constexpr int kBar = kBias + 2 * BIAS_BYTES;
  static constexpr int kTotal = kBar + 1024 + 1024;           // barriers + alignment slack
  // barrier block (byte offsets from kBar)
  static constexpr int bWFull = 0, bWEmpty = bWFull + 8 * NW, bDFull = bWEmpty + 8 * NW, bDEmpty = bDFull + 16 * NDR;
  static constexpr int bXFull = bDEmpty + 16 * NDR, bXEmpty = bXFull + 8 * NX, bUFull = bXEmpty + 8 * NX, bUEmpty = bUFull + 8 * NU;
  static constexpr int bZFull = bUEmpty + 8 * NU, bZReady = bZFull + 8, bBias = bZReady + 8, bTmem = bBias + 16, bStored = bTmem + 4;
  static constexpr int bGates = bStored + 4;                  // kChainMaxStages floats
};
static_assert(Smem::bGates + 4 * kChainMaxStages <= 1024, "barrier block");
static_assert(Smem::kTotal <= 232448, "shared memory budget");

// ---- explicit shared-window primitives (32-bit addresses) ----
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_A:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE_A;\n"
      "bra WAIT_LOOP_A;\n"
      "WAIT_DONE_A:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// single-thread roles: let the hardware park the thread (suspend-time hint) instead of spinning on the issue port that the
// epilogue warps of the same scheduler need
__device__ __forceinline__ void mbar_wait_park(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_P:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE_P;\n"
      "bra WAIT_LOOP_P;\n"
      "WAIT_DONE_P:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(20000u)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_a(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void mma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t a) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 rows x 16 bf16 per MMA, packed two per 32-bit column (8 columns per K step)
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- packed f32x2 arithmetic ----
__device__ __forceinline__ uint64_t f2pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t u2pack(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
// two bf16 (one 32-bit word, low element first) -> two fp32
__device__ __forceinline__ uint64_t bf2(uint32_t w) {
  uint64_t r;
  asm("{\n.reg .b32 lo, hi;\nshl.b32 lo, %1, 16;\nand.b32 hi, %1, 0xffff0000;\nmov.b64 %0, {lo, hi};\n}" : "=l"(r) : "r"(w));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// two fp32 (register pair) -> one word of two bf16, round to nearest even, low element first
__device__ __forceinline__ uint32_t pack2x(uint64_t v) {
  uint32_t r;
  asm("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\ncvt.rn.bf16x2.f32 %0, hi, lo;\n}" : "=r"(r) : "l"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <bool V> struct BoolTag { static constexpr bool value = V; };
}  // namespace c2

using namespace c2;

// ---- optional wait-time accounting (build variant "trace", -DIISAN_CHAIN_TRACE; scripts/chain2_trace.py).  Every role of the
// middle CTA of a tower sums, in registers, the cycles spent at each wait site and writes them once at the end. ----
#ifdef IISAN_CHAIN_TRACE
constexpr int TR_SITES = 8;
#define TR_DECL() unsigned int tr_acc[TR_SITES] = {0, 0, 0, 0, 0, 0, 0, 0}; const unsigned int tr_life0 = clock()
#define TR(site, stmt) do { const unsigned int tr_t0 = clock(); stmt; tr_acc[site] += clock() - tr_t0; } while (0)
#define TR_FLUSH(role)                                                                             \
  do {                                                                                             \
    if (blockIdx.x == gridDim.x / 2) {                                                             \
      tr_acc[TR_SITES - 1] = clock() - tr_life0;                                                   \
      for (int i = 0; i < TR_SITES; ++i) C2_TRACE_BUF[blockIdx.y][role][i] = tr_acc[i];             \
    }                                                                                              \
  } while (0)
#else
#define TR_DECL() do {} while (0)
#define TR(site, stmt) do { stmt; } while (0)
#define TR_FLUSH(role) do {} while (0)
#endif


}  // namespace iisan
