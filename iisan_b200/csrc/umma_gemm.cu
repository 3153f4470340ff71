// tcgen05 / TMEM / TMA GEMM (see umma_gemm.cuh).  Persistent: one CTA per SM walks the 128 x BN output tiles of all problems
// of the launch (tile = blockIdx.x + i * gridDim.x; column tiles fastest, so CTAs running side by side share their A rows in
// L2), with TWO accumulators in TMEM: the epilogue of tile i overlaps the TMA / MMA main loop of tile i + 1.
//   warp 0   : TMA producer   (one elected lane streams A/B k-blocks through a 4-stage smem ring, across tile boundaries)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, commits to mbarriers)
//   warps 2-9: epilogue       (tcgen05.ld 32 lanes x 32 columns -> bias / ReLU / mask / residual -> global)
#include "umma_gemm.cuh"

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "launch.cuh"
#include "umma.cuh"

namespace iisan {

using namespace umma;

constexpr int UBM = 128;          // rows per tile (UMMA M)
constexpr int UBK = 64;           // k-block: one 128-byte swizzle atom of bf16
constexpr int USTAGES = 4;        // at most; short reductions use fewer so that two CTAs share an SM
constexpr int UTHREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int UEPI_WARPS = 8;
constexpr int USTG_FLOATS = 32 * 33;   // per-warp transpose staging (padded: conflict-free both ways)

struct UmmaDevProblem {
  CUtensorMap map_a, map_b;   // MC: map_b boxes cover HALF of the column tile (each CTA of the pair loads one half for both)
  int M, N, K;
  int kblocks_per_split;
  int tile_start;             // first tile of this problem in the launch-wide tile list (tiles: split-major, then m, then n)
  UmmaEpilogue epi;
};
template <int NP>
struct UmmaDevBatchT {
  UmmaDevProblem p[NP];
  int stages;
  int n_probs, total_tiles;
};

struct UmmaTile { int prob, m0, n0, kb_beg, kb_end, split; };
// MC (multicast cluster of two CTAs): the list holds PAIRS of row tiles that share a column tile; CTA `crank` of the cluster
// takes row tile 2 * pair + crank (a pair past the last row tile computes on zero-filled rows and stores nothing)
template <int BN, int NP, bool MC>
__device__ __forceinline__ UmmaTile umma_decode_tile(const UmmaDevBatchT<NP>& batch, int tile, int crank) {
  int p = 0;
  while (p + 1 < batch.n_probs && tile >= batch.p[p + 1].tile_start) ++p;
  const UmmaDevProblem& P = batch.p[p];
  const int tiles_n = (P.N + BN - 1) / BN;
  const int tiles_m = MC ? ((P.M + UBM - 1) / UBM + 1) / 2 : (P.M + UBM - 1) / UBM;
  const int l = tile - P.tile_start;
  const int split = l / (tiles_m * tiles_n), r = l % (tiles_m * tiles_n);
  UmmaTile t;
  t.prob = p; t.split = split; t.m0 = ((r / tiles_n) * (MC ? 2 : 1) + crank) * UBM; t.n0 = (r % tiles_n) * BN;
  const int kb_total = (P.K + UBK - 1) / UBK;
  t.kb_beg = split * P.kblocks_per_split;
  t.kb_end = min(kb_total, t.kb_beg + P.kblocks_per_split);
  return t;
}

template <int BN>
struct UmmaSmem {
  static constexpr int kABytes = UBM * UBK * 2;
  static constexpr int kBBytes = BN * UBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = UEPI_WARPS * USTG_FLOATS * 4;
  static int total(int stages) { return stages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes; }
};


// One 32x32 output block of the epilogue, transposed domain (lane = column).  The feature set is a template so that the
// per-element code carries no pointer checks; rows advance by pointer increments.
template <bool MASK, bool RESB, bool RESF, bool OUTF, bool OUTB, bool ATOMIC>
__device__ __forceinline__ void epi_block(const UmmaEpilogue& E, const float* __restrict__ stg, int lane, int64_t row_base, int nrow,
                                          int col, float bias_v, float lo) {
  const __nv_bfloat16* mp = MASK ? E.mask + row_base * E.ld_mask + col : nullptr;
  const __nv_bfloat16* rb = RESB ? E.resid_bf16 + row_base * E.ld_resid_bf16 + col : nullptr;
  const float* rf = RESF ? E.resid_f32 + row_base * E.ld_resid_f32 + col : nullptr;
  float* of = OUTF ? E.out_f32 + row_base * E.ld_f32 + col : nullptr;
  __nv_bfloat16* ob = OUTB ? E.out_bf16 + row_base * E.ld_bf16 + col : nullptr;
#pragma unroll 1
  for (int r0 = 0; r0 < nrow; r0 += 8) {
    float add[8]; bool keep[8];
    // all global loads of the group first: the output pointers may alias the inputs as far as the compiler knows
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      add[i] = 0.f; keep[i] = true;
      if (r0 + i < nrow) {
        if (MASK) keep[i] = __bfloat162float(mp[(int64_t)i * E.ld_mask]) > 0.f;
        if (RESB) add[i] += __bfloat162float(rb[(int64_t)i * E.ld_resid_bf16]);
        if (RESF) add[i] += rf[(int64_t)i * E.ld_resid_f32];
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (r0 + i < nrow) {
        float v = fmaxf(stg[(r0 + i) * 33 + lane] + bias_v, lo);
        if (MASK) v = keep[i] ? v : 0.f;
        v += add[i];
        if (OUTF) { if (ATOMIC) atomicAdd(of + (int64_t)i * E.ld_f32, v); else of[(int64_t)i * E.ld_f32] = v; }
        if (OUTB) ob[(int64_t)i * E.ld_bf16] = __float2bfloat16_rn(v);
      }
    }
    if (MASK) mp += 8 * E.ld_mask;
    if (RESB) rb += 8 * E.ld_resid_bf16;
    if (RESF) rf += 8 * E.ld_resid_f32;
    if (OUTF) of += 8 * E.ld_f32;
    if (OUTB) ob += 8 * E.ld_bf16;
  }
}

// generic fallback: every feature decided at run time
__device__ __noinline__ void epi_block_generic(const UmmaEpilogue& E, const float* stg, int lane, int64_t row_base, int nrow, int col,
                                               float bias_v, float lo) {
  for (int r = 0; r < nrow; ++r) {
    const int64_t row = row_base + r;
    if (E.gelu) {          // AdapterBlock down-projection with GELU: the activation feeds the up-projection, the backward needs the pre-activation
      const float pre = stg[r * 33 + lane] + bias_v;
      if (E.out_f32) E.out_f32[row * E.ld_f32 + col] = pre;
      if (E.out_bf16) E.out_bf16[row * E.ld_bf16 + col] = __float2bfloat16_rn(gelu_erf(pre));
      continue;
    }
    float v = fmaxf(stg[r * 33 + lane] + bias_v, lo);
    if (E.gelu_pre) v *= gelu_erf_grad(E.gelu_pre[row * E.ld_gelu_pre + col]);
    if (E.mask) { if (!(__bfloat162float(E.mask[row * E.ld_mask + col]) > 0.f)) v = 0.f; }
    if (E.resid_bf16) v += __bfloat162float(E.resid_bf16[row * E.ld_resid_bf16 + col]);
    if (E.resid_f32) v += E.resid_f32[row * E.ld_resid_f32 + col];
    if (E.out_f32) { if (E.atomic) atomicAdd(E.out_f32 + row * E.ld_f32 + col, v); else E.out_f32[row * E.ld_f32 + col] = v; }
    if (E.out_bf16) E.out_bf16[row * E.ld_bf16 + col] = __float2bfloat16_rn(v);
  }
}

enum : int { EPI_MASK = 1, EPI_RESB = 2, EPI_RESF = 4, EPI_OUTF = 8, EPI_OUTB = 16, EPI_ATOMIC = 32, EPI_GELU = 64 };

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load that lands at the same shared-memory offset of every CTA in `mask` and completes bytes on each one's mbarrier
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// MC: the kernel runs as clusters of two CTAs that work on two row tiles of the same column tile: each CTA loads its own A tile
// and ONE HALF of the B tile, multicast into both CTAs (operand traffic from L2 per row tile: A + B/2 instead of A + B; the
// 768 x 768 heads are L2-bandwidth bound: 396 tiles x 589 KB).  A ring slot is refilled only when BOTH consumers released it.
template <int BN, bool A_MN, bool B_MN, int NP, bool MC>
__global__ void __launch_bounds__(UTHREADS, 1) umma_gemm_kernel(const __grid_constant__ UmmaDevBatchT<NP> batch) {
  const int NSTG = batch.stages;
  const int crank = MC ? (int)cluster_ctarank() : 0;
  const int tile_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = UmmaSmem<BN>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTG * S::kStageBytes);
  uint64_t* empty = full + USTAGES;
  uint64_t* accum_full = empty + USTAGES;        // [2]
  uint64_t* accum_empty = accum_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_empty + 2);
  float* staging = reinterpret_cast<float*>(smem + NSTG * S::kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MC ? 2 : 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&accum_full[b], 1); mbar_init(&accum_empty[b], UEPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();          // the peer's barriers exist before any multicast traffic / remote commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_first; tile < batch.total_tiles; tile += tile_step) {
        const UmmaTile t = umma_decode_tile<BN, NP, MC>(batch, tile, crank);
        const UmmaDevProblem& P = batch.p[t.prob];
        for (int kb = t.kb_beg; kb < t.kb_end; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full[stage], S::kStageBytes);
          const int k0 = kb * UBK;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < UBM / 64; ++i) tma_load_2d(sa + i * (64 * UBK * 2), &P.map_a, &full[stage], t.m0 + i * 64, k0);
          } else {
            tma_load_2d(sa, &P.map_a, &full[stage], k0, t.m0);
          }
          if (MC) {          // this CTA's half of the column tile, into both CTAs
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i) {
                const int bi = crank * (BN / 128) + i;
                tma_load_2d_mc(sb + bi * (64 * UBK * 2), &P.map_b, &full[stage], t.n0 + bi * 64, k0, (uint16_t)3);
              }
            } else {
              tma_load_2d_mc(sb + crank * (BN / 2) * UBK * 2, &P.map_b, &full[stage], k0, t.n0 + crank * (BN / 2), (uint16_t)3);
            }
          } else if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * (64 * UBK * 2), &P.map_b, &full[stage], t.n0 + i * 64, k0);
          } else {
            tma_load_2d(sb, &P.map_b, &full[stage], k0, t.n0);
          }
          if (++stage == NSTG) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_bf16(UBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < batch.total_tiles; tile += tile_step, ++it) {
        const UmmaTile t = umma_decode_tile<BN, NP, MC>(batch, tile, crank);
        const int ab = it & 1; const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&accum_empty[ab], aph ^ 1u);           // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tm = tmem_base + ab * BN;
        for (int kb = t.kb_beg; kb < t.kb_end; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < UBK / 16; ++k) {
            // K-major: step 16 elements (32 B) inside the 128-byte swizzle row; MN-major: step 16 k-rows (2048 B)
            const uint64_t ad = A_MN ? smem_desc_sw128(sa + k * 2048, 64 * UBK * 2, 1024) : smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? smem_desc_sw128(sb + k * 2048, 64 * UBK * 2, 1024) : smem_desc_sw128(sb + k * 32, 16, 1024);
            mma_bf16_ss(tm, ad, bd, idesc, (kb > t.kb_beg || k > 0) ? 1u : 0u);
          }
          if (MC) mma_commit_mc(&empty[stage], (uint16_t)3);   // both CTAs' producers write into this slot of both CTAs
          else mma_commit(&empty[stage]);            // frees the smem slot when these MMAs retire
          if (++stage == NSTG) { stage = 0; phase ^= 1; }
        }
        mma_commit(&accum_full[ab]);
      }
    }
  } else {
    // ---- epilogue: 8 warps; warp w may touch TMEM lanes [32*(w%4), +32) == output rows m0 + 32*(w%4) + lane; the two warps
    //      of a quadrant take alternate 32-column chunks.  Each 32x32 block is transposed through padded shared memory so
    //      that global loads (bias / mask / residual) and stores are row-contiguous (coalesced 64/128-byte segments). ----
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int csel = ew >> 2;
    const int lane = threadIdx.x & 31;
    float* stg = staging + ew * USTG_FLOATS;
    int it = 0;
    for (int tile = tile_first; tile < batch.total_tiles; tile += tile_step, ++it) {
      const UmmaTile t = umma_decode_tile<BN, NP, MC>(batch, tile, crank);
      const UmmaDevProblem& P = batch.p[t.prob];
      const int m0 = t.m0, n0 = t.n0;
      const int ab = it & 1; const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&accum_full[ab], aph);
      tc_fence_after();
      const uint32_t tm = tmem_base + ab * BN;
      const UmmaEpilogue& E = P.epi;
      const bool first_split = (t.split == 0);
      const int row_base = m0 + quad * 32;
      const int features = (E.mask ? EPI_MASK : 0) | (E.resid_bf16 ? EPI_RESB : 0) | (E.resid_f32 ? EPI_RESF : 0) | (E.out_f32 ? EPI_OUTF : 0) |
                           (E.out_bf16 ? EPI_OUTB : 0) | ((E.atomic && E.out_f32) ? EPI_ATOMIC : 0) |
                           ((E.gelu || E.gelu_pre) ? EPI_GELU : 0);      // GELU: the run-time generic block below
#pragma unroll 1
      for (int c0 = csel * 32; c0 < BN; c0 += 64) {
        if (n0 + c0 >= P.N) break;                   // warp-uniform
        uint32_t raw[32];
        tmem_ld_32x32(tm + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
        tmem_ld_wait();
        if (E.transpose_out) {
          // out[n * ld + m]: the lanes (rows m) are already contiguous in memory
          const int row = row_base + lane;
          if (row < P.M) {
            const int ncol = min(32, P.N - (n0 + c0));
            float* op = E.out_f32 + (int64_t)(n0 + c0) * E.ld_f32 + row;
            if (E.atomic) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncol) atomicAdd(op + (int64_t)j * E.ld_f32, __uint_as_float(raw[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < ncol) op[(int64_t)j * E.ld_f32] = __uint_as_float(raw[j]);
            }
          }
          continue;
        }
        // ---- direct path (one output, no mask / residual): the thread owns 32 consecutive columns of its row, i.e. a 64 B (bf16)
        //      or 128 B (fp32) contiguous segment -> 128-bit stores / vector reductions instead of 32 two-byte stores per lane ----
        if ((features == EPI_OUTB || features == EPI_OUTF || features == (EPI_ATOMIC | EPI_OUTF)) && n0 + c0 + 32 <= P.N) {
          const int row = row_base + lane;
          if (row < P.M) {
            float v[32];
            const bool add_bias = E.bias && (!E.atomic || first_split);
            const float lo = E.relu ? 0.f : -INFINITY;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (add_bias) b4 = __ldg(reinterpret_cast<const float4*>(E.bias + n0 + c0) + q);
              v[4 * q] = fmaxf(__uint_as_float(raw[4 * q]) + b4.x, lo); v[4 * q + 1] = fmaxf(__uint_as_float(raw[4 * q + 1]) + b4.y, lo);
              v[4 * q + 2] = fmaxf(__uint_as_float(raw[4 * q + 2]) + b4.z, lo); v[4 * q + 3] = fmaxf(__uint_as_float(raw[4 * q + 3]) + b4.w, lo);
            }
            if (features == EPI_OUTB) {
              __nv_bfloat16* ob = E.out_bf16 + (int64_t)row * E.ld_bf16 + n0 + c0;
              if ((reinterpret_cast<uintptr_t>(ob) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 pk;
                  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                  for (int z = 0; z < 4; ++z) h2[z] = __floats2bfloat162_rn(v[8 * q + 2 * z], v[8 * q + 2 * z + 1]);
                  reinterpret_cast<uint4*>(ob)[q] = pk;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) ob[j] = __float2bfloat16_rn(v[j]);
              }
            } else {
              float* of = E.out_f32 + (int64_t)row * E.ld_f32 + n0 + c0;
              const bool al = (reinterpret_cast<uintptr_t>(of) & 15) == 0;
              if (features == EPI_OUTF) {
                if (al) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(of)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) of[j] = v[j];
                }
              } else if (al) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(of + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]),
                               "f"(v[4 * q + 3])
                               : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(of + j, v[j]);
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(raw[j]);
        __syncwarp();
        const int col = n0 + c0 + lane;
        const bool col_ok = col < P.N;
        const float bias_v = (E.bias && col_ok && (!E.atomic || first_split)) ? __ldg(E.bias + col) : 0.f;
        const int nrow = min(32, P.M - row_base);    // warp-uniform
        if (col_ok && nrow > 0) {
          const float lo = E.relu ? 0.f : -INFINITY;
          switch (features) {
            case EPI_OUTB: epi_block<false, false, false, false, true, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_RESB | EPI_OUTB: epi_block<false, true, false, false, true, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_OUTF: epi_block<false, false, false, true, false, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_OUTF | EPI_OUTB: epi_block<false, false, false, true, true, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_MASK | EPI_OUTB: epi_block<true, false, false, false, true, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_RESF | EPI_OUTF | EPI_OUTB: epi_block<false, false, true, true, true, false>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            case EPI_ATOMIC | EPI_OUTF: epi_block<false, false, false, true, false, true>(E, stg, lane, row_base, nrow, col, bias_v, lo); break;
            default: epi_block_generic(E, stg, lane, row_base, nrow, col, bias_v, lo);
          }
        }
        __syncwarp();
      }
      // accumulator drained (all tcgen05.ld of this warp have completed)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&accum_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int64_t rows, cols, pitch; int box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](int64_t v) { h ^= std::hash<int64_t>()(v) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.pitch); mix(k.box_inner); mix(k.box_outer);
    return h;
  }
};

// 2-D bf16 tensor map over a row-major [rows, cols] matrix, 128-byte swizzle, zero fill out of bounds.
int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_inner, int box_outer) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, rows, cols, pitch, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return IISAN_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return IISAN_ECUDA;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((pitch * 2) & 15)) return IISAN_EINVAL;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return IISAN_EINVAL;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return IISAN_OK;
}

// kernel class the next launches are accounted under (san_lr.cu times its hidden-state GEMM pass separately)
thread_local int g_umma_launch_class = IISAN_K_GEMM;

static const bool g_no_multicast = [] { const char* e = getenv("IISAN_B200_NO_MULTICAST"); return e && e[0] == '1'; }();

template <int BN, bool A_MN, bool B_MN, int NP, bool MC = false>
static int launch_cfg(const UmmaProblem* probs, int n_probs, cudaStream_t st) {
  static thread_local UmmaDevBatchT<NP> dev;      // large for the batched variant: keep it off the stack (launches are serialised per thread by the callers)
  int total_tiles = 0, max_kb = 1;
  for (int i = 0; i < n_probs; ++i) {
    const UmmaProblem& P = probs[i];
    UmmaDevProblem& D = dev.p[i];
    if (P.M <= 0 || P.N <= 0 || P.K <= 0 || (P.N % 8)) return IISAN_EINVAL;
    // A
    if (A_MN) IISAN_TRY(make_tensor_map_bf16(&D.map_a, P.A.ptr, P.A.rows, P.A.cols, P.A.pitch, 64, UBK));
    else IISAN_TRY(make_tensor_map_bf16(&D.map_a, P.A.ptr, P.A.rows, P.A.cols, P.A.pitch, UBK, UBM));
    if (B_MN) IISAN_TRY(make_tensor_map_bf16(&D.map_b, P.B.ptr, P.B.rows, P.B.cols, P.B.pitch, 64, UBK));
    else IISAN_TRY(make_tensor_map_bf16(&D.map_b, P.B.ptr, P.B.rows, P.B.cols, P.B.pitch, UBK, MC ? BN / 2 : BN));
    D.M = P.M; D.N = P.N; D.K = P.K; D.epi = P.epi;
    const int kb_total = (P.K + UBK - 1) / UBK;
    int split = P.splitk < 1 ? 1 : (P.splitk > kb_total ? kb_total : P.splitk);
    D.kblocks_per_split = (kb_total + split - 1) / split;
    split = (kb_total + D.kblocks_per_split - 1) / D.kblocks_per_split;
    if (split > 1 && !(P.epi.atomic && P.epi.out_f32 && !P.epi.out_bf16)) return IISAN_EINVAL;
    const int tiles_m = (P.M + UBM - 1) / UBM;
    const int tiles = (MC ? (tiles_m + 1) / 2 : tiles_m) * ((P.N + BN - 1) / BN);
    D.tile_start = total_tiles;
    total_tiles += tiles * split;
    if (D.kblocks_per_split > max_kb) max_kb = D.kblocks_per_split;
  }
  dev.n_probs = n_probs; dev.total_tiles = total_tiles;
  static std::atomic<uint64_t> attr_done{0};      // devices on which the attribute has been set
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<BN, A_MN, B_MN, NP, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaSmem<BN>::total(USTAGES)));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  // one CTA per SM: the stage count no longer has to leave room for a second CTA (the ring runs across tile boundaries); the
  // full-size ring also keeps a second CTA (and its 2 * BN TMEM columns) off the SM
  dev.stages = USTAGES;
  (void)max_kb;
  static int n_sm = 0;
  if (n_sm == 0) {
    int devid = 0;
    IISAN_CUDA_OK(cudaGetDevice(&devid));
    IISAN_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, devid));
  }
  if (MC) {
    const int pairs = total_tiles < n_sm / 2 ? total_tiles : n_sm / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(UTHREADS); cfg.dynamicSmemBytes = UmmaSmem<BN>::total(dev.stages); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    { LaunchScope ls_(g_umma_launch_class, st); IISAN_CUDA_OK(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<BN, A_MN, B_MN, NP, MC>, dev)); }
    IISAN_LAUNCH_OK();
    return IISAN_OK;
  }
  const int grid = total_tiles < n_sm ? total_tiles : n_sm;
  { LaunchScope ls_(g_umma_launch_class, st); umma_gemm_kernel<BN, A_MN, B_MN, NP, MC><<<grid, UTHREADS, UmmaSmem<BN>::total(dev.stages), st>>>(dev); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

int launch_umma_gemm(const UmmaBatch& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  if (b.n > kUmmaMaxProbs) return IISAN_EINVAL;
  const bool a_mn = b.p[0].a_mn_major != 0, b_mn = b.p[0].b_mn_major != 0;
  int maxN = 0;
  for (int i = 0; i < b.n; ++i) {
    if ((b.p[i].a_mn_major != 0) != a_mn || (b.p[i].b_mn_major != 0) != b_mn) return IISAN_EINVAL;
    if (b.p[i].N > maxN) maxN = b.p[i].N;
  }
  if (a_mn && !b_mn) return IISAN_EUNSUPPORTED;
  int bn = maxN <= 64 ? 64 : (maxN <= 128 ? 128 : 256);
  if (bn == 256) {   // fewer 256-wide tiles than SMs: 128-wide tiles spread the work over more of them
    int64_t tiles = 0;
    for (int i = 0; i < b.n; ++i) tiles += (int64_t)((b.p[i].M + UBM - 1) / UBM) * ((b.p[i].N + 255) / 256) * (b.p[i].splitk < 1 ? 1 : b.p[i].splitk);
    if (tiles < 148) bn = 128;
  }
  // Large launches of 256-wide tiles (>= 4 tiles per SM: operand traffic from L2 is what bounds them): clusters of two CTAs share
  // the column tile by TMA multicast.  Measured: 8192^3 1250 -> 1297 TFLOP/s; the 396-tile head GEMMs of the base config are
  // unchanged within noise (0.282 vs 0.286 ms for the whole GEMM class), so they keep the simpler single-CTA schedule.
  bool mc = (bn == 256) && !g_no_multicast;
  int64_t tiles256 = 0;
  for (int i = 0; i < b.n; ++i) {
    mc = mc && (b.p[i].M > UBM);
    tiles256 += (int64_t)((b.p[i].M + UBM - 1) / UBM) * ((b.p[i].N + 255) / 256) * (b.p[i].splitk < 1 ? 1 : b.p[i].splitk);
  }
  mc = mc && tiles256 >= 4 * 148;
  if (!a_mn && b_mn) {       // data gradient x W with the nn.Linear weight [out, in] read in place as a [K, N] operand
    if (bn == 64) return launch_cfg<64, false, true, kUmmaMaxProbs>(b.p, b.n, st);
    if (bn == 128) return launch_cfg<128, false, true, kUmmaMaxProbs>(b.p, b.n, st);
    return mc ? launch_cfg<256, false, true, kUmmaMaxProbs, true>(b.p, b.n, st) : launch_cfg<256, false, true, kUmmaMaxProbs>(b.p, b.n, st);
  }
  if (!a_mn) {
    if (bn == 64) return launch_cfg<64, false, false, kUmmaMaxProbs>(b.p, b.n, st);
    if (bn == 128) return launch_cfg<128, false, false, kUmmaMaxProbs>(b.p, b.n, st);
    return mc ? launch_cfg<256, false, false, kUmmaMaxProbs, true>(b.p, b.n, st) : launch_cfg<256, false, false, kUmmaMaxProbs>(b.p, b.n, st);
  }
  if (bn == 64) return launch_cfg<64, true, true, kUmmaMaxProbs>(b.p, b.n, st);
  if (bn == 128) return launch_cfg<128, true, true, kUmmaMaxProbs>(b.p, b.n, st);
  return mc ? launch_cfg<256, true, true, kUmmaMaxProbs, true>(b.p, b.n, st) : launch_cfg<256, true, true, kUmmaMaxProbs>(b.p, b.n, st);
}

int launch_umma_gemm_big(const UmmaBatchBig& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  if (b.n > kUmmaBigProbs) return IISAN_EINVAL;
  for (int i = 0; i < b.n; ++i)
    if (!b.p[i].a_mn_major || !b.p[i].b_mn_major || b.p[i].N > 64) return IISAN_EUNSUPPORTED;
  return launch_cfg<64, true, true, kUmmaBigProbs>(b.p, b.n, st);
}

int launch_umma_gemm_many(const UmmaBatchBig& b, cudaStream_t st) {
  if (b.n <= 0) return IISAN_OK;
  if (b.n > kUmmaBigProbs) return IISAN_EINVAL;
  const bool a_mn = b.p[0].a_mn_major != 0, b_mn = b.p[0].b_mn_major != 0;
  int maxN = 0;
  for (int i = 0; i < b.n; ++i) {
    if ((b.p[i].a_mn_major != 0) != a_mn || (b.p[i].b_mn_major != 0) != b_mn) return IISAN_EINVAL;
    if (b.p[i].N > maxN) maxN = b.p[i].N;
  }
  if (a_mn && !b_mn) return IISAN_EUNSUPPORTED;
  const bool wide = maxN > 64;
  // (Clusters of two CTAs sharing the column tile by TMA multicast were measured on the 17-problem hidden-state pass of san_lr.cu:
  // 93.6 vs 95.5 us -- the 128 x 256 tile is bound by the shared-memory operand fetch of the MMA, not by L2 -- and are not used here.)
  if (a_mn) return wide ? launch_cfg<256, true, true, kUmmaBigProbs>(b.p, b.n, st) : launch_cfg<64, true, true, kUmmaBigProbs>(b.p, b.n, st);
  if (b_mn) return wide ? launch_cfg<256, false, true, kUmmaBigProbs>(b.p, b.n, st) : launch_cfg<64, false, true, kUmmaBigProbs>(b.p, b.n, st);
  return wide ? launch_cfg<256, false, false, kUmmaBigProbs>(b.p, b.n, st) : launch_cfg<64, false, false, kUmmaBigProbs>(b.p, b.n, st);
}

}  // namespace iisan
