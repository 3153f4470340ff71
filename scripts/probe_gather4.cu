// Standalone probe (not on the product path): does cp.async.bulk.tensor.2d ... tile::gather4 deliver the rows of an item table
// into the SAME 128-byte-swizzled [128 x 64] bf16 tile that one ordinary 2-D TMA box load of 128 consecutive rows produces, and
// how fast is a tile assembled from 32 gather4 instructions (one per lane of a warp) against one box load?
//
// Why: with the HBM-resident cached-state store (iisan_b200/store.py) a batch is a list of item ids; today iisan_gather_states
// materialises the [N, A, d] batch (121 MB written + re-read per step at B = 512) before the chain kernels stream it.  If the
// answer here is "same tile, same speed", the hidden-state producers of san_chain3.cu / umma_gemm.cu can read the store directly.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o iisan_b200/lib/probe_gather4 scripts/probe_gather4.cu
//   ./iisan_b200/lib/probe_gather4            -> JSON lines on stdout
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
static __device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// four rows r0..r3 of the 2-D tensor, columns [c0, c0 + box0): lands as four consecutive 128-byte rows at dst
static __device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

constexpr int ROWS = 128, CW = 64, TILE_BYTES = ROWS * CW * 2;

// ---- correctness: one tile, dumped raw ----
__global__ void dump_kernel(const __grid_constant__ CUtensorMap map_box, const __grid_constant__ CUtensorMap map_row, const int* __restrict__ idx,
                            int col, int mode, uint4* __restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TILE_BYTES + 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane == 0) mbar_expect_tx(bar, TILE_BYTES);
    __syncwarp();
    if (mode == 0) {
      if (lane == 0) tma_load_2d(smem, &map_box, bar, col, idx[0]);        // 128 consecutive rows starting at idx[0]
    } else {
      tma_gather4(smem + lane * 512, &map_row, bar, col, idx[4 * lane], idx[4 * lane + 1], idx[4 * lane + 2], idx[4 * lane + 3]);
    }
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < TILE_BYTES / 16; i += blockDim.x) out[i] = reinterpret_cast<const uint4*>(smem)[i];
}

// ---- throughput: every CTA streams `n_tiles` tiles through a ring; consumer warps touch one word per thread ----
constexpr int SLOTS = 8, CONS = 4;
__global__ void __launch_bounds__(32 * (1 + CONS), 1)
stream_kernel(const __grid_constant__ CUtensorMap map_box, const __grid_constant__ CUtensorMap map_row, const int* __restrict__ idx,
              int n_rows_batch, int n_chunks, int n_layers, int layers_total, int mode, unsigned long long* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SLOTS * TILE_BYTES);
  uint64_t* empty = full + SLOTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int m0 = blockIdx.x * ROWS;
  const int n_tiles = n_layers * n_chunks;
  if (warp == 0) {
    // rows of this CTA's tile: table row = item * layers_total + layer (the packed store layout [items, A, d] seen as 2-D)
    int it[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int r = m0 + 4 * lane + k; it[k] = r < n_rows_batch ? idx[r] : 0; }
    for (int t = 0; t < n_tiles; ++t) {
      const int slot = t % SLOTS; const uint32_t ph = (uint32_t)(t / SLOTS) & 1u;
      const int l = t / n_chunks, c = t % n_chunks;
      mbar_wait(&empty[slot], ph ^ 1u);
      if (lane == 0) mbar_expect_tx(&full[slot], TILE_BYTES);
      __syncwarp();
      if (mode == 0) {          // ordinary box load from a materialised [N, A*d] batch
        if (lane == 0) tma_load_2d(smem + slot * TILE_BYTES, &map_box, &full[slot], l * n_chunks * CW + c * CW, m0);
      } else {
        tma_gather4(smem + slot * TILE_BYTES + lane * 512, &map_row, &full[slot], c * CW, it[0] * layers_total + l, it[1] * layers_total + l,
                    it[2] * layers_total + l, it[3] * layers_total + l);
      }
    }
  } else {
    unsigned long long acc = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int slot = t % SLOTS; const uint32_t ph = (uint32_t)(t / SLOTS) & 1u;
      mbar_wait(&full[slot], ph);
      acc += *reinterpret_cast<const unsigned int*>(smem + slot * TILE_BYTES + ((warp - 1) * 32 + lane) * 128);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn fn, CUtensorMap* m, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_c, uint32_t box_r) {
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_c, box_r};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  if (!fn) { printf("{\"error\": \"no cuTensorMapEncodeTiled\"}\n"); return 2; }

  // ---------------- correctness ----------------
  const int R = 4096, C = 768;
  std::vector<__nv_bfloat16> h((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
  __nv_bfloat16* d_tab; CK(cudaMalloc(&d_tab, h.size() * 2)); CK(cudaMemcpy(d_tab, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap map_box, map_row;
  int e1 = make_map(fn, &map_box, d_tab, R, C, CW, ROWS), e2 = make_map(fn, &map_row, d_tab, R, C, CW, 1);
  if (e1 || e2) { printf("{\"error\": \"encode failed %d %d\"}\n", e1, e2); return 2; }
  std::vector<int> idx(ROWS);
  uint4* d_out; CK(cudaMalloc(&d_out, TILE_BYTES));
  int* d_idx; CK(cudaMalloc(&d_idx, ROWS * 4));
  std::vector<uint16_t> got(TILE_BYTES / 2);
  CK(cudaFuncSetAttribute(dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 2048));
  const int col = 3 * CW;
  for (int variant = 0; variant < 3; ++variant) {
    // 0: box load of rows 256..383 ; 1: gather4 with the SAME consecutive rows ; 2: gather4 with scattered rows (repeats, row 0)
    for (int i = 0; i < ROWS; ++i) idx[i] = variant < 2 ? 256 + i : (int)(((unsigned)i * 2654435761u) % R);
    if (variant == 2) { idx[5] = 0; idx[6] = 0; idx[77] = idx[3]; }
    CK(cudaMemcpy(d_idx, idx.data(), ROWS * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xEE, TILE_BYTES));
    dump_kernel<<<1, 128, TILE_BYTES + 2048>>>(map_box, map_row, d_idx, col, variant == 0 ? 0 : 1, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"check\": %d, \"error\": \"%s\"}\n", variant, cudaGetErrorString(e)); return 3; }
    CK(cudaMemcpy(got.data(), d_out, TILE_BYTES, cudaMemcpyDeviceToHost));
    // expected: row m at (m >> 3) * 1024 + (m & 7) * 128, its 16-byte group g at ((g ^ (m & 7)) << 4)
    int bad = 0, first_bad = -1;
    for (int m = 0; m < ROWS; ++m)
      for (int c = 0; c < CW; ++c) {
        const int g = c >> 3, w = c & 7;
        const size_t off = (size_t)(m >> 3) * 1024 + (m & 7) * 128 + ((g ^ (m & 7)) << 4) + w * 2;
        const __nv_bfloat16 ex = h[(size_t)idx[m] * C + col + c];
        if (got[off / 2] != *reinterpret_cast<const uint16_t*>(&ex)) { if (first_bad < 0) first_bad = m * CW + c; ++bad; }
      }
    printf("{\"check\": \"%s\", \"mismatches\": %d, \"first_bad\": %d}\n",
           variant == 0 ? "box_load_consecutive" : (variant == 1 ? "gather4_consecutive" : "gather4_scattered"), bad, first_bad);
  }

  // ---------------- throughput: Instrument shape (5632 rows, 7 layers, d = 768), catalogue of 19,247 items ----------------
  const int N = 5632, A = 7, ITEMS = 19247, NCH = C / CW;
  __nv_bfloat16 *d_store, *d_batch;
  CK(cudaMalloc(&d_store, (size_t)ITEMS * A * C * 2)); CK(cudaMemset(d_store, 0, (size_t)ITEMS * A * C * 2));
  CK(cudaMalloc(&d_batch, (size_t)N * A * C * 2)); CK(cudaMemset(d_batch, 0, (size_t)N * A * C * 2));
  std::vector<int> ids(N);
  for (int i = 0; i < N; ++i) ids[i] = 1 + (int)(((unsigned)i * 2246822519u + 12345u) % (ITEMS - 1));
  int* d_ids; CK(cudaMalloc(&d_ids, N * 4)); CK(cudaMemcpy(d_ids, ids.data(), N * 4, cudaMemcpyHostToDevice));
  CUtensorMap mb, mr;
  if (make_map(fn, &mb, d_batch, N, (uint64_t)A * C, CW, ROWS) || make_map(fn, &mr, d_store, (uint64_t)ITEMS * A, C, CW, 1)) {
    printf("{\"error\": \"encode failed (stream)\"}\n"); return 2;
  }
  unsigned long long* d_sink; CK(cudaMalloc(&d_sink, 8));
  void* flush; const size_t FL = 256u << 20; CK(cudaMalloc(&flush, FL));
  const int smem = SLOTS * TILE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int mode = 0; mode < 2; ++mode) {
    float best = 1e9f, sum = 0.f; const int reps = 10;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaMemsetAsync(flush, r, FL));          // evict L2 (126 MB)
      CK(cudaEventRecord(a));
      stream_kernel<<<N / ROWS, 32 * (1 + CONS), smem>>>(mb, mr, d_ids, N, NCH, A, A, mode, d_sink);
      CK(cudaEventRecord(b));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"stream_mode\": %d, \"error\": \"%s\"}\n", mode, cudaGetErrorString(e)); return 3; }
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (r >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    const double bytes = (double)N * A * C * 2;
    printf("{\"stream\": \"%s\", \"ctas\": %d, \"bytes\": %.0f, \"us_avg\": %.1f, \"us_min\": %.1f, \"gbs_avg\": %.0f}\n",
           mode == 0 ? "box_load_materialised_batch" : "gather4_from_store", N / ROWS, bytes, sum / reps * 1e3, best * 1e3,
           bytes / (sum / reps * 1e-3) / 1e9);
  }
  return 0;
}
