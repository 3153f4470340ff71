// Measurement probe (not on the product path): the access pattern of the fused chain kernels' hidden-state tiles, alone.
//
// A ring tile of san_chain.cu is a TMA box of 128 item rows x 64 bf16 columns: 128 separate 128-byte segments, one per item,
// at the row pitch of the caller's [N, layers, d] tensor (19,968 B for the BERT-base / ViT-B/16 states).  This kernel issues
// exactly those loads -- same tensor map (128-byte swizzle, 64 x 128 box), same per-CTA order (all chunks of the first selected
// layer, then of the next ...), a ring of `slots` 16 KB tiles, one CTA per 128-item tile -- and nothing else: consumer warps
// read one 16-byte word per thread and release the slot.  With `contiguous != 0` the same number of tiles is read from a
// tile-contiguous layout [N/128, A, d/64, 128, 64] (one 16 KB burst per tile), which is what the HBM-resident store could
// provide.  The ratio of the two GB/s figures says whether in-place streaming of the reference layout is limited by the
// 128-byte segment granularity (first measurement of round 2).
#include "common.cuh"
#include "launch.cuh"
#include "umma.cuh"

namespace iisan {

using namespace umma;

int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t pitch, int box_inner, int box_outer);

constexpr int PB_ROWS = 128, PB_CW = 64, PB_TILE_BYTES = PB_ROWS * PB_CW * 2;
constexpr int PB_MAX_SLOTS = 12;
constexpr int PB_CONSUMER_WARPS = 4;
constexpr int PB_THREADS = 32 * (1 + PB_CONSUMER_WARPS);
constexpr int PB_MAX_LAYERS = 16;

struct ProbeArgs {
  CUtensorMap map;
  int n_chunks;            // d / 64
  int n_layers;            // selected layers
  int layer[PB_MAX_LAYERS];
  int d;
  int slots;
  int repeat;              // passes over the (layer, chunk) list: 2 emulates the inter-modal tower's second read (L2 hits)
  int contiguous;
  unsigned long long* sink;
};

__global__ void __launch_bounds__(PB_THREADS, 1) tile_stream_probe_kernel(const __grid_constant__ ProbeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.slots * PB_TILE_BYTES);
  uint64_t* empty = full + PB_MAX_SLOTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&a.map);
    for (int s = 0; s < a.slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], PB_CONSUMER_WARPS); }
    fence_barrier_init();
  }
  __syncthreads();
  const int tiles_per_pass = a.n_layers * a.n_chunks;
  const int n_tiles = tiles_per_pass * a.repeat;
  const int m_tile = blockIdx.x;
  if (warp == 0) {
    if (elect_one()) {
      for (int t = 0; t < n_tiles; ++t) {
        const int slot = t % a.slots; const uint32_t ph = (uint32_t)(t / a.slots) & 1u;
        const int q = t % tiles_per_pass;
        const int s = q / a.n_chunks, c = q % a.n_chunks;
        mbar_wait(&empty[slot], ph ^ 1u);
        mbar_expect_tx(&full[slot], PB_TILE_BYTES);
        if (a.contiguous) tma_load_2d(smem + slot * PB_TILE_BYTES, &a.map, &full[slot], 0, ((m_tile * a.n_layers + s) * a.n_chunks + c) * PB_ROWS);
        else tma_load_2d(smem + slot * PB_TILE_BYTES, &a.map, &full[slot], a.layer[s] * a.d + c * PB_CW, m_tile * PB_ROWS);
      }
    }
  } else {
    unsigned long long acc = 0;
    const int ct = threadIdx.x - 32;                      // 0 .. 127: one 16-byte word of row ct per tile
    for (int t = 0; t < n_tiles; ++t) {
      const int slot = t % a.slots; const uint32_t ph = (uint32_t)(t / a.slots) & 1u;
      mbar_wait(&full[slot], ph);
      const uint4 v = *reinterpret_cast<const uint4*>(smem + slot * PB_TILE_BYTES + ct * 128 + ((t & 7) << 4));
      acc ^= ((unsigned long long)v.x << 32 | v.y) + ((unsigned long long)v.z << 32 | v.w);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 0x9e3779b97f4a7c15ull) a.sink[0] = acc;    // keeps the loads alive; practically never true
  }
}

}  // namespace iisan

using namespace iisan;

extern "C" int iisan_probe_tile_stream(const void* base, int64_t n_rows, int32_t layers, int32_t d, const int32_t* sel, int32_t n_sel,
                                       int32_t slots, int32_t repeat, int32_t contiguous, void* sink, iisan_stream_t stream) {
  if (!base || !sel || !sink || n_rows <= 0 || layers <= 0 || d <= 0 || d % PB_CW || n_sel <= 0 || n_sel > PB_MAX_LAYERS || slots < 1 ||
      slots > PB_MAX_SLOTS || repeat < 1)
    return IISAN_EINVAL;
  ProbeArgs a{};
  const int tiles = (int)((n_rows + PB_ROWS - 1) / PB_ROWS);
  a.n_chunks = d / PB_CW; a.n_layers = n_sel; a.d = d; a.slots = slots; a.repeat = repeat; a.contiguous = contiguous ? 1 : 0;
  a.sink = static_cast<unsigned long long*>(sink);
  for (int i = 0; i < n_sel; ++i) {
    if (sel[i] < 0 || sel[i] >= layers) return IISAN_EINVAL;
    a.layer[i] = sel[i];
  }
  if (contiguous) {
    // `base` holds tiles * n_sel * (d / 64) tiles of [128, 64] bf16 back to back
    const int64_t rows = (int64_t)tiles * n_sel * a.n_chunks * PB_ROWS;
    IISAN_TRY(make_tensor_map_bf16(&a.map, base, rows, PB_CW, PB_CW, PB_CW, PB_ROWS));
  } else {
    IISAN_TRY(make_tensor_map_bf16(&a.map, base, n_rows, (int64_t)layers * d, (int64_t)layers * d, PB_CW, PB_ROWS));
  }
  const int smem = slots * PB_TILE_BYTES + 2 * PB_MAX_SLOTS * 8 + 1024;
  static std::atomic<uint64_t> attr_done{0};
  const uint64_t dev_bit = device_bit();
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(tile_stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PB_MAX_SLOTS * PB_TILE_BYTES + 2 * PB_MAX_SLOTS * 8 + 1024));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  cudaStream_t st = as_stream(stream);
  { LaunchScope ls_(IISAN_K_MISC, st); tile_stream_probe_kernel<<<tiles, PB_THREADS, smem, st>>>(a); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}
