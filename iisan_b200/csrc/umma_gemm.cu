// tcgen05 / TMEM / TMA GEMM (see umma_gemm.cuh).  One CTA computes one 128 x BN output tile:
//   warp 0   : TMA producer   (one elected lane streams A/B k-blocks through a 4-stage smem ring)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, commits to mbarriers)
//   warps 2-5: epilogue       (tcgen05.ld 32 lanes x 32 columns -> bias / ReLU / mask / residual -> global)
#include "umma_gemm.cuh"

#include <mutex>
#include <unordered_map>

#include "launch.cuh"
#include "umma.cuh"

namespace iisan {

using namespace umma;

constexpr int UBM = 128;          // rows per tile (UMMA M)
constexpr int UBK = 64;           // k-block: one 128-byte swizzle atom of bf16
constexpr int USTAGES = 4;
constexpr int UTHREADS = 192;

struct UmmaDevProblem {
  CUtensorMap map_a, map_b;
  int M, N, K;
  int kblocks_per_split;
  UmmaEpilogue epi;
};
struct UmmaDevBatch {
  UmmaDevProblem p[kUmmaMaxProbs];
};

template <int BN>
struct UmmaSmem {
  static constexpr int kABytes = UBM * UBK * 2;
  static constexpr int kBBytes = BN * UBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = USTAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(UTHREADS, 1) umma_gemm_kernel(const __grid_constant__ UmmaDevBatch batch) {
  const UmmaDevProblem& P = batch.p[blockIdx.z];
  const int tiles_n = (P.N + BN - 1) / BN;
  const int tiles_m = (P.M + UBM - 1) / UBM;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int kb_total = (P.K + UBK - 1) / UBK;
  const int kb_beg = blockIdx.y * P.kblocks_per_split;
  const int kb_end = min(kb_total, kb_beg + P.kblocks_per_split);
  if (kb_beg >= kb_end) return;
  const int m0 = (blockIdx.x / tiles_n) * UBM, n0 = (blockIdx.x % tiles_n) * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  using S = UmmaSmem<BN>;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + USTAGES * S::kStageBytes);
  uint64_t* empty = full + USTAGES;
  uint64_t* accum_full = empty + USTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.map_a);
    tma_prefetch_desc(&P.map_b);
    for (int s = 0; s < USTAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full[stage], S::kStageBytes);
        const int k0 = kb * UBK;
        if (A_MN) {
#pragma unroll
          for (int i = 0; i < UBM / 64; ++i) tma_load_2d(sa + i * (64 * UBK * 2), &P.map_a, &full[stage], m0 + i * 64, k0);
        } else {
          tma_load_2d(sa, &P.map_a, &full[stage], k0, m0);
        }
        if (B_MN) {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * (64 * UBK * 2), &P.map_b, &full[stage], n0 + i * 64, k0);
        } else {
          tma_load_2d(sb, &P.map_b, &full[stage], k0, n0);
        }
        if (++stage == USTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc_bf16(UBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_beg; kb < kb_end; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < UBK / 16; ++k) {
          // K-major: step 16 elements (32 B) inside the 128-byte swizzle row; MN-major: step 16 k-rows (2048 B)
          const uint64_t ad = A_MN ? smem_desc_sw128(sa + k * 2048, 64 * UBK * 2, 1024) : smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? smem_desc_sw128(sb + k * 2048, 64 * UBK * 2, 1024) : smem_desc_sw128(sb + k * 32, 16, 1024);
          mma_bf16_ss(tmem_base, ad, bd, idesc, (kb > kb_beg || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty[stage]);                 // frees the smem slot when these MMAs retire
        if (++stage == USTAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(accum_full);
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32) == output rows m0 + 32*(w%4) + lane ----
    const int quad = warp & 3;
    const int lane = threadIdx.x & 31;
    const int row = m0 + quad * 32 + lane;
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const UmmaEpilogue& E = P.epi;
    const bool first_split = (blockIdx.y == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= P.N) break;                   // warp-uniform
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
      tmem_ld_wait();
      if (row < P.M) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const int ncol = min(32, P.N - (n0 + c0));
        const int col = n0 + c0;
        if (E.bias && (!E.atomic || first_split)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < ncol) v[j] += __ldg(E.bias + col + j);
        }
        if (E.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (E.mask) {
          const __nv_bfloat16* mp = E.mask + (int64_t)row * E.ld_mask + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncol) {
              const uint4 q = *reinterpret_cast<const uint4*>(mp + j);
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
              for (int t = 0; t < 8; ++t) if (!(__bfloat162float(h[t]) > 0.f)) v[j + t] = 0.f;
            }
          }
        }
        if (E.resid_bf16) {
          const __nv_bfloat16* rp = E.resid_bf16 + (int64_t)row * E.ld_resid_bf16 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncol) {
              const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
              for (int t = 0; t < 8; ++t) v[j + t] += __bfloat162float(h[t]);
            }
          }
        }
        if (E.resid_f32) {
          const float* rp = E.resid_f32 + (int64_t)row * E.ld_resid_f32 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncol) {
              const float4 q = *reinterpret_cast<const float4*>(rp + j);
              v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
            }
          }
        }
        if (E.out_f32 && E.transpose_out) {
          float* op = E.out_f32 + (int64_t)col * E.ld_f32 + row;      // lanes (rows) are contiguous: coalesced
          if (E.atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) atomicAdd(op + (int64_t)j * E.ld_f32, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) op[(int64_t)j * E.ld_f32] = v[j];
          }
        } else if (E.out_f32) {
          float* op = E.out_f32 + (int64_t)row * E.ld_f32 + col;
          if (E.atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncol) atomicAdd(op + j, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) if (j < ncol) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        if (E.out_bf16) {
          __nv_bfloat16* op = E.out_bf16 + (int64_t)row * E.ld_bf16 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j < ncol) {
              uint4 q;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
              for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
              *reinterpret_cast<uint4*>(op + j) = q;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
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

template <int BN, bool A_MN, bool B_MN>
static int launch_cfg(const UmmaBatch& b, cudaStream_t st) {
  UmmaDevBatch dev;
  int max_tiles = 0, max_split = 1;
  for (int i = 0; i < b.n; ++i) {
    const UmmaProblem& P = b.p[i];
    UmmaDevProblem& D = dev.p[i];
    if (P.M <= 0 || P.N <= 0 || P.K <= 0 || (P.N % 8)) return IISAN_EINVAL;
    // A
    if (A_MN) IISAN_TRY(make_tensor_map_bf16(&D.map_a, P.A.ptr, P.A.rows, P.A.cols, P.A.pitch, 64, UBK));
    else IISAN_TRY(make_tensor_map_bf16(&D.map_a, P.A.ptr, P.A.rows, P.A.cols, P.A.pitch, UBK, UBM));
    if (B_MN) IISAN_TRY(make_tensor_map_bf16(&D.map_b, P.B.ptr, P.B.rows, P.B.cols, P.B.pitch, 64, UBK));
    else IISAN_TRY(make_tensor_map_bf16(&D.map_b, P.B.ptr, P.B.rows, P.B.cols, P.B.pitch, UBK, BN));
    D.M = P.M; D.N = P.N; D.K = P.K; D.epi = P.epi;
    const int kb_total = (P.K + UBK - 1) / UBK;
    int split = P.splitk < 1 ? 1 : (P.splitk > kb_total ? kb_total : P.splitk);
    D.kblocks_per_split = (kb_total + split - 1) / split;
    split = (kb_total + D.kblocks_per_split - 1) / D.kblocks_per_split;
    if (split > 1 && !(P.epi.atomic && P.epi.out_f32 && !P.epi.out_bf16)) return IISAN_EINVAL;
    const int tiles = ((P.M + UBM - 1) / UBM) * ((P.N + BN - 1) / BN);
    if (tiles > max_tiles) max_tiles = tiles;
    if (split > max_split) max_split = split;
  }
  static bool attr_set = false;
  if (!attr_set) {
    IISAN_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, UmmaSmem<BN>::kTotal));
    attr_set = true;
  }
  dim3 grid(max_tiles, max_split, b.n);
  { LaunchScope ls_(IISAN_K_GEMM, st); umma_gemm_kernel<BN, A_MN, B_MN><<<grid, UTHREADS, UmmaSmem<BN>::kTotal, st>>>(dev); }
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
  if (a_mn != b_mn) return IISAN_EUNSUPPORTED;
  const int bn = maxN <= 64 ? 64 : (maxN <= 128 ? 128 : 256);
  if (!a_mn) {
    if (bn == 64) return launch_cfg<64, false, false>(b, st);
    if (bn == 128) return launch_cfg<128, false, false>(b, st);
    return launch_cfg<256, false, false>(b, st);
  }
  if (bn == 64) return launch_cfg<64, true, true>(b, st);
  if (bn == 128) return launch_cfg<128, true, true>(b, st);
  return launch_cfg<256, true, true>(b, st);
}

}  // namespace iisan
