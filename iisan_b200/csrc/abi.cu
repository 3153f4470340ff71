// C-ABI plumbing: version/status/error reporting, the dense layer (com_dense) and the cached-state gather.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "linear_tf32.cuh"
#include "launch.cuh"
#include "umma_gemm.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace iisan {
thread_local cudaError_t g_last_cuda_error = cudaSuccess;

// ---- launch accounting / timing ------------------------------------------------------------------------
static std::atomic<int64_t> g_launches[IISAN_K_COUNT];
static std::atomic<int> g_timing_on{0};
struct TimedLaunch { cudaEvent_t beg, end; int kclass; bool used; };
static std::vector<TimedLaunch> g_timed;
static std::mutex g_timed_mu;
constexpr size_t kMaxTimed = 1u << 16;

void launch_scope_begin(int kclass, cudaStream_t st, int* slot) {
  g_launches[kclass].fetch_add(1, std::memory_order_relaxed);
  *slot = -1;
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timed_mu);
  int idx = -1;
  for (size_t i = 0; i < g_timed.size(); ++i) if (!g_timed[i].used) { idx = (int)i; break; }
  if (idx < 0) {
    if (g_timed.size() >= kMaxTimed) return;
    TimedLaunch t{};
    if (cudaEventCreate(&t.beg) != cudaSuccess || cudaEventCreate(&t.end) != cudaSuccess) return;
    g_timed.push_back(t);
    idx = (int)g_timed.size() - 1;
  }
  g_timed[idx].used = true; g_timed[idx].kclass = kclass;
  cudaEventRecord(g_timed[idx].beg, st);
  *slot = idx;
}
void launch_scope_end(int slot, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_timed_mu);
  cudaEventRecord(g_timed[slot].end, st);
}

// ---- cached-state gather -----------------------------------------------------------------------------
// out[i, a, :] = table[ids[i], sel[a], :]  (zeros when ids[i] == 0).  One CTA handles a group of
// (row, selected layer) pairs; every 2*d-byte (or 4*d-byte) layer row is moved with 128-bit loads and
// stores, fully coalesced.  The table may live in HBM or in mapped pinned host memory (zero-copy
// over the host link): the access pattern is identical.
//   restates Build_MM_Dataset.__getitem__ (CC/data_utils/dataset.py:65-92): per-item load of the
//   cached [layers, d] tensor, left padding with zeros, stacking into the batch.
// One WARP per (row, selected layer) pair: the pair's 2*d (4*d) bytes are one contiguous run in the table and in the batch, moved
// as 128-bit words, up to four in flight per lane; the (row, layer) split costs one 32-bit division per pair instead of four
// 64-bit div/mod per 16 bytes.
__global__ void __launch_bounds__(256) gather_states_kernel(const uint4* __restrict__ table, int64_t n_table_items, int layers, int vec_per_row,
                                                            const int64_t* __restrict__ ids, int n, const int* __restrict__ sel, int n_sel,
                                                            uint4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned pairs = (unsigned)n * (unsigned)n_sel;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < pairs; p += warps) {
    const unsigned r = p / (unsigned)n_sel, a = p - r * (unsigned)n_sel;
    const int64_t id = ids[r];
    const bool live = id > 0 && id < n_table_items;
    const uint4* src = table + (id * layers + sel[a]) * (int64_t)vec_per_row;
    uint4* dst = out + (int64_t)p * vec_per_row;
    for (int v0 = lane; v0 < vec_per_row; v0 += 128) {
      uint4 val[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        val[k] = make_uint4(0u, 0u, 0u, 0u);
        if (live && v0 + 32 * k < vec_per_row) val[k] = ld_stream_128(src + v0 + 32 * k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (v0 + 32 * k < vec_per_row) dst[v0 + 32 * k] = val[k];
    }
  }
}

// ---- layer selection + cast to bf16 (fast mode, states stored in fp32 / fp16) ------------------------
// out[i, a, :] = bf16_rn(in[i, sel[a], :]).  One warp per (row, selected layer) pair, eight elements per lane and trip: two
// (fp32) or one (fp16) 128-bit loads, one 128-bit store.  Same rounding as torch's .bfloat16().
template <typename T>
__global__ void __launch_bounds__(256) pack_states_kernel(const T* __restrict__ in, int layers, int d, int n, const int* __restrict__ sel, int n_sel,
                                                          __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned pairs = (unsigned)n * (unsigned)n_sel;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < pairs; p += warps) {
    const unsigned r = p / (unsigned)n_sel, a = p - r * (unsigned)n_sel;
    const T* src = in + ((int64_t)r * layers + sel[a]) * d;
    __nv_bfloat16* dst = out + (int64_t)p * d;
    for (int c = lane * 8; c < d; c += 256) {
      const float4 lo = load4<T>(src + c), hi = load4<T>(src + c + 4);
      uint4 pk;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
      h2[0] = __floats2bfloat162_rn(lo.x, lo.y); h2[1] = __floats2bfloat162_rn(lo.z, lo.w);
      h2[2] = __floats2bfloat162_rn(hi.x, hi.y); h2[3] = __floats2bfloat162_rn(hi.z, hi.w);
      *reinterpret_cast<uint4*>(dst + c) = pk;
    }
  }
}
}  // namespace iisan

using namespace iisan;

extern "C" int iisan_abi_version(void) { return IISAN_ABI_VERSION; }

extern "C" const char* iisan_status_string(int s) {
  switch (s) {
    case IISAN_OK: return "ok";
    case IISAN_EINVAL: return "invalid argument (descriptor, pointer or unsupported shape)";
    case IISAN_ECUDA: return "CUDA runtime error (see iisan_last_cuda_error_string)";
    case IISAN_EWORKSPACE: return "workspace too small";
    case IISAN_EUNSUPPORTED: return "configuration valid in the reference but not built here";
  }
  return "unknown status";
}

extern "C" int64_t iisan_launch_count(int kclass) {
  int64_t t = 0;
  for (int k = 0; k < IISAN_K_COUNT; ++k) if (kclass < 0 || kclass == k) t += g_launches[k].load();
  return t;
}
extern "C" int iisan_timing_enable(int on) { g_timing_on.store(on ? 1 : 0); return IISAN_OK; }
extern "C" int iisan_timing_read(int kclass, double* total_ms, int64_t* launches) {
  if (kclass < 0 || kclass >= IISAN_K_COUNT || !total_ms || !launches) return IISAN_EINVAL;
  std::lock_guard<std::mutex> lk(g_timed_mu);
  double tot = 0.0; int64_t n = 0;
  for (auto& t : g_timed) {
    if (!t.used || t.kclass != kclass) continue;
    IISAN_CUDA_OK(cudaEventSynchronize(t.end));
    float ms = 0.f;
    IISAN_CUDA_OK(cudaEventElapsedTime(&ms, t.beg, t.end));
    tot += ms; ++n; t.used = false;
  }
  *total_ms = tot; *launches = n;
  return IISAN_OK;
}

extern "C" size_t iisan_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(iisan_san_desc);
    case 1: return sizeof(iisan_san_params);
    case 2: return sizeof(iisan_ue_desc);
    case 3: return sizeof(iisan_ue_params);
    case 4: return sizeof(iisan_ce_desc);
    case 5: return sizeof(iisan_adam_tensor);
  }
  return 0;
}

extern "C" int iisan_last_cuda_error(void) { return (int)g_last_cuda_error; }
extern "C" const char* iisan_last_cuda_error_string(void) { return cudaGetErrorString(g_last_cuda_error); }

extern "C" int iisan_linear_forward(int32_t rows, int32_t out_features, int32_t in_features, const float* x, int64_t ldx,
                                    const float* w, const float* b, float* y, int64_t ldy, int32_t compute, iisan_stream_t stream) {
  if (!x || !w || !y || rows <= 0 || out_features <= 0 || in_features <= 0) return IISAN_EINVAL;
  if (compute == IISAN_COMPUTE_BF16 && linear_tf32_supported(rows, out_features, in_features, x, ldx, w, ldy, y))
    return linear_tf32_forward(rows, out_features, in_features, x, ldx, w, b, y, ldy, as_stream(stream));
  GemmBatch g{}; g.n = 1;
  g.p[0] = prob_linear(x, ldx, w, b, y, ldy, rows, out_features, in_features);
  return launch_gemm(g, as_stream(stream));
}

extern "C" int iisan_linear_backward(int32_t rows, int32_t out_features, int32_t in_features, const float* x, int64_t ldx,
                                     const float* w, const float* dy, int64_t lddy, float* dx, int64_t lddx, float* dw, float* db,
                                     int32_t compute, iisan_stream_t stream) {
  if (!x || !w || !dy || rows <= 0 || out_features <= 0 || in_features <= 0) return IISAN_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (compute == IISAN_COMPUTE_BF16 && dw && linear_tf32_supported(rows, out_features, in_features, x, ldx, w, lddy, dy) &&
      (!dx || (lddx % 4 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0)))
    return linear_tf32_backward(rows, out_features, in_features, x, ldx, w, dy, lddy, dx, lddx, dw, db, st);
  if (dx) {
    GemmBatch g{}; g.n = 1;
    g.p[0] = prob_dgrad(dy, lddy, w, dx, lddx, rows, out_features, in_features);
    IISAN_TRY(launch_gemm(g, st));
  }
  if (dw) {
    GemmBatch g{}; g.n = 1;
    g.p[0] = prob_wgrad(dy, lddy, x, ldx, dw, rows, out_features, in_features);
    IISAN_TRY(launch_gemm(g, st));
  }
  if (db) {
    ColsumBatch c{}; c.n = 1;
    c.p[0] = {dy, lddy, rows, out_features, db};
    IISAN_TRY(launch_colsum(c, st));
  }
  return IISAN_OK;
}

extern "C" int iisan_gather_states(const void* table, int32_t dtype, int64_t n_table_items, int32_t layers, int32_t d,
                                   const int64_t* ids, int32_t n, const int32_t* sel, int32_t n_sel, void* out,
                                   iisan_stream_t stream) {
  if (!table || !ids || !sel || !out || n <= 0 || n_sel <= 0 || layers <= 0 || d <= 0 || n_table_items <= 0) return IISAN_EINVAL;
  const int bpe = (int)dtype_size(dtype);
  if ((d * bpe) % 16) return IISAN_EINVAL;
  if ((int64_t)n * n_sel >= ((int64_t)1 << 31)) return IISAN_EINVAL;
  const int64_t pairs = (int64_t)n * n_sel;
  const int blocks = (int)imin64((pairs + 7) / 8, 148 * 8);      // 8 warps per block, 64 warps per SM resident
  cudaStream_t st = as_stream(stream);
  { LaunchScope ls_(IISAN_K_MISC, st); gather_states_kernel<<<blocks, 256, 0, st>>>((const uint4*)table, n_table_items, layers, d * bpe / 16, ids, n, sel, n_sel, (uint4*)out); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

extern "C" int iisan_pack_states(const void* states, int32_t dtype, int64_t n, int32_t layers, int32_t d, const int32_t* sel,
                                 int32_t n_sel, void* out_bf16, iisan_stream_t stream) {
  if (!states || !sel || !out_bf16 || n <= 0 || n_sel <= 0 || layers <= 0 || d <= 0 || (d % 8)) return IISAN_EINVAL;
  if (n * n_sel >= ((int64_t)1 << 31)) return IISAN_EINVAL;
  if ((reinterpret_cast<uintptr_t>(states) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) return IISAN_EINVAL;
  if (dtype != IISAN_F32 && dtype != IISAN_F16 && dtype != IISAN_BF16) return IISAN_EINVAL;
  const int64_t pairs = n * n_sel;
  const int blocks = (int)imin64((pairs + 7) / 8, 148 * 8);
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  LaunchScope ls_(IISAN_K_MISC, st);
  switch (dtype) {
    case IISAN_F32: pack_states_kernel<float><<<blocks, 256, 0, st>>>((const float*)states, layers, d, (int)n, sel, n_sel, out); break;
    case IISAN_F16: pack_states_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)states, layers, d, (int)n, sel, n_sel, out); break;
    case IISAN_BF16: pack_states_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)states, layers, d, (int)n, sel, n_sel, out); break;
    default: break;
  }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

extern "C" int iisan_gemm_bf16(int32_t M, int32_t N, int32_t K, const void* A, int64_t a_pitch, int32_t a_mn_major, const void* B,
                               int64_t b_pitch, int32_t b_mn_major, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                               const float* bias, int32_t relu, int32_t splitk, iisan_stream_t stream) {
  if (!A || !B || (!out_f32 && !out_bf16) || M <= 0 || N <= 0 || K <= 0) return IISAN_EINVAL;
  UmmaBatch b{}; b.n = 1;
  UmmaProblem& P = b.p[0];
  P.A.ptr = (const __nv_bfloat16*)A; P.A.pitch = a_pitch;
  if (a_mn_major) { P.A.rows = K; P.A.cols = M; } else { P.A.rows = M; P.A.cols = K; }
  P.B.ptr = (const __nv_bfloat16*)B; P.B.pitch = b_pitch;
  if (b_mn_major) { P.B.rows = K; P.B.cols = N; } else { P.B.rows = N; P.B.cols = K; }
  P.a_mn_major = a_mn_major; P.b_mn_major = b_mn_major;
  P.M = M; P.N = N; P.K = K; P.splitk = splitk < 1 ? 1 : splitk;
  P.epi.out_f32 = out_f32; P.epi.ld_f32 = ld_f32;
  P.epi.out_bf16 = (__nv_bfloat16*)out_bf16; P.epi.ld_bf16 = ld_bf16;
  P.epi.bias = bias; P.epi.relu = relu; P.epi.atomic = P.splitk > 1 ? 1 : 0;
  return launch_umma_gemm(b, as_stream(stream));
}

extern "C" int iisan_stage_states_h2d(const void* host_src, void* dev_dst, int64_t n_rows, int32_t layers, int32_t d, int32_t dtype,
                                      const int32_t* sel, int32_t n_sel, iisan_stream_t stream) {
  if (!host_src || !dev_dst || !sel || n_rows <= 0 || layers <= 0 || d <= 0 || n_sel <= 0 || n_sel > layers) return IISAN_EINVAL;
  const size_t bpe = dtype_size(dtype);
  const size_t pitch = (size_t)layers * d * bpe;
  cudaStream_t st = as_stream(stream);
  int i = 0;
  while (i < n_sel) {
    if (sel[i] < 0 || sel[i] >= layers || (i > 0 && sel[i] <= sel[i - 1])) return IISAN_EINVAL;
    int j = i;
    while (j + 1 < n_sel && sel[j + 1] == sel[j] + 1) ++j;            // run of adjacent layers -> one DMA
    const size_t off = (size_t)sel[i] * d * bpe, width = (size_t)(j - i + 1) * d * bpe;
    IISAN_CUDA_OK(cudaMemcpy2DAsync((char*)dev_dst + off, pitch, (const char*)host_src + off, pitch, width, (size_t)n_rows, cudaMemcpyHostToDevice, st));
    i = j + 1;
  }
  return IISAN_OK;
}

namespace iisan { int set_chain_generation(int gen); }
extern "C" int iisan_debug_chain_generation(int32_t gen) { return iisan::set_chain_generation(gen); }
