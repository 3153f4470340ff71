// Shared device/host helpers for the iisan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/iisan_b200.h"

namespace iisan {

// ---- error plumbing (no exceptions across the C ABI) ------------------------------------------
extern thread_local cudaError_t g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = e;
  return IISAN_ECUDA;
}

#define IISAN_CUDA_OK(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::iisan::cuda_fail(_e);     \
  } while (0)

#define IISAN_LAUNCH_OK()                                     \
  do {                                                        \
    cudaError_t _e = cudaPeekAtLastError();                   \
    if (_e != cudaSuccess) return ::iisan::cuda_fail(_e);     \
  } while (0)

#define IISAN_TRY(expr)                                       \
  do {                                                        \
    int _s = (expr);                                          \
    if (_s != IISAN_OK) return _s;                            \
  } while (0)

inline cudaStream_t as_stream(iisan_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute applies to the CURRENT device only.  Each launcher keeps one std::atomic<uint64_t> of the devices it has
// configured (one process per GPU is the normal deployment, but one process driving several GPUs must work too).
inline uint64_t device_bit() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return 1ull << (dev & 63);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-owned workspace (256 B aligned slices).
struct Arena {
  char* base;
  size_t off;
  explicit Arena(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T), 256);
    return p;
  }
};

__host__ __device__ inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

inline size_t dtype_size(int dt) { return dt == IISAN_F32 ? 4 : 2; }

// ---- device helpers -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// exact GELU, torch's nn.GELU() default (CC/model/modules.py:104-105):  x Phi(x)  and its derivative  Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * expf(-0.5f * x * x);
}

// 128-bit read-only streaming load (L1 no-allocate): hidden states are read once per pass.
__device__ __forceinline__ uint4 ld_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Load 4 consecutive elements of a hidden-state row as fp32 (p must be 4-element aligned).
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  uint4 r = ld_stream_128(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&r.x);
  __half2 b = *reinterpret_cast<__half2*>(&r.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sigmoid(p / 0.1): CC/model/model.py:321.  Division first, like the reference.
__device__ __forceinline__ float gate_value(const float* p) {
  float x = __ldg(p) / 0.1f;
  return 1.0f / (1.0f + expf(-x));
}

}  // namespace iisan
