// Elementwise stage kernels of the side-adapter network: layer-select gather + gated fusion (forward) and the
// gate-gradient / state-gradient pass (backward).  Shared by the exact (san.cu) and the bf16 (san_bf16.cu) paths.
//   gated fusion       CC/model/model.py:319-326   x = g*h_l + (1-g)*last,  g = sigmoid(p/0.1)
//   inter-modal mix    CC/model/model.py:335-337   x = last + g*h_cv + (1-g)*h_text
// The cached states are read in place from the caller's [N, layers, d] tensors (row pitch layers*d, 128-bit
// loads of the selected layer only); nothing else of the 13 layers is touched.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "launch.cuh"

namespace iisan {

struct MixSrc {
  const void* p;       // null => zeros
  int64_t row_stride;  // elements between consecutive rows
  int is_state;        // 1: cached hidden state of dtype T; 0: dense fp32; 2: dense bf16
};

template <typename T>
__device__ __forceinline__ float4 mix_load(const MixSrc& s, int64_t row, int col) {
  if (s.p == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.is_state == 1) return load4<T>(reinterpret_cast<const T*>(s.p) + row * s.row_stride + col);
  if (s.is_state == 2) return load4<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(s.p) + row * s.row_stride + col);
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(s.p) + row * s.row_stride + col);
}

__device__ __forceinline__ void store4_bf16(__nv_bfloat16* p, const float4& v) {
  uint2 q;
  *reinterpret_cast<__nv_bfloat162*>(&q.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(&q.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = q;
}

template <typename T>
static MixSrc state_src(const void* base, int layers, int d, int layer) {
  MixSrc s;
  s.p = reinterpret_cast<const T*>(base) + (int64_t)layer * d;
  s.row_stride = (int64_t)layers * d;
  s.is_state = 1;
  return s;
}
static inline MixSrc dense_src(const float* p, int64_t ld) {
  MixSrc s; s.p = p; s.row_stride = ld; s.is_state = 0; return s;
}
static inline MixSrc dense_bf16_src(const __nv_bfloat16* p, int64_t ld) {
  MixSrc s; s.p = p; s.row_stride = ld; s.is_state = 2; return s;
}

struct MixProb {
  MixSrc P, Q, R;
  const float* gate;  // device pointer to the [1] gate parameter
  int mode;           // 0: x = g*P + (1-g)*R ; 1: x = R + g*P + (1-g)*Q ; 2: x = P (layer gather / cast)
  float* X;           // [N, d] dense fp32, or null
  __nv_bfloat16* Xb;  // [N, d] dense bf16 (GEMM operand of the fast mode), or null
  int N, d;
};
struct MixBatch { MixProb p[kMaxProbs]; int n; };

template <typename T>
__global__ void __launch_bounds__(256) mix_kernel(const MixBatch batch) {
  const MixProb& M = batch.p[blockIdx.y];
  const int d4 = M.d / 4;
  const int64_t total = (int64_t)M.N * d4;
  float g = 0.f;
  if (M.mode != 2) g = gate_value(M.gate);
  const float omg = 1.0f - g;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const int col = (int)(i % d4) * 4;
    const float4 p = mix_load<T>(M.P, row, col);
    float4 x;
    if (M.mode == 0) {
      const float4 r = mix_load<T>(M.R, row, col);
      // exact reference order, no fma contraction: (g*h) + ((1-g)*last)
      x.x = __fadd_rn(__fmul_rn(g, p.x), __fmul_rn(omg, r.x));
      x.y = __fadd_rn(__fmul_rn(g, p.y), __fmul_rn(omg, r.y));
      x.z = __fadd_rn(__fmul_rn(g, p.z), __fmul_rn(omg, r.z));
      x.w = __fadd_rn(__fmul_rn(g, p.w), __fmul_rn(omg, r.w));
    } else if (M.mode == 1) {
      const float4 q = mix_load<T>(M.Q, row, col);
      const float4 r = mix_load<T>(M.R, row, col);
      // (last + g*h_cv) + (1-g)*h_text
      x.x = __fadd_rn(__fadd_rn(r.x, __fmul_rn(g, p.x)), __fmul_rn(omg, q.x));
      x.y = __fadd_rn(__fadd_rn(r.y, __fmul_rn(g, p.y)), __fmul_rn(omg, q.y));
      x.z = __fadd_rn(__fadd_rn(r.z, __fmul_rn(g, p.z)), __fmul_rn(omg, q.z));
      x.w = __fadd_rn(__fadd_rn(r.w, __fmul_rn(g, p.w)), __fmul_rn(omg, q.w));
    } else {
      x = p;
    }
    if (M.X) *reinterpret_cast<float4*>(M.X + row * M.d + col) = x;
    if (M.Xb) store4_bf16(M.Xb + row * M.d + col, x);
  }
}

// Backward of the fusion: given dx [N,d]
//   mode 0: dgate += sum dx*(P - R) * g(1-g)/0.1 ; dR = (1-g)*dx            (written to dPrev if non-null)
//   mode 1: dgate += sum dx*(P - Q) * g(1-g)/0.1 ; dR = dx (caller aliases) ;
//           dP_out = g*dx (if non-null), dQ_out = (1-g)*dx (if non-null)      (down_project inputs)
// Every fp32 output has an optional bf16 twin (the GEMM operand of the fast mode).
struct MixBwdProb {
  MixSrc P, Q, R;
  const float* gate;
  float* dgate;
  int mode;
  const float* dX;   // [N,d]
  float* dPrev;      // mode 0: [N,d] or null
  float* dP_out;     // mode 1: [N,d] or null
  float* dQ_out;     // mode 1: [N,d] or null
  __nv_bfloat16* dPrevb;
  __nv_bfloat16* dP_outb;
  __nv_bfloat16* dQ_outb;
  int N, d;
};
struct MixBwdBatch { MixBwdProb p[kMaxProbs]; int n; };

template <typename T>
__global__ void __launch_bounds__(256) mix_bwd_kernel(const MixBwdBatch batch) {
  const MixBwdProb& M = batch.p[blockIdx.y];
  const int d4 = M.d / 4;
  const int64_t total = (int64_t)M.N * d4;
  const float g = gate_value(M.gate);
  const float omg = 1.0f - g;
  double part = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d4;
    const int col = (int)(i % d4) * 4;
    const float4 dx = *reinterpret_cast<const float4*>(M.dX + row * M.d + col);
    const float4 p = mix_load<T>(M.P, row, col);
    const float4 o = (M.mode == 0) ? mix_load<T>(M.R, row, col) : mix_load<T>(M.Q, row, col);
    float s = dx.x * (p.x - o.x);
    s = fmaf(dx.y, p.y - o.y, s);
    s = fmaf(dx.z, p.z - o.z, s);
    s = fmaf(dx.w, p.w - o.w, s);
    part += (double)s;
    const int64_t off = row * M.d + col;
    if (M.mode == 0) {
      const float4 v = make_float4(omg * dx.x, omg * dx.y, omg * dx.z, omg * dx.w);
      if (M.dPrev) *reinterpret_cast<float4*>(M.dPrev + off) = v;
      if (M.dPrevb) store4_bf16(M.dPrevb + off, v);
    } else {
      if (M.dP_out || M.dP_outb) {
        const float4 v = make_float4(g * dx.x, g * dx.y, g * dx.z, g * dx.w);
        if (M.dP_out) *reinterpret_cast<float4*>(M.dP_out + off) = v;
        if (M.dP_outb) store4_bf16(M.dP_outb + off, v);
      }
      if (M.dQ_out || M.dQ_outb) {
        const float4 v = make_float4(omg * dx.x, omg * dx.y, omg * dx.z, omg * dx.w);
        if (M.dQ_out) *reinterpret_cast<float4*>(M.dQ_out + off) = v;
        if (M.dQ_outb) store4_bf16(M.dQ_outb + off, v);
      }
    }
  }
  // block reduction (double) -> one atomic per CTA
  __shared__ double red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    // d sigmoid(p/0.1)/dp = g(1-g)/0.1
    atomicAdd(M.dgate, (float)(t * (double)g * (double)omg / 0.1));
  }
}

template <typename T>
static int launch_mix(const MixBatch& b, cudaStream_t st) {
  if (b.n == 0) return IISAN_OK;
  int64_t mx = 0;
  for (int i = 0; i < b.n; ++i) mx = max(mx, (int64_t)b.p[i].N * (b.p[i].d / 4));
  int blocks = (int)imin64((mx + 255) / 256, 148 * 8);
  { LaunchScope ls_(IISAN_K_STREAM, st); mix_kernel<T><<<dim3(blocks, b.n), 256, 0, st>>>(b); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}
template <typename T>
static int launch_mix_bwd(const MixBwdBatch& b, cudaStream_t st) {
  if (b.n == 0) return IISAN_OK;
  int64_t mx = 0;
  for (int i = 0; i < b.n; ++i) mx = max(mx, (int64_t)b.p[i].N * (b.p[i].d / 4));
  int blocks = (int)imin64((mx + 255) / 256, 148 * 4);
  { LaunchScope ls_(IISAN_K_STREAM, st); mix_bwd_kernel<T><<<dim3(blocks, b.n), 256, 0, st>>>(b); }
  IISAN_LAUNCH_OK();
  return IISAN_OK;
}

}  // namespace iisan
