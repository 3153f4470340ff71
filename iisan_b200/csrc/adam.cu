// Fused multi-tensor Adam for the hot path's 146 small parameter tensors (CC/run.py:260-307 builds 5 name-routed learning-rate
// groups and steps torch.optim.Adam over them; :383-385).  One launch updates up to IISAN_ADAM_MAX_TENSORS tensors (one CTA per 4096-element chunk of the flat chunk list), each with
// its own learning rate; the step counter lives on the device so that the launch can be captured in a CUDA graph.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)      (torch.optim.Adam defaults,
//   weight_decay = 0, amsgrad off -- the reference passes only lr).
#include "common.cuh"
#include "launch.cuh"

namespace iisan {

constexpr int kAdamChunk = 4096;        // elements per CTA

struct AdamArgs {
  iisan_adam_tensor t[IISAN_ADAM_MAX_TENSORS];
  int chunk_start[IISAN_ADAM_MAX_TENSORS + 1];   // CTA b works on tensor i with chunk_start[i] <= b < chunk_start[i+1]
  int n;
  float beta1, beta2, eps;
  const float* step;   // device scalar: step count t (already incremented for this step)
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  // which tensor: binary search over the chunk prefix (<= 8 probes, warp-uniform)
  int lo = 0, hi = a.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.chunk_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const iisan_adam_tensor& T = a.t[lo];
  const int64_t beg = (int64_t)((int)blockIdx.x - a.chunk_start[lo]) * kAdamChunk;
  const int64_t end = imin64(T.numel, beg + kAdamChunk);
  const float t = __ldg(a.step);
  const float bc1 = 1.0f - powf(a.beta1, t), bc2 = 1.0f - powf(a.beta2, t);
  const float step_size = T.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;
  auto upd = [&](float g, float& m, float& v, float& p) {
    m = b1 * m + (1.0f - b1) * g;
    v = b2 * v + (1.0f - b2) * g * g;
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
  };
  const bool vec = (((reinterpret_cast<uintptr_t>(T.param) | reinterpret_cast<uintptr_t>(T.grad) | reinterpret_cast<uintptr_t>(T.exp_avg) |
                      reinterpret_cast<uintptr_t>(T.exp_avg_sq)) & 15) == 0);
  int64_t i = beg;
  if (vec) {                                   // chunk starts are multiples of 4096: 128-bit accesses
    const int64_t n4 = (end - beg) >> 2;
    for (int64_t q = threadIdx.x; q < n4; q += blockDim.x) {
      const int64_t j = (beg >> 2) + q;
      const float4 g = __ldg(reinterpret_cast<const float4*>(T.grad) + j);
      float4 m = reinterpret_cast<float4*>(T.exp_avg)[j], v = reinterpret_cast<float4*>(T.exp_avg_sq)[j], p = reinterpret_cast<float4*>(T.param)[j];
      upd(g.x, m.x, v.x, p.x); upd(g.y, m.y, v.y, p.y); upd(g.z, m.z, v.z, p.z); upd(g.w, m.w, v.w, p.w);
      reinterpret_cast<float4*>(T.exp_avg)[j] = m; reinterpret_cast<float4*>(T.exp_avg_sq)[j] = v; reinterpret_cast<float4*>(T.param)[j] = p;
    }
    i = beg + (n4 << 2);
  }
  for (int64_t j = i + threadIdx.x; j < end; j += blockDim.x) {
    float m = T.exp_avg[j], v = T.exp_avg_sq[j], p = T.param[j];
    upd(T.grad[j], m, v, p);
    T.exp_avg[j] = m; T.exp_avg_sq[j] = v; T.param[j] = p;
  }
}

__global__ void adam_tick_kernel(float* step) { *step += 1.0f; }

}  // namespace iisan

using namespace iisan;

extern "C" int iisan_adam_step(const iisan_adam_tensor* tensors, int32_t n, float beta1, float beta2, float eps, float* step_dev,
                               int32_t advance_step, iisan_stream_t stream) {
  if (!tensors || n <= 0 || !step_dev) return IISAN_EINVAL;
  cudaStream_t st = as_stream(stream);
  if (advance_step) {
    { LaunchScope ls_(IISAN_K_MISC, st); adam_tick_kernel<<<1, 1, 0, st>>>(step_dev); }
    IISAN_LAUNCH_OK();
  }
  for (int base = 0; base < n; base += IISAN_ADAM_MAX_TENSORS) {
    AdamArgs a;
    a.n = n - base < IISAN_ADAM_MAX_TENSORS ? n - base : IISAN_ADAM_MAX_TENSORS;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.step = step_dev;
    int64_t chunks = 0;
    for (int i = 0; i < a.n; ++i) {
      a.t[i] = tensors[base + i];
      if (!a.t[i].param || !a.t[i].grad || !a.t[i].exp_avg || !a.t[i].exp_avg_sq || a.t[i].numel <= 0) return IISAN_EINVAL;
      a.chunk_start[i] = (int)chunks;
      chunks += (a.t[i].numel + kAdamChunk - 1) / kAdamChunk;
      if (chunks > 0x7fffffff) return IISAN_EINVAL;
    }
    a.chunk_start[a.n] = (int)chunks;
    { LaunchScope ls_(IISAN_K_MISC, st); adam_kernel<<<(unsigned)chunks, 256, 0, st>>>(a); }
    IISAN_LAUNCH_OK();
  }
  return IISAN_OK;
}
