// Fused multi-tensor Adam for the hot path's 146 small parameter tensors (CC/run.py:260-307 builds 5 name-routed learning-rate
// groups and steps torch.optim.Adam over them; :383-385).  One launch updates up to IISAN_ADAM_MAX_TENSORS tensors, each with
// its own learning rate; the step counter lives on the device so that the launch can be captured in a CUDA graph.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)      (torch.optim.Adam defaults,
//   weight_decay = 0, amsgrad off -- the reference passes only lr).
#include "common.cuh"
#include "launch.cuh"

namespace iisan {

constexpr int kAdamChunk = 4096;

struct AdamArgs {
  iisan_adam_tensor t[IISAN_ADAM_MAX_TENSORS];
  int n;
  float beta1, beta2, eps;
  const float* step;   // device scalar: step count t (already incremented for this step)
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  const iisan_adam_tensor& T = a.t[blockIdx.y];
  const int64_t beg = (int64_t)blockIdx.x * kAdamChunk;
  if (beg >= T.numel) return;
  const int64_t end = imin64(T.numel, beg + kAdamChunk);
  const float t = __ldg(a.step);
  const float bc1 = 1.0f - powf(a.beta1, t), bc2 = 1.0f - powf(a.beta2, t);
  const float step_size = T.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float b1 = a.beta1, b2 = a.beta2, eps = a.eps;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const float g = T.grad[i];
    const float m = b1 * T.exp_avg[i] + (1.0f - b1) * g;
    const float v = b2 * T.exp_avg_sq[i] + (1.0f - b2) * g * g;
    T.exp_avg[i] = m; T.exp_avg_sq[i] = v;
    T.param[i] -= step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + eps));
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
    int64_t mx = 1;
    for (int i = 0; i < a.n; ++i) {
      a.t[i] = tensors[base + i];
      if (!a.t[i].param || !a.t[i].grad || !a.t[i].exp_avg || !a.t[i].exp_avg_sq || a.t[i].numel <= 0) return IISAN_EINVAL;
      mx = imax64(mx, (a.t[i].numel + kAdamChunk - 1) / kAdamChunk);
    }
    { LaunchScope ls_(IISAN_K_MISC, st); adam_kernel<<<dim3((unsigned)mx, a.n), 256, 0, st>>>(a); }
    IISAN_LAUNCH_OK();
  }
  return IISAN_OK;
}
