// Launch accounting + optional per-kernel-class CUDA-event timing (bench.py's live roofline numbers).
// Disabled by default: a LaunchScope then costs one relaxed atomic increment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iisan {

void launch_scope_begin(int kclass, cudaStream_t st, int* slot);
void launch_scope_end(int slot, cudaStream_t st);

struct LaunchScope {
  int slot;
  cudaStream_t st;
  LaunchScope(int kclass, cudaStream_t s) : slot(-1), st(s) { launch_scope_begin(kclass, s, &slot); }
  ~LaunchScope() { if (slot >= 0) launch_scope_end(slot, st); }
};

}  // namespace iisan
