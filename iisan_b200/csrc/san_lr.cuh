// Third-generation fast path of the side-adapter network (san_lr.cu): resident-state chain forward + low-rank adjoint backward.
#pragma once
#include "common.cuh"

namespace iisan {

bool san_lr_eligible(const iisan_san_desc& D);                          // shape / plan conditions
bool san_lr_usable(const iisan_san_desc& D, const iisan_san_params& P); // + pointer alignment of the parameters (bulk copies, 128-bit loads)
size_t san_lr_workspace_bytes(const iisan_san_desc& D);
int san_lr_forward(const iisan_san_desc* D, const iisan_san_params* P, const void* image, const void* text, void* lr_ws, float* out,
                   cudaStream_t st);
int san_lr_backward(const iisan_san_desc* D, const iisan_san_params* P, const iisan_san_params* G, const void* image, const void* text,
                    void* lr_ws, const float* d_out, cudaStream_t st);

}  // namespace iisan
