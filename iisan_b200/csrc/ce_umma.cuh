// Fast-mode (tcgen05) in-batch cross-entropy: host entry points used by the C ABI in inbatch_ce.cu.
#pragma once
#include "common.cuh"

namespace iisan {

int ce_fast_splits(int owner_tiles, int stream_tiles);
int ce_fast_supported(const iisan_ce_desc& d);
size_t ce_fast_workspace_bytes(const iisan_ce_desc& d);
int ce_fast_forward(const iisan_ce_desc& d, const float* prec, const float* score, const int64_t* ids_rows, const int64_t* ids_cols,
                    const float* lm_rows, const float* lm_cols, const float* pop, void* ws, float* loss_sum, int32_t* n_valid,
                    float* loss, cudaStream_t st);
int ce_fast_backward(const iisan_ce_desc& d, const float* lm_rows, const float* lm_cols, void* ws, const float* g_sum,
                     const float* g_mean, const int32_t* n_valid, float* d_prec, float* d_score, cudaStream_t st);
int ce_fast_masks(const iisan_ce_desc& d, const int64_t* ids_rows, const int64_t* ids_cols, const float* lm_rows, const float* lm_cols,
                  void* ws, uint8_t* out, cudaStream_t st);

}  // namespace iisan
