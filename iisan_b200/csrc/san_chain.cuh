// Fused side-adapter chain kernels (san_chain.cu): argument blocks and launchers.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace iisan {

constexpr int kChainMaxStages = 8;

struct ChainTower {
  CUtensorMap map_h;        // cached states of this tower's modality as a [N, layers*d] bf16 matrix (mm: image)
  CUtensorMap map_h2;       // mm tower: text states
  CUtensorMap map_wd;       // packed down weights [A*r, d]
  CUtensorMap map_wu;       // packed up weights   [A*d, r]
  int mode;                 // 0: intra-modal tower (x = g h + (1-g) last) ; 1: inter-modal (x = last + g h + (1-g) h2)
  int layer[kChainMaxStages], layer2[kChainMaxStages];
  const float* gate[kChainMaxStages];
  const float* b_down[kChainMaxStages];
  const float* b_up[kChainMaxStages];
  __nv_bfloat16* x_stash[kChainMaxStages];     // [N, d]
  __nv_bfloat16* z_stash[kChainMaxStages];     // [N, r]
  __nv_bfloat16* last_stash[kChainMaxStages];  // [N, d] or null (the final stage's must be set: it feeds the heads)
};

struct ChainArgs {
  ChainTower tower[3];
  int n_items, d, n_stages;
};

struct ChainBwdTower {
  CUtensorMap map_h;        // cached states [N, layers*d] (mm: image)
  CUtensorMap map_aux;      // intra: last stash of all stages [A*N, d] ; mm: text states [N, layers*d]
  CUtensorMap map_dy;       // d last_s of all stages [A*N, d] (stage A-1 filled by the head gradient GEMM, the rest by this kernel)
  CUtensorMap map_wd, map_wu;
  int mode;
  int layer[kChainMaxStages], layer2[kChainMaxStages];
  const float* gate[kChainMaxStages];
  float* g_gate[kChainMaxStages];                 // gradients (accumulated)
  float* g_b_down[kChainMaxStages];
  float* g_b_up[kChainMaxStages];
  const __nv_bfloat16* z_stash[kChainMaxStages];  // [N, r]
  __nv_bfloat16* dz_stash[kChainMaxStages];       // [N, r]   out: wgrad operand
  __nv_bfloat16* dy_stash;                        // [A, N, d] in/out
};

struct ChainBwdArgs {
  ChainBwdTower tower[3];
  int n_items, d, n_stages;
};

int chain_fill_bwd_tower(ChainBwdTower* T, int mode, const void* h, int64_t h_rows, int64_t h_pitch_cols, const void* h2, int64_t h2_pitch_cols,
                         const __nv_bfloat16* wd_pack, const __nv_bfloat16* wu_pack, const __nv_bfloat16* dy_all,
                         const __nv_bfloat16* last_all, int n_stages, int d);
int launch_san_chain_bwd(const ChainBwdArgs& args, int n_towers, cudaStream_t st);

int chain_fill_tower(ChainTower* T, int mode, const void* h, int64_t h_rows, int64_t h_pitch_cols, const void* h2, int64_t h2_pitch_cols,
                     const __nv_bfloat16* wd_pack, const __nv_bfloat16* wu_pack, int n_stages, int d);
int launch_san_chain_fwd(const ChainArgs& args, int n_towers, cudaStream_t st);

}  // namespace iisan
