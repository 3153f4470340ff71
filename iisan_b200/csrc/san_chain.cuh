// Fused side-adapter chain kernels (san_chain.cu): argument blocks and launchers.
// Stash matrices of a tower are contiguous over the stages with the item count padded to the 128-row tile:
// x / last / dy : [A * n_pad, d] ; z / dz : [A * n_pad, 64]   (stage s starts at row s * n_pad).
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
  CUtensorMap map_x;        // x stash (stored, and loaded back one stage later as the residual)
  CUtensorMap map_last;     // last stash (stored: every stage when store_last, else only the final stage, which feeds the heads)
  CUtensorMap map_z;        // z stash (stored)
  int mode;                 // 0: intra-modal tower (x = g h + (1-g) last) ; 1: inter-modal (x = last + g h + (1-g) h2)
  int store_last;
  int layer[kChainMaxStages], layer2[kChainMaxStages];
  const float* gate[kChainMaxStages];
  const float* b_down[kChainMaxStages];
  const float* b_up[kChainMaxStages];
  __nv_bfloat16* z_out;     // z stash as a plain pointer (second-generation kernel: direct stores)
};

struct ChainArgs {
  ChainTower tower[3];
  int n_items, n_pad, d, n_stages;
  int pf;                   // second generation: hidden-state tiles are requested into L2 this many chunks ahead (0: off)
  int rows;                 // second generation: rows per tile (TMA box rows), <= 128
};

struct ChainBwdTower {
  CUtensorMap map_h;        // cached states [N, layers*d] (mm: image)
  CUtensorMap map_aux;      // intra: last stash ; mm: text states [N, layers*d]
  CUtensorMap map_dy;       // d last_s of all stages (stage A-1 filled by the head gradient GEMM, the rest stored by the kernel)
  CUtensorMap map_dz;       // dz stash (stored: weight-gradient operand)
  CUtensorMap map_wd, map_wu;
  int mode;
  int layer[kChainMaxStages], layer2[kChainMaxStages];
  const float* gate[kChainMaxStages];
  float* g_gate[kChainMaxStages];                 // gradients (accumulated)
  float* g_b_down[kChainMaxStages];
  float* g_b_up[kChainMaxStages];
  const __nv_bfloat16* z_stash;                   // [A * n_pad, r]
  __nv_bfloat16* dz_out;                          // dz stash as a plain pointer (second-generation kernel: direct stores)
};

struct ChainBwdArgs {
  ChainBwdTower tower[3];
  int n_items, n_pad, d, n_stages;
  int pf;
  int rows;
};

int chain_n_pad(int n_items);
int chain_tile_rows(int n_items);          // row tile of the second-generation kernels (the first generation: 128)
int chain_fill_tower(ChainTower* T, int mode, const void* h, int64_t n_items, int64_t h_pitch_cols, const void* h2, int64_t h2_pitch_cols,
                     const __nv_bfloat16* wd_pack, const __nv_bfloat16* wu_pack, const __nv_bfloat16* x_all, const __nv_bfloat16* last_all,
                     const __nv_bfloat16* z_all, int n_stages, int d, int box_rows = 128);
int chain_fill_bwd_tower(ChainBwdTower* T, int mode, const void* h, int64_t n_items, int64_t h_pitch_cols, const void* h2,
                         int64_t h2_pitch_cols, const __nv_bfloat16* wd_pack, const __nv_bfloat16* wu_pack, const __nv_bfloat16* dy_all,
                         const __nv_bfloat16* last_all, const __nv_bfloat16* dz_all, int n_stages, int d, int box_rows = 128);
int launch_san_chain_fwd(const ChainArgs& args, int n_towers, cudaStream_t st);
bool chain2_shape_supported(int d);                                                  // widths the second generation covers
int launch_san_chain2_fwd(const ChainArgs& args, int n_towers, cudaStream_t st);   // san_chain2.cu
int launch_san_chain2_bwd(const ChainBwdArgs& args, int n_towers, cudaStream_t st);   // san_chain2_bwd.cu
int launch_san_chain_bwd(const ChainBwdArgs& args, int n_towers, cudaStream_t st);

}  // namespace iisan
