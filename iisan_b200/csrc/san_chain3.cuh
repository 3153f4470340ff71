// Third-generation fused chain (san_chain3.cu) + the low-rank adjoint path (san_lr.cu): argument blocks and launchers.
//
// What changes against the second generation: the running state x_s of a 128-row tile never leaves the SM.  It is
// resident as packed bf16 in tensor memory (9 chunks of 64 columns) and shared memory (the other 3), updated in place by the
// epilogue warps, and is at once the residual of the next stage and the A operand of its down-projection.  Nothing of width d is
// written to HBM: the forward emits only relu(z_s) [N, 64] per stage and the E outputs of the merged head (fc o pre_fc is ONE
// [E, d] matrix: CC/model/model.py:340-347 applies two Linear layers with nothing in between).  The backward (san_lr.cu) needs no
// width-d stash either: see tests/lowrank_reference.py for the algebra.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "san_chain.cuh"

namespace iisan {

struct Chain3Tower {
  CUtensorMap map_h;        // cached states of this tower's modality as a [N, layers*d] bf16 matrix (mm: image)
  CUtensorMap map_h2;       // mm tower: text states
  CUtensorMap map_wd;       // packed down weights [(A+1)*64, d]  (block A = merged head  W_pre W_fc)
  CUtensorMap map_wu;       // packed up weights   [A*d, 64]
  int mode;                 // 0: intra-modal tower (x = g h + (1-g) last) ; 1: inter-modal (x = last + g h + (1-g) h2)
  int layer[kChainMaxStages], layer2[kChainMaxStages];
  const float* gate[kChainMaxStages];
  const float* b_down[kChainMaxStages + 1];   // [A]: merged head bias  W_pre b_fc + b_pre
  const float* b_up[kChainMaxStages];
  __nv_bfloat16* r_out;     // relu(z_s) stash of this tower [N, A, 64]
  int out_col;              // first column of this tower's E = 64 outputs
};

struct Chain3Args {
  Chain3Tower tower[3];
  float* out;               // [N, out_ld] fp32
  int out_ld;
  int n_items, d, n_stages;
  int one_issuer;           // 1: every tcgen05.mma of a CTA from ONE thread (measurement / fault-hunt switch IISAN_B200_C3_ONE_ISSUER)
};

bool chain3_shape_supported(int d, int emb);
int launch_san_chain3_fwd(const Chain3Args& args, int n_towers, cudaStream_t st);

}  // namespace iisan
