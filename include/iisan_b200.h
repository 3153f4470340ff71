/*
 * iisan_b200 -- C ABI of the B200-native IISAN(Cached) / IISAN-Versa training hot path.
 *
 * The reference (GAIR-Lab/IISAN) is pure Python/PyTorch and has no FFI of its own; the interfaces
 * these entry points replace are the PyTorch module calls on its train step.  Each entry point
 * cites the reference lines it stands in for (paths relative to the reference root; CC =
 * Code_Cached, CA = Code_Cached_Asym).  The Python mirror of the reference `model` package
 * (iisan_b200/model) binds them through ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, POD descriptors, a CUDA stream handle.  No torch types.
 *   - every call is stream-ordered and asynchronous, never synchronises the device, never
 *     allocates: scratch and the forward->backward stash live in a caller-owned workspace whose
 *     size comes from the matching *_workspace_bytes() query.
 *   - return value: 0 (IISAN_OK) or an iisan_status code; no exceptions cross the ABI.
 *   - parameters are the reference's fp32 nn.Linear tensors ([out, in] row-major) passed by
 *     pointer; gradients are ACCUMULATED (+=) into caller-zeroed fp32 buffers of the same shapes.
 *   - sm_100a only.  There is no CPU fallback.
 */
#ifndef IISAN_B200_H_
#define IISAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IISAN_ABI_VERSION 3
#define IISAN_MAX_STAGES 64
#define IISAN_MAX_BLOCKS 8

typedef struct CUstream_st* iisan_stream_t; /* == cudaStream_t */

enum iisan_status {
  IISAN_OK = 0,
  IISAN_EINVAL = 1,       /* bad descriptor / null pointer / unsupported shape */
  IISAN_ECUDA = 2,        /* a CUDA runtime call failed: see iisan_last_cuda_error() */
  IISAN_EWORKSPACE = 3,   /* workspace too small */
  IISAN_EUNSUPPORTED = 4  /* valid in the reference, not built here (e.g. GELU adapters) */
};

enum iisan_dtype { IISAN_F32 = 0, IISAN_BF16 = 1, IISAN_F16 = 2 };

/* arithmetic mode: FP32 = fp32 FMA everywhere (<=1e-5 rel vs the reference run in fp32);
 * BF16 = the fast mode (<=1e-2 rel on loss / embeddings): bf16 tensor-core (tcgen05) operands for the side-adapter network and
 * the in-batch loss, TF32 tensor-core (mma.sync) operands -- the mantissa of the reference's fp16 autocast GEMMs -- for the
 * SASRec user encoder (emb 64, seq_len 10) and the dense layer; fp32 accumulation, LayerNorm, softmax, log-sum-exp. */
enum iisan_compute { IISAN_COMPUTE_FP32 = 0, IISAN_COMPUTE_BF16 = 1 };

int iisan_abi_version(void);
const char* iisan_status_string(int status);
/* last cudaError_t observed by this library on the calling thread (0 = none) and its message */
int iisan_last_cuda_error(void);
const char* iisan_last_cuda_error_string(void);
/* sizeof() of the POD structs as compiled, so that a foreign-language binding can verify its mirror:
 * which = 0 iisan_san_desc, 1 iisan_san_params, 2 iisan_ue_desc, 3 iisan_ue_params, 4 iisan_ce_desc, 5 iisan_adam_tensor. */
size_t iisan_sizeof(int which);

/* Launch accounting and live kernel timing (used by bench.py for the roofline numbers).
 * Every kernel launch of the library belongs to one class. */
enum iisan_kernel_class {
  IISAN_K_STREAM = 0, /* hidden-state streaming: layer-select gather + gate fusion (fwd) and its backward */
  IISAN_K_GEMM = 1,   /* adapter / head / dense-layer GEMMs */
  IISAN_K_USER = 2,   /* SASRec layer-norm / attention kernels */
  IISAN_K_CE = 3,     /* fused in-batch cross-entropy */
  IISAN_K_MISC = 4,   /* reductions, gathers, optimizer */
  IISAN_K_CHAIN = 5,  /* fused tcgen05 adapter-chain kernel, forward */
  IISAN_K_CHAIN_BWD = 6, /* fused tcgen05 adapter-chain kernel, backward (third generation: the rank-space chain) */
  IISAN_K_WGRAD_STREAM = 7, /* third generation: the backward's one GEMM pass over the cached hidden states (G = dz^T h) + Gram blocks */
  IISAN_K_COUNT = 8
};
/* total kernels launched by this process through the library, per class (kclass < 0: all classes) */
int64_t iisan_launch_count(int kclass);
/* on != 0: bracket every launch with CUDA events on its stream (pool of 1<<16 pairs, then stops recording) */
int iisan_timing_enable(int on);
/* synchronise the recorded events, return summed device milliseconds and launches of `kclass` since the last
 * read, and recycle the pool entries of that class */
int iisan_timing_read(int kclass, double* total_ms, int64_t* launches);

/* ---------------------------------------------------------------------------------------------
 * Side-adapter network  (CC/model/model.py:257-349  IISANAdaptedMModel ;
 *                        CA/model/model.py:257-429  IISAN-Versa: group layer-drop + dim alignment ;
 *                        CC/model/modules.py:98-116 AdapterBlock)
 * ------------------------------------------------------------------------------------------- */
typedef struct iisan_adapter_ptrs {
  float* w_down; /* [r, d] */
  float* b_down; /* [r]    */
  float* w_up;   /* [d, r] */
  float* b_up;   /* [d]    */
} iisan_adapter_ptrs;

typedef struct iisan_linear_ptrs {
  float* w; /* [out, in] */
  float* b; /* [out]     */
} iisan_linear_ptrs;

typedef struct iisan_san_params {
  iisan_adapter_ptrs text[IISAN_MAX_STAGES];        /* bert_adapter_list */
  iisan_adapter_ptrs img[IISAN_MAX_STAGES];         /* cv_adapter_list   */
  iisan_adapter_ptrs mm[IISAN_MAX_STAGES];          /* mm_adapter_list   */
  iisan_linear_ptrs down_project[IISAN_MAX_STAGES]; /* CA down_project_list (wide -> narrow) */
  float* gate_text[IISAN_MAX_STAGES];               /* side_gate_params_text[i], shape [1] */
  float* gate_img[IISAN_MAX_STAGES];                /* side_gate_params_cv   */
  float* gate_mm[IISAN_MAX_STAGES];                 /* side_gate_params_mm   */
  iisan_linear_ptrs fc_text, fc_img, fc_mm;         /* fc_bert, fc_cv, fc_mm */
  iisan_linear_ptrs pre_text, pre_img, mm_down;     /* bert_pre_fc, cv_pre_fc, fc_mm_down */
} iisan_san_params;

/* AdapterBlock activation (CC/model/modules.py:104-107: nn.GELU() when args.adapter_activation == "GELU", else nn.ReLU()).
 * GELU is the exact (erf) form, torch's default.  The fused chain kernels are ReLU-only: GELU runs on the layered path. */
typedef enum iisan_activation { IISAN_ACT_RELU = 0, IISAN_ACT_GELU = 1 } iisan_activation;

typedef struct iisan_san_desc {
  int32_t n_items;                  /* rows N: B*11 for a train batch, b for the eval sweep */
  int32_t d_text, d_img, d_mm;      /* hidden widths; d_mm = min(d_text, d_img) */
  int32_t layers_text, layers_img;  /* cached states per item (n_layers + 1) */
  int32_t r_text, r_img, r_mm;      /* adapter bottlenecks */
  int32_t emb;                      /* E */
  int32_t n_stages;
  /* stage plan (CC/model/model.py:318-338 ; CA/model/model.py:353-417); -1 = tower idle */
  int32_t text_adapter[IISAN_MAX_STAGES], text_layer[IISAN_MAX_STAGES];
  int32_t img_adapter[IISAN_MAX_STAGES], img_layer[IISAN_MAX_STAGES];
  int32_t mm_index[IISAN_MAX_STAGES];
  int32_t asym;         /* 0: CC heads (fc d->d, pre_fc d->E); 1: CA heads (fc d->E, pre_fc E->E) */
  int32_t remove_first; /* 1: towers start from hidden state 0 (CC/model/model.py:305-308) */
  int32_t state_dtype;  /* iisan_dtype of the cached hidden states */
  int32_t compute;      /* iisan_compute */
  int32_t out_ld;       /* leading dimension of `out` (>= 3*emb) */
  int32_t activation;   /* iisan_activation of the AdapterBlocks: args.adapter_activation (CC/model/modules.py:104-107) */
} iisan_san_desc;

size_t iisan_san_workspace_bytes(const iisan_san_desc* desc);

/* image: [N, layers_img, d_img], text: [N, layers_text, d_text] of desc->state_dtype, contiguous.
 * out: fp32 [N, out_ld]; columns [0,E) = cv, [E,2E) = text, [2E,3E) = mm embedding -- the operand
 * order of torch.cat at CC/model/model.py:72.  Replaces CC/model/model.py:300-349. */
int iisan_san_forward(const iisan_san_desc* desc, const iisan_san_params* params,
                      const void* image, const void* text, void* workspace, size_t workspace_bytes,
                      float* out, iisan_stream_t stream);

/* d_out: fp32 [N, out_ld] gradient of `out`.  Accumulates every parameter gradient into `grads`
 * (same pointer layout as params).  `workspace` must be the one the forward filled.
 * Replaces the autograd backward of CC/model/model.py:300-349. */
int iisan_san_backward(const iisan_san_desc* desc, const iisan_san_params* params,
                       const iisan_san_params* grads, const void* image, const void* text,
                       void* workspace, size_t workspace_bytes, const float* d_out,
                       iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense layer  y = x W^T + b   (com_dense: CC/model/model.py:37-38,72)
 * ------------------------------------------------------------------------------------------- */
int iisan_linear_forward(int32_t rows, int32_t out_features, int32_t in_features, const float* x,
                         int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                         int32_t compute, iisan_stream_t stream);
/* dx (nullable) is overwritten; dw/db are accumulated. */
int iisan_linear_backward(int32_t rows, int32_t out_features, int32_t in_features, const float* x,
                          int64_t ldx, const float* w, const float* dy, int64_t lddy, float* dx,
                          int64_t lddx, float* dw, float* db, int32_t compute,
                          iisan_stream_t stream);

/* bf16 tensor-core GEMM primitive (tcgen05/TMEM/TMA) that the fast mode is built from; exported so that the
 * parity tests can pin it in isolation.  D[M,N] = act(A x B + bias), fp32 accumulate.
 *   a_mn_major = 0: A stored [M,K] row-major (pitch a_pitch) ; 1: A stored [K,M] row-major
 *   b_mn_major = 0: B stored [N,K] row-major (nn.Linear weight) ; 1: B stored [K,N] row-major
 * Built combinations: K/K (forward), K/MN (data gradients: the weight is read in place), MN/MN (weight gradients); MN/K is
 * not.  out_f32 / out_bf16: either may be null.  splitk > 1 accumulates
 * atomically into a caller-zeroed out_f32.  N % 8 == 0, pitches % 8 == 0, 16-byte aligned pointers. */
int iisan_gemm_bf16(int32_t M, int32_t N, int32_t K, const void* A, int64_t a_pitch, int32_t a_mn_major,
                    const void* B, int64_t b_pitch, int32_t b_mn_major, float* out_f32, int64_t ld_f32,
                    void* out_bf16, int64_t ld_bf16, const float* bias, int32_t relu, int32_t splitk,
                    iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SASRec user encoder  (CC/model/encoders.py:37-58 User_Encoder ;
 *                       CC/model/modules.py:6-96 FFN / attention / TransformerEncoder)
 * ------------------------------------------------------------------------------------------- */
typedef struct iisan_ue_block_ptrs {
  float *w_q, *w_k, *w_v, *w_fc; /* [E,E], bias-free */
  float *ln1_w, *ln1_b;          /* multi_head_attention.layer_norm */
  float *w1, *b1;                /* feed_forward.w_1 [4E,E],[4E] */
  float *w2, *b2;                /* feed_forward.w_2 [E,4E],[E]  */
  float *ln2_w, *ln2_b;          /* feed_forward.layer_norm */
} iisan_ue_block_ptrs;

typedef struct iisan_ue_params {
  float* pos_emb;      /* position_embedding.weight [max_seq_len, E] */
  float *ln_w, *ln_b;  /* transformer_encoder.layer_norm */
  iisan_ue_block_ptrs blocks[IISAN_MAX_BLOCKS];
} iisan_ue_params;

typedef struct iisan_ue_desc {
  int32_t users;      /* B */
  int32_t seq_len;    /* L = max_seq_len (10) */
  int32_t emb;        /* E */
  int32_t heads;
  int32_t n_blocks;
  int32_t training;   /* 1: apply dropout with (seed, offset) */
  float dropout_p;
  uint64_t seed, offset; /* Philox counter-based dropout stream */
  int32_t compute;
  int32_t reserved;
  const uint64_t* offset_dev; /* optional device counter read instead of `offset` (CUDA-graph replays advance it on the device) */
} iisan_ue_desc;

size_t iisan_user_encoder_workspace_bytes(const iisan_ue_desc* desc);

/* embs: fp32, element (u, t, e) at embs[u*ld_user + t*E + e] (ld_user = 11*E for the
 * input_embs[:, :-1, :] slice at CC/model/model.py:76).  log_mask fp32 [B, L].  out fp32 [B, L, E]. */
int iisan_user_encoder_forward(const iisan_ue_desc* desc, const iisan_ue_params* params,
                               const float* embs, int64_t ld_user, const float* log_mask,
                               void* workspace, size_t workspace_bytes, float* out,
                               iisan_stream_t stream);
/* d_embs is OVERWRITTEN at the (u, t<L) positions with the same strides as embs. */
int iisan_user_encoder_backward(const iisan_ue_desc* desc, const iisan_ue_params* params,
                                const iisan_ue_params* grads, const float* embs, int64_t ld_user,
                                const float* log_mask, void* workspace, size_t workspace_bytes,
                                const float* d_out, float* d_embs, iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * In-batch softmax cross-entropy with popularity debias, column-pad mask and reject mask
 * (CC/model/model.py:63-64, 81-105).  Rows are the local users' B*L positions; columns are the
 * item slots of `col_users` users (== the local batch in the reference; the all-gathered global
 * batch for the global negative pool).  Logits are never materialised.
 * ------------------------------------------------------------------------------------------- */
typedef struct iisan_ce_desc {
  int32_t row_users;   /* B (local) */
  int32_t col_users;   /* users whose 11 slots form the columns */
  int32_t seq_len;     /* L */
  int32_t emb;         /* E */
  int64_t user_offset; /* index of local user 0 inside the column pool */
  int32_t compute;
  int32_t reserved;
} iisan_ce_desc;

size_t iisan_inbatch_ce_workspace_bytes(const iisan_ce_desc* desc);

/* prec fp32 [B*L, E]; score fp32 [col_users*(L+1), E]; ids int64; log_mask fp32; pop_prob fp32
 * table indexed by item id.  Outputs (device): loss_sum = sum over valid rows of the row CE,
 * n_valid = number of valid rows (int32), loss = loss_sum / n_valid. */
int iisan_inbatch_ce_forward(const iisan_ce_desc* desc, const float* prec, const float* score,
                             const int64_t* ids_rows, const int64_t* ids_cols,
                             const float* log_mask_rows, const float* log_mask_cols,
                             const float* pop_prob, void* workspace, size_t workspace_bytes,
                             float* loss_sum, int32_t* n_valid, float* loss,
                             iisan_stream_t stream);
/* Upstream gradients are device fp32 scalars, either may be null: grad_loss_sum (d/d loss_sum) and
 * grad_loss_mean (d/d loss, where loss = loss_sum / n_valid).  Every logit gradient is scaled by
 * (grad_loss_sum + grad_loss_mean / n_valid); n_valid is the device int32 the forward wrote.  (The
 * global negative pool divides loss_sum by the all-reduced row count outside and feeds
 * grad_loss_sum.)  d_prec [B*L,E] and d_score [col_users*(L+1),E] are OVERWRITTEN. */
int iisan_inbatch_ce_backward(const iisan_ce_desc* desc, const float* prec, const float* score,
                              const int64_t* ids_rows, const int64_t* ids_cols,
                              const float* log_mask_rows, const float* log_mask_cols,
                              const float* pop_prob, void* workspace, size_t workspace_bytes,
                              const float* grad_loss_sum, const float* grad_loss_mean,
                              const int32_t* n_valid, float* d_prec, float* d_score,
                              iisan_stream_t stream);

/* Bit-exact mask probe used by the parity tests: writes uint8 [B*L, cols] with
 * bit0 = column-pad masked, bit1 = reject masked, bit2 = label column, bit3 = row valid. */
int iisan_inbatch_ce_masks(const iisan_ce_desc* desc, const int64_t* ids_rows,
                           const int64_t* ids_cols, const float* log_mask_rows,
                           const float* log_mask_cols, uint8_t* out, iisan_stream_t stream);

/* Same probe for the fast mode (desc->compute == IISAN_COMPUTE_BF16, emb == 64): the tensor-core CE evaluates its masks
 * from one bit per (row-user, column) built by exact int64 compares; this expands those bits: bit0 = masked as applied by
 * the kernels (column-pad OR reject, the label column escaping the reject), bit2 = label column, bit3 = row valid.
 * `workspace` as for iisan_inbatch_ce_forward. */
int iisan_inbatch_ce_masks_fast(const iisan_ce_desc* desc, const int64_t* ids_rows, const int64_t* ids_cols,
                                const float* log_mask_rows, const float* log_mask_cols, void* workspace,
                                size_t workspace_bytes, uint8_t* out, iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Cached hidden-state path (CC/data_utils/dataset.py:29-34,65-92 ; CC/run.py:368-377):
 * per-item, per-layer gather from an item table into the dense train batch.
 * table: [n_table_items, layers, d] of `dtype`, in HBM or in mapped pinned host memory.
 * out:   [n, n_sel, d] holding for row i the selected layers sel[0..n_sel) of item ids[i];
 *        rows whose id is 0 (padding) are zero-filled without touching the table.
 * ------------------------------------------------------------------------------------------- */
int iisan_gather_states(const void* table, int32_t dtype, int64_t n_table_items, int32_t layers,
                        int32_t d, const int64_t* ids, int32_t n, const int32_t* sel, int32_t n_sel,
                        void* out, iisan_stream_t stream);

/* Layer selection + cast of a dense batch of cached states to bf16 (fast mode with states stored as the reference writes them:
 * fp32 .pt files, CC/preprocess_vectors.py:27-31; fp16 for the LLaMA / EVA-CLIP caches of Code_Cached_Asym):
 *   out_bf16[i, a, :] = bf16_rn(states[i, sel[a], :]),  states [n, layers, d] of `dtype`, out [n, n_sel, d] bf16.
 * The result is a "packed" batch (what iisan_gather_states produces from a bf16 store): the fused chain kernels then serve the
 * reference's on-disk dtype too.  d % 8 == 0, 16-byte aligned pointers. */
int iisan_pack_states(const void* states, int32_t dtype, int64_t n, int32_t layers, int32_t d, const int32_t* sel,
                      int32_t n_sel, void* out_bf16, iisan_stream_t stream);

/* 1 if this (fast-mode) configuration runs on the fused side-adapter chain kernels once its states are bf16 (equal widths,
 * bottleneck 64, every tower active in every stage, towers starting from zero: CC/model/model.py:300-338 as launched by the
 * reference's scripts), else 0 (it runs on the layered path whatever the stored dtype). */
int iisan_san_fused_eligible(const iisan_san_desc* desc);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step (CC/run.py:260-307 name-routed LR groups, :383-385 step): fused multi-tensor Adam with torch.optim.Adam's
 * default hyper-parameters semantics (weight_decay 0, amsgrad off).  `step_dev` is a device fp32 scalar holding the step count;
 * with advance_step != 0 it is incremented (stream-ordered) before the update, so the call is CUDA-graph safe.
 * ------------------------------------------------------------------------------------------- */
#define IISAN_ADAM_MAX_TENSORS 192
typedef struct iisan_adam_tensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t numel;
  float lr;
  int32_t reserved;
} iisan_adam_tensor;
int iisan_adam_step(const iisan_adam_tensor* tensors, int32_t n, float beta1, float beta2, float eps, float* step_dev,
                    int32_t advance_step, iisan_stream_t stream);

/* Host -> device staging of one train batch of cached states with layer selection (CC/run.py:370-374 moves all
 * `layers` states of every slot; only the `n_sel` layers the towers read are needed).  host_src: PINNED host memory,
 * [n_rows, layers, d] of `dtype`; dev_dst: device buffer of the same shape -- rows of unselected layers are left untouched
 * (the kernels never read them).  `sel` is a HOST array, strictly increasing; adjacent layers are merged into one 2-D DMA.
 * Asynchronous on `stream`. */
int iisan_stage_states_h2d(const void* host_src, void* dev_dst, int64_t n_rows, int32_t layers, int32_t d, int32_t dtype,
                           const int32_t* sel, int32_t n_sel, iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation scoring: 1-based rank of each user's held-out item among the catalogue.
 * Replaces the per-user Python loop of CC/data_utils/metrics.py:212-222 + metrics_topK (:59-67):
 *   scores = prec . item_embs^T over ids 0..item_num ; scores[history] = -inf ; id 0 dropped ; rank = position of the target
 *   in the descending order = 1 + #{ i in 1..item_num, i not in history : score_i > score_target }.
 * prec fp32 [users, emb]; item_embs fp32 [n_items1 = item_num + 1, emb] (16-byte aligned, emb % 4 == 0, emb <= 256);
 * targets int64 [users] (ids in 1..item_num); history int64 [users, hist_len], padded with 0 (may be NULL when hist_len == 0,
 * duplicates allowed); ranks int32 [users].  Hit@K = [rank <= K], nDCG@K = 1 / log2(rank + 1).
 * ------------------------------------------------------------------------------------------- */
int iisan_eval_ranks(const float* prec, const float* item_embs, const int64_t* targets, const int64_t* history, int32_t users,
                     int32_t n_items1, int32_t emb, int32_t hist_len, int32_t* ranks, iisan_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement probe (not on the product path; scripts/probe_tile_stream.py): issues the TMA box loads of the fused chain
 * kernels' hidden-state tiles and nothing else.  base: bf16 [n_rows, layers, d] (contiguous == 0: the reference layout, a
 * tile = 128 segments of 128 B at the row pitch layers*d*2) or the same number of [128, 64] tiles stored back to back
 * (contiguous != 0).  One CTA per 128 rows walks sel[0..n_sel) x d/64 chunks `repeat` times through a ring of `slots` tiles.
 * sink: 8 writable device bytes.  Replaces nothing in the reference: it measures the ceiling of streaming
 * CC/run.py:373-374's [B, 11, 13, 768] tensors in place.
 * ------------------------------------------------------------------------------------------- */
int iisan_probe_tile_stream(const void* base, int64_t n_rows, int32_t layers, int32_t d, const int32_t* sel, int32_t n_sel,
                            int32_t slots, int32_t repeat, int32_t contiguous, void* sink, iisan_stream_t stream);

/* Measurement switch (not on the product path; scripts/chain_ab.py): generation of the fused chain kernels used from now on
 * (1 = first, 2 = second generation, 3 = resident-state forward + low-rank adjoint backward: the default wherever it applies;
 * the newest generation that covers a shape is used).  Returns the previous value; gen < 1 only queries. */
int iisan_debug_chain_generation(int32_t gen);

#ifdef __cplusplus
}
#endif
#endif /* IISAN_B200_H_ */
