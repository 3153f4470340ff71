"""TEST INFRASTRUCTURE -- CPU fp32 restatement of the IISAN(Cached) train-step algorithm.

Not product code (see ``oracle/__init__.py``).  Parity status: PINNED against the reference's own
``model`` package by ``tests/golden/*.npz`` (generated with ``oracle/make_golden.py``).

Every function cites the reference lines it restates (paths relative to /root/reference):
  CC = Code_Cached, CA = Code_Cached_Asym.
Integer / index / mask logic is written in numpy integer arithmetic (bit-exact); floating point is
plain fp32 PyTorch on CPU so that autograd supplies the reference gradients.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .synthetic import PathConfig, _adapter_counts

NEG_MASK = -1e4      # CC/model/model.py:89,100
ATT_NEG = -1e9       # CC/model/encoders.py:57
LN_EPS = 1e-6        # CC/model/modules.py:11,52,83
GATE_T = 0.1         # CC/model/model.py:321


# ----------------------------------------------------------------------------------------------
# integer side: layer selection, labels, masks  (bit-exact)
# ----------------------------------------------------------------------------------------------

def stage_plan(cfg: PathConfig):
    """List of stages ``(text_adapter, text_layer, img_adapter, img_layer, mm_index)``; entries are
    ``None`` where a tower does not run in that stage.

    CC/model/model.py:318-338 (all towers every stage);
    CA/model/model.py:353-417 (group layer-drop: the longer tower runs ``diff`` solo stages first).
    """
    lt, li = cfg.text_layers_selected(), cfg.img_layers_selected()
    a_text, a_img, _ = _adapter_counts(cfg)
    plan = []
    if not cfg.asym:
        assert a_text == a_img
        for s in range(a_img):
            plan.append((s, lt[s], s, li[s], s))
        return plan
    diff_text = max(0, a_text - a_img)
    diff_cv = max(0, a_img - a_text)
    for s in range(diff_text):
        plan.append((s, lt[s], None, None, None))
    for s in range(diff_cv):
        plan.append((None, None, s, li[s], None))
    for s in range(min(a_text, a_img)):
        plan.append((s + diff_text, lt[s + diff_text], s + diff_cv, li[s + diff_cv], s))
    return plan


def ce_labels(B: int, L: int, user_offset: int = 0) -> np.ndarray:
    """Label column of row (i, j): ``i*L + i + j`` for j in 1..L == (i)*(L+1) + j'+1 for j' in 0..L-1.

    CC/model/model.py:82-85.  ``user_offset`` shifts users for the global-negative pool (rank w owns
    users [w*B, (w+1)*B) of the concatenated batch).
    """
    i = np.arange(B, dtype=np.int64)[:, None] + user_offset
    j = np.arange(L, dtype=np.int64)[None, :]
    return (i * (L + 1) + j + 1).reshape(-1)


def column_valid(log_mask_cols: np.ndarray) -> np.ndarray:
    """Column (u, p) is kept iff ``cat(log_mask, 1)[u, p] != 0``.  CC/model/model.py:88-89."""
    Bc = log_mask_cols.shape[0]
    full = np.concatenate([log_mask_cols, np.ones((Bc, 1), dtype=log_mask_cols.dtype)], axis=1)
    return (full.reshape(-1) != 0)


def reject_mask(ids_rows: np.ndarray, ids_cols: np.ndarray, L: int, user_offset: int = 0) -> np.ndarray:
    """bool [B*L, C]: True where the logit is overwritten with -1e4 by the reject loop.

    CC/model/model.py:91-100: for row-user i every column whose item id occurs anywhere in
    ``id_list[i]`` (all 11 slots, id 0 included) is rejected for all L rows of that user, except the
    row's own label column.
    """
    B, S = ids_rows.shape
    cols = ids_cols.reshape(-1)
    member = (cols[None, None, :] == ids_rows[:, :, None]).any(axis=1)      # [B, C]
    m = np.repeat(member[:, None, :], L, axis=1).reshape(B * L, -1).copy()
    lab = ce_labels(B, L, user_offset)
    m[np.arange(B * L), lab] = False
    return m


def valid_rows(log_mask_rows: np.ndarray) -> np.ndarray:
    """Row indices kept for the loss.  CC/model/model.py:102."""
    return np.nonzero(log_mask_rows.reshape(-1) != 0)[0]


# ----------------------------------------------------------------------------------------------
# floating-point side
# ----------------------------------------------------------------------------------------------

def _gate(p: torch.Tensor) -> torch.Tensor:
    return torch.sigmoid(p / GATE_T)                                          # CC/model/model.py:321


def _adapter(P, prefix: str, x: torch.Tensor, act: str = "RELU") -> torch.Tensor:
    """``fc_up(act(fc_down(x))) + x`` -- CC/model/modules.py:113-116 (dropout is never applied)."""
    z = F.linear(x, P[prefix + ".fc_down.weight"], P[prefix + ".fc_down.bias"])
    z = F.gelu(z) if act == "GELU" else F.relu(z)
    return F.linear(z, P[prefix + ".fc_up.weight"], P[prefix + ".fc_up.bias"]) + x


def san_forward(P, image: torch.Tensor, text: torch.Tensor, cfg: PathConfig, prefix="mm_encoder."):
    """Side-adapter network.  CC/model/model.py:300-349 ; CA/model/model.py:326-429.

    image ``[..., layers_img, d_img]``, text ``[..., layers_text, d_text]`` (4-D train batch or 3-D
    eval batch -- the ``dim() == 4`` switch at CC/model/model.py:301).  Returns (cv, text, mm)
    embeddings ``[N, E]``.
    """
    h_cv = image.reshape(-1, image.shape[-2], image.shape[-1]).float()
    h_tx = text.reshape(-1, text.shape[-2], text.shape[-1]).float()
    N = h_cv.shape[0]
    d_mm = min(cfg.d_text, cfg.d_img) if cfg.asym else cfg.d_text
    if cfg.remove_first == "TRUE":                                            # CC :305-308
        last_cv, last_tx = h_cv[:, 0], h_tx[:, 0]
    else:                                                                     # CC :311-313
        last_cv = torch.zeros(N, cfg.d_img); last_tx = torch.zeros(N, cfg.d_text)
    last_mm = torch.zeros(N, d_mm)                                            # CA :349-351 broadcasts a 1-D zero
    act = cfg.adapter_activation
    for (ta, tl, ia, il, mi) in stage_plan(cfg):
        if ia is not None:
            g = _gate(P[f"{prefix}side_gate_params_cv.{ia}"])
            fusion_cv = g * h_cv[:, il] + (1 - g) * last_cv                   # CC :320-322
        if ta is not None:
            g = _gate(P[f"{prefix}side_gate_params_text.{ta}"])
            fusion_tx = g * h_tx[:, tl] + (1 - g) * last_tx                   # CC :324-326
        if ta is not None:
            last_tx = _adapter(P, f"{prefix}bert_adapter_list.{ta}", fusion_tx, act)   # CC :331
        if ia is not None:
            last_cv = _adapter(P, f"{prefix}cv_adapter_list.{ia}", fusion_cv, act)     # CC :332
        if mi is not None:
            mm_tx, mm_cv = h_tx[:, tl], h_cv[:, il]
            if cfg.asym and cfg.d_text > cfg.d_img:                           # CA :406-408
                mm_tx = F.linear(mm_tx, P[f"{prefix}down_project_list.{mi}.weight"], P[f"{prefix}down_project_list.{mi}.bias"])
            elif cfg.asym and cfg.d_img > cfg.d_text:                         # CA :409-411
                mm_cv = F.linear(mm_cv, P[f"{prefix}down_project_list.{mi}.weight"], P[f"{prefix}down_project_list.{mi}.bias"])
            g = _gate(P[f"{prefix}side_gate_params_mm.{mi}"])
            last_mm = last_mm + g * mm_cv + (1 - g) * mm_tx                   # CC :335-337
            last_mm = _adapter(P, f"{prefix}mm_adapter_list.{mi}", last_mm, act)       # CC :338
    lin = lambda n, x: F.linear(x, P[f"{prefix}{n}.weight"], P[f"{prefix}{n}.bias"])
    e_tx = lin("bert_pre_fc", lin("fc_bert", last_tx))                        # CC :340,346
    e_cv = lin("cv_pre_fc", lin("fc_cv", last_cv))                            # CC :341,345
    e_mm = lin("fc_mm_down", lin("fc_mm", last_mm))                           # CC :342,347
    return e_cv, e_tx, e_mm


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def user_encoder_forward(P, embs: torch.Tensor, log_mask: torch.Tensor, cfg: PathConfig,
                         prefix="user_encoder.transformer_encoder."):
    """SASRec encoder, dropout disabled (eval / drop_rate 0).

    CC/model/encoders.py:53-58 (mask), CC/model/modules.py:89-96 (embedding + LN),
    :54-64 (attention block), :14-18 (FFN block).
    """
    B, L, E = embs.shape
    H = cfg.heads; dk = E // H
    keep = (log_mask != 0)[:, None, None, :].expand(B, 1, L, L)
    keep = torch.tril(keep)
    att_mask = torch.where(keep, 0.0, ATT_NEG)                                # [B,1,L,L]
    x = _ln(embs + P[prefix + "position_embedding.weight"][None, :L], P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"])
    for b in range(cfg.blocks):
        p = f"{prefix}transformer_blocks.{b}."
        a = p + "multi_head_attention."
        q = F.linear(x, P[a + "w_Q.weight"]).view(B, L, H, dk).transpose(1, 2)
        k = F.linear(x, P[a + "w_K.weight"]).view(B, L, H, dk).transpose(1, 2)
        v = F.linear(x, P[a + "w_V.weight"]).view(B, L, H, dk).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-2, -1)) / (dk ** 0.5) + att_mask
        ctx = torch.matmul(torch.softmax(att, dim=-1), v).transpose(1, 2).reshape(B, L, E)
        x = _ln(x + F.linear(ctx, P[a + "fc.weight"]), P[a + "layer_norm.weight"], P[a + "layer_norm.bias"])
        f = p + "feed_forward."
        y = F.linear(F.relu(F.linear(x, P[f + "w_1.weight"], P[f + "w_1.bias"])), P[f + "w_2.weight"], P[f + "w_2.bias"])
        x = _ln(x + y, P[f + "layer_norm.weight"], P[f + "layer_norm.bias"])
    return x


def inbatch_ce(prec: torch.Tensor, score: torch.Tensor, debias: torch.Tensor,
               ids_rows: np.ndarray, log_mask_rows: np.ndarray,
               ids_cols: np.ndarray, log_mask_cols: np.ndarray, user_offset: int = 0,
               n_valid_total: int | None = None):
    """In-batch softmax CE with debias, column-pad mask and reject mask.

    CC/model/model.py:81-105.  Rows = local users, columns = ``ids_cols`` users (the same batch for
    the reference; the concatenated global batch for the global-negative pool).
    Returns (loss, logits_masked[valid rows]).
    """
    B, S = ids_rows.shape; L = S - 1
    logits = prec @ score.t() - debias[None, :]                               # :86-87
    colv = torch.from_numpy(column_valid(log_mask_cols))
    logits = torch.where(colv[None, :], logits, torch.tensor(NEG_MASK))       # :88-89
    rej = torch.from_numpy(reject_mask(ids_rows, ids_cols, L, user_offset))
    logits = torch.where(rej, torch.tensor(NEG_MASK), logits)                 # :92-100
    rows = torch.from_numpy(valid_rows(log_mask_rows))
    labels = torch.from_numpy(ce_labels(B, L, user_offset))
    lg = logits[rows]
    if n_valid_total is None:
        loss = F.cross_entropy(lg, labels[rows])                              # :104 (mean)
    else:
        loss = F.cross_entropy(lg, labels[rows], reduction="sum") / n_valid_total
    return loss, lg


def model_forward(P, batch, pop_prob: np.ndarray, cfg: PathConfig):
    """Whole ``ModelMM.forward`` (CC/model/model.py:61-105).  ``batch`` holds numpy arrays
    ids [B,11] int64, log_mask [B,10] f32, image, text.  Returns dict of tensors."""
    ids, lm = batch["ids"], batch["log_mask"]
    B, S = ids.shape
    image = torch.as_tensor(batch["image"]); text = torch.as_tensor(batch["text"])
    debias = torch.log(torch.from_numpy(pop_prob)[torch.from_numpy(ids.reshape(-1))])      # :63-64
    e_cv, e_tx, e_mm = san_forward(P, image, text, cfg)
    score = F.linear(torch.cat([e_cv, e_tx, e_mm], dim=1), P["com_dense.weight"], P["com_dense.bias"])  # :72
    embs = score.view(B, S, cfg.embedding_dim)
    prec = user_encoder_forward(P, embs[:, :-1], torch.from_numpy(lm), cfg).reshape(-1, cfg.embedding_dim)  # :76-79
    loss, lg = inbatch_ce(prec, score, debias, ids, lm, ids, lm)
    return {"loss": loss, "score_embs": score, "prec_vec": prec, "e_cv": e_cv, "e_text": e_tx,
            "e_mm": e_mm, "logits_valid": lg}


def params_to_torch(params_np, requires_grad=True):
    return {k: torch.tensor(v, dtype=torch.float32, requires_grad=requires_grad) for k, v in params_np.items()}


def train_step_grads(params_np, batch, pop_prob, cfg: PathConfig):
    """loss + gradient of every parameter (fp32 CPU autograd).  Unused parameters -> None."""
    P = params_to_torch(params_np)
    out = model_forward(P, batch, pop_prob, cfg)
    out["loss"].backward()
    grads = {k: (None if v.grad is None else v.grad.detach().numpy()) for k, v in P.items()}
    return {k: v.detach().numpy() for k, v in out.items()}, grads
