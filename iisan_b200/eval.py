"""Evaluation scoring of the IISAN(Cached) path (Code_Cached/data_utils/metrics.py:59-67, 162-250,
Code_Cached/data_utils/dataset.py:172-223).

The reference scores every user against the whole catalogue, masks the user's history, drops id 0 and ranks the held-out
item with a per-user ``argsort`` in Python.  Here the rank comes from one kernel (``iisan_eval_ranks``: no sort, no
[users, items] score matrix) and Hit@K / nDCG@K follow from it.  ``evaluate`` runs the whole flow on the device: item sweep
through the side-adapter network (3-D input mode) -> com_dense -> user encoder on the padded histories -> ranks.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def eval_ranks(prec, item_embs, targets, history=None):
    """1-based rank [U] (int32) of ``targets`` among ids 1..item_num for the user vectors ``prec`` [U, E] against
    ``item_embs`` [item_num + 1, E]; ``history`` [U, H] int64 padded with 0: ids that do not compete (metrics.py:217-219).

    Edge cases against the reference's argsort (metrics.py:212-222): exact score ties resolve to the BEST rank here (strict
    ``>`` count) instead of argsort order; a target that also appears in the user's history is ranked normally here, while the
    reference masks it to -inf (its rank is then item_num) -- the reference's data pipeline never produces such a pair (the
    history is the sequence without its last item).  A target outside [0, item_num] is a caller error (the kernel scores it +inf)."""
    lib = L.load()
    L.require_cuda(prec, "prec_emb")
    prec = prec.contiguous().float()
    item_embs = item_embs.contiguous().float()
    targets = targets.to(prec.device, torch.int64).contiguous().view(-1)
    u, e = prec.shape
    h = 0
    hp = None
    if history is not None and history.numel() > 0:
        history = history.to(prec.device, torch.int64).contiguous().view(u, -1)
        h, hp = history.shape[1], C.c_void_p(history.data_ptr())
    ranks = torch.empty(u, dtype=torch.int32, device=prec.device)
    L.check(lib.iisan_eval_ranks(C.c_void_p(prec.data_ptr()), C.c_void_p(item_embs.data_ptr()), C.c_void_p(targets.data_ptr()), hp, u,
                                 item_embs.shape[0], e, h, C.c_void_p(ranks.data_ptr()),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "iisan_eval_ranks")
    return ranks


def hit_ndcg(ranks, topk=10):
    """(Hit@K, nDCG@K) per user from the ranks -- metrics_topK (metrics.py:59-67): hit = [rank <= K], nDCG = 1/log2(rank + 1)."""
    r = ranks.to(torch.float64)
    hit = (r <= topk).to(torch.float64)
    ndcg = hit / torch.log2(r + 1.0)
    return hit, ndcg


def pad_sequences(seqs, max_seq_len):
    """BuildMMEvalDataset.__getitem__ (dataset.py:185-191): tokens = seq[:-1] left-padded with 0 to ``max_seq_len`` slots,
    log_mask marks the real ones, target = seq[-1].  Returns (tokens int64 [U, L], log_mask fp32 [U, L], targets int64 [U])."""
    u = len(seqs)
    tokens = torch.zeros(u, max_seq_len, dtype=torch.int64)
    log_mask = torch.zeros(u, max_seq_len, dtype=torch.float32)
    targets = torch.zeros(u, dtype=torch.int64)
    for i, seq in enumerate(seqs):
        toks = list(seq[:-1])[-max_seq_len:]
        n = len(toks)
        if n:
            tokens[i, max_seq_len - n:] = torch.as_tensor(toks, dtype=torch.int64)
            log_mask[i, max_seq_len - n:] = 1.0
        targets[i] = int(seq[-1])
    return tokens, log_mask, targets


@torch.no_grad()
def item_embedding_table(model, image_states, text_states, batch=4096):
    """score embedding of every catalogue item: com_dense(cat(mm_encoder(states)))  (get_MM_item_embeddings + eval_model head,
    metrics.py:71-111, 180-186).  ``*_states`` [item_num + 1, layers, d] on the device (row 0 = the padding item)."""
    outs = []
    for i in range(0, image_states.shape[0], batch):
        cv, (tx, mm) = model.mm_encoder(image_states[i:i + batch], text_states[i:i + batch])
        outs.append(model.com_dense(torch.cat([cv, tx, mm], dim=1)))
    return torch.cat(outs, dim=0)


@torch.no_grad()
def evaluate(model, image_states, text_states, eval_seqs, user_history, topk=10, batch=1024):
    """Hit@K / nDCG@K over ``eval_seqs`` (list of id sequences, last id = held-out target) with ``user_history`` (list of id
    lists that must not compete) -- eval_model (metrics.py:162-250) without the host round trips.  Returns (hit, ndcg) means."""
    was_training = model.training
    model.eval()
    try:
        dev = image_states.device
        table = item_embedding_table(model, image_states, text_states)              # [item_num + 1, E]
        L_ = model.max_seq_len
        tokens, log_mask, targets = pad_sequences(eval_seqs, L_)
        hmax = max((len(h) for h in user_history), default=0)
        hist = torch.zeros(len(eval_seqs), max(hmax, 1), dtype=torch.int64)
        for i, h in enumerate(user_history):
            if len(h):
                hist[i, :len(h)] = torch.as_tensor(list(h), dtype=torch.int64)
        hits, ndcgs = [], []
        for i in range(0, len(eval_seqs), batch):
            tok = tokens[i:i + batch].to(dev)
            lm = log_mask[i:i + batch].to(dev)
            embs = table[tok]                                                       # [b, L, E]: the reference gathers cv/text/mm then com_dense; com_dense is per item, so gather after it
            prec = model.user_encoder(embs, lm, dev)[:, -1]
            ranks = eval_ranks(prec, table, targets[i:i + batch].to(dev), hist[i:i + batch].to(dev))
            h_, n_ = hit_ndcg(ranks, topk)
            hits.append(h_); ndcgs.append(n_)
        return float(torch.cat(hits).mean()), float(torch.cat(ndcgs).mean())
    finally:
        model.train(was_training)
