"""TEST INFRASTRUCTURE -- CPU restatement of the reference's evaluation scoring (never imported by the product path).

Follows Code_Cached/data_utils/metrics.py:59-67 (metrics_topK) and :212-222 (history masking, id 0 dropped).
Pinned against outputs of the reference's own metrics_topK (tests/golden/eval_topk.npz, made by oracle/make_golden_eval.py).
"""
import math

import numpy as np


def rank_from_scores(scores, target, history):
    """scores [item_num + 1] (id-indexed, float); returns the 1-based position of ``target`` after
    ``scores[history] = -inf ; scores = scores[1:] ; argsort(descending)``  (metrics.py:59-62, 217-221)."""
    s = np.array(scores, dtype=np.float64, copy=True)
    if len(history):
        s[np.asarray(history, dtype=np.int64)] = -np.inf
    s = s[1:]
    order = np.argsort(-s, kind="stable")
    return int(np.nonzero(order == (target - 1))[0][0]) + 1


def metrics_topk(scores, target, history, topk=10):
    """(hit, ndcg) of one user -- metrics.py:59-67."""
    rank = rank_from_scores(scores, target, history)
    if rank <= topk:
        return 1.0, 1.0 / math.log2(rank + 1)
    return 0.0, 0.0


def ranks_from_embeddings(prec, item_embs, targets, histories):
    """float64 scores from the fp32 operands: prec [U, E] . item_embs [I + 1, E]^T, then rank_from_scores per user.
    Also returns the smallest |score_i - score_target| over competing items (ties within fp32 noise have no defined order)."""
    sc = prec.astype(np.float64) @ item_embs.astype(np.float64).T
    ranks, margins = [], []
    for u in range(prec.shape[0]):
        ranks.append(rank_from_scores(sc[u], int(targets[u]), histories[u]))
        s = sc[u].copy()
        st = s[int(targets[u])]
        s[0] = np.inf
        s[int(targets[u])] = np.inf
        if len(histories[u]):
            s[np.asarray(histories[u], dtype=np.int64)] = np.inf
        margins.append(float(np.min(np.abs(s - st))))
    return np.array(ranks), np.array(margins)
