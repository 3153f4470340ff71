"""Freeze outputs of the REFERENCE's own metrics_topK (Code_Cached/data_utils/metrics.py:59-67) plus the history masking of
eval_model (:217-221) on seeded random scores -> tests/golden/eval_topk.npz.

metrics.py cannot be imported here (its `from .dataset import *` needs lmdb), so the function's source text is read from
/root/reference and executed as is (nothing is copied into the repository).  Run in the build container only:
    python -m oracle.make_golden_eval
"""
import ast
import math
import os

import numpy as np
import torch

REF = "/root/reference/Code_Cached/data_utils/metrics.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eval_topk.npz")


def reference_metrics_topk():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "metrics_topK")
    ns = {"torch": torch, "math": math, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["metrics_topK"]


def main():
    f = reference_metrics_topk()
    g = torch.Generator().manual_seed(12345)
    item_num, users, topk = 300, 64, 10
    scores = torch.randn(users, item_num + 1, generator=g)
    targets = torch.randint(1, item_num + 1, (users,), generator=g)
    hist = []
    res = []
    item_rank = torch.Tensor(np.arange(item_num) + 1)
    for u in range(users):
        n = int(torch.randint(0, 12, (1,), generator=g))
        h = torch.randint(1, item_num + 1, (n,), generator=g)
        h = h[h != targets[u]]
        hist.append(h.numpy())
        score = scores[u].clone()
        if u % 3 == 0:                                   # make the target a top item now and then
            score[targets[u]] = 3.0 + 0.01 * u
            scores[u, targets[u]] = score[targets[u]]
        score[h] = -np.inf                               # metrics.py:217-219
        score = score[1:]                                # :220
        labels = torch.zeros(item_num); labels[targets[u] - 1] = 1.0        # dataset.py:207-208
        res.append(f(score, labels, item_rank, topk, "cpu").numpy())
    hmax = max(len(h) for h in hist)
    hp = np.zeros((users, hmax), dtype=np.int64)
    for u, h in enumerate(hist):
        hp[u, :len(h)] = h
    np.savez_compressed(OUT, scores=scores.numpy(), targets=targets.numpy(), history=hp, hit_ndcg=np.stack(res), topk=topk)
    print("wrote", OUT, np.stack(res).mean(0))


if __name__ == "__main__":
    main()
