"""Evaluation scoring.  CPU: the oracle restatement against outputs of the reference's own metrics_topK
(tests/golden/eval_topk.npz, oracle/make_golden_eval.py).  GPU: iisan_eval_ranks against the oracle, and the whole device-side
evaluate() flow against the reference flow restated with the product's own sub-modules."""
import math
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_topk.npz")


def test_oracle_matches_reference_metrics_topk():
    from oracle import eval_oracle as EO
    z = np.load(GOLD)
    topk = int(z["topk"])
    for u in range(z["scores"].shape[0]):
        h = z["history"][u]; h = h[h > 0]
        hit, ndcg = EO.metrics_topk(z["scores"][u], int(z["targets"][u]), h, topk)
        assert hit == z["hit_ndcg"][u, 0]
        assert abs(ndcg - z["hit_ndcg"][u, 1]) < 1e-6


def test_host_helpers_follow_the_reference_dataset():
    """pad_sequences == BuildMMEvalDataset.__getitem__ (dataset.py:185-191); hit_ndcg == metrics_topK's outputs from a rank."""
    from iisan_b200.eval import hit_ndcg, pad_sequences
    tok, lm, tgt = pad_sequences([[5, 7, 9], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], [4, 2]], 10)
    assert tok[0].tolist() == [0] * 8 + [5, 7] and lm[0].tolist() == [0.0] * 8 + [1.0, 1.0] and int(tgt[0]) == 9
    assert tok[1].tolist() == list(range(1, 11)) and lm[1].sum() == 10 and int(tgt[1]) == 11
    assert tok[2].tolist() == [0] * 9 + [4] and int(tgt[2]) == 2
    z = np.load(GOLD)
    from oracle import eval_oracle as EO
    ranks = [EO.rank_from_scores(z["scores"][u], int(z["targets"][u]), z["history"][u][z["history"][u] > 0]) for u in range(len(z["targets"]))]
    hit, ndcg = hit_ndcg(torch.tensor(ranks), int(z["topk"]))
    assert np.array_equal(hit.numpy(), z["hit_ndcg"][:, 0])
    assert np.allclose(ndcg.numpy(), z["hit_ndcg"][:, 1], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("users,item_num,emb,hmax", [(64, 300, 64, 12), (257, 5000, 64, 40), (33, 1000, 32, 0), (8, 19246, 64, 11)])
def test_eval_ranks_match_oracle(users, item_num, emb, hmax):
    from iisan_b200.eval import eval_ranks, hit_ndcg
    from oracle import eval_oracle as EO
    g = torch.Generator().manual_seed(users + item_num)
    prec = torch.randn(users, emb, generator=g)
    items = torch.randn(item_num + 1, emb, generator=g)
    targets = torch.randint(1, item_num + 1, (users,), generator=g)
    hist = torch.zeros(users, max(hmax, 1), dtype=torch.int64)
    hl = []
    for u in range(users):
        n = int(torch.randint(0, hmax + 1, (1,), generator=g)) if hmax else 0
        h = torch.randint(1, item_num + 1, (n,), generator=g)
        if n >= 2:
            h[1] = h[0]                                   # duplicates in the history are legal
        h = h[h != targets[u]]
        hist[u, :len(h)] = h
        hl.append(h.numpy())
    ranks = eval_ranks(prec.cuda(), items.cuda(), targets.cuda(), hist.cuda() if hmax else None).cpu().numpy()
    ref, margin = EO.ranks_from_embeddings(prec.numpy(), items.numpy(), targets.numpy(), hl)
    ok = margin > 2e-5                                    # a competitor within fp32 noise of the target has no defined order
    assert ok.mean() >= 0.7
    assert np.array_equal(ranks[ok], ref[ok])
    assert np.all(np.abs(ranks[~ok] - ref[~ok]) <= 2)
    hit, ndcg = hit_ndcg(torch.from_numpy(ranks), 10)
    for u in np.nonzero(ok)[0][:50]:
        e = (1.0, 1.0 / math.log2(ref[u] + 1)) if ref[u] <= 10 else (0.0, 0.0)
        assert float(hit[u]) == e[0] and abs(float(ndcg[u]) - e[1]) < 1e-12


@pytest.mark.gpu
def test_evaluate_flow_matches_reference_flow():
    """evaluate() (device-side sweep, gather after com_dense, rank kernel) vs the reference's flow (metrics.py:162-250) restated
    with the same product modules: per-item (cv, text, mm) embeddings gathered per user, com_dense on the gathered [b, L, 3E],
    dense score matrix, per-user masking and argsort."""
    from iisan_b200.eval import evaluate, item_embedding_table, pad_sequences
    from iisan_b200.precision import set_compute_mode
    from oracle import eval_oracle as EO
    from oracle.synthetic import PathConfig, make_params, make_pop_prob
    from product_util import build_product
    cfg = PathConfig(item_num=400)
    params = make_params(cfg, 5, perturb=True)
    pop = make_pop_prob(cfg, 5)
    set_compute_mode("fp32")
    try:
        model = build_product(cfg, params, pop).eval()
        g = torch.Generator().manual_seed(9)
        img = torch.randn(cfg.item_num + 1, 13, 768, generator=g).cuda(); txt = torch.randn(cfg.item_num + 1, 13, 768, generator=g).cuda()
        img[0] = 0; txt[0] = 0
        seqs, hists = [], []
        for u in range(70):
            n = int(torch.randint(2, 12, (1,), generator=g))
            s = torch.randint(1, cfg.item_num + 1, (n,), generator=g).tolist()
            seqs.append(s); hists.append(s[:-1])
        hit, ndcg = evaluate(model, img, txt, seqs, hists, topk=10, batch=32)
        with torch.no_grad():
            cv, (tx, mm) = model.mm_encoder(img, txt)
            table = model.com_dense(torch.cat([cv, tx, mm], dim=1))
            tok, lm, tgt = pad_sequences(seqs, cfg.max_seq_len)
            gathered = torch.cat([cv[tok.cuda()], tx[tok.cuda()], mm[tok.cuda()]], dim=2)        # metrics.py:207-209
            prec = model.user_encoder(model.com_dense(gathered), lm.cuda(), "cuda")[:, -1]
            scores = (prec @ table.t()).cpu().numpy()
        res = np.array([EO.metrics_topk(scores[u], int(tgt[u]), np.array(hists[u]), 10) for u in range(len(seqs))])
        assert abs(hit - res[:, 0].mean()) <= 1.0 / len(seqs) + 1e-9          # at most one near-tie may flip
        assert abs(ndcg - res[:, 1].mean()) <= 1.0 / len(seqs) + 1e-9
        assert torch.allclose(item_embedding_table(model, img, txt, batch=128), table, rtol=1e-5, atol=1e-6)
    finally:
        set_compute_mode(None)
