"""The oracle (oracle/iisan_oracle.py) must reproduce the reference's own outputs frozen in
tests/golden (fp32, CPU): loss, embeddings, bit-exact masks/labels, and every parameter gradient."""
import numpy as np
import pytest

from golden_util import CASES, ORACLE_ONLY_CASES, VERSA_CASES, check_grads, golden_masked, load_case, rebuild_inputs


@pytest.mark.parametrize("name", CASES + VERSA_CASES + ORACLE_ONLY_CASES)
def test_oracle_matches_reference_fixture(name):
    from oracle import iisan_oracle as O
    z, meta = load_case(name)
    cfg, batch, params, pop = rebuild_inputs(meta)
    out, grads = O.train_step_grads(params, batch, pop, cfg)
    np.testing.assert_allclose(out["loss"], z["loss"], rtol=1e-5)
    for k in ("score_embs", "prec_vec", "e_cv", "e_text", "e_mm"):
        np.testing.assert_allclose(out[k], z[k], rtol=1e-4, atol=2e-6, err_msg=k)
    # bit-exact: which logits the reference overwrote with -1e4, and the labels of the valid rows
    masked = out["logits_valid"] == O.NEG_MASK
    assert masked.shape == tuple(z["masked_shape"])
    assert np.array_equal(masked, golden_masked(z))
    L = cfg.max_seq_len
    rows = O.valid_rows(batch["log_mask"])
    assert np.array_equal(O.ce_labels(meta["B"], L)[rows], z["labels_valid"])
    check_grads(z, grads, rtol=2e-4)


def test_mask_restatement_properties():
    """Size-independent properties of the integer restatement (bit-exact side)."""
    from oracle import iisan_oracle as O
    from oracle.synthetic import PathConfig, make_ids
    cfg = PathConfig(item_num=30)
    ids, lm = make_ids(64, cfg, 9, "realistic")
    L = cfg.max_seq_len
    rej = O.reject_mask(ids, ids, L)
    lab = O.ce_labels(64, L)
    assert not rej[np.arange(64 * L), lab].any()                 # label never rejected
    colv = O.column_valid(lm)
    assert colv.reshape(64, 11)[:, -1].all()                     # last slot always valid
    assert np.array_equal(colv.reshape(64, 11)[:, :-1], lm != 0)
    # own items are rejected for every row of that user (except the label column)
    for u in range(64):
        for p in range(11):
            c = u * 11 + p
            rows = np.arange(u * L, (u + 1) * L)
            exp = np.ones(L, bool); 
            if p >= 1:
                exp[p - 1] = False
            assert np.array_equal(rej[rows, c], exp)
    # global pool: the rows of rank 1 against the concatenated columns equal the corresponding slice
    rej_g = O.reject_mask(ids[32:], ids, L, user_offset=32)
    assert np.array_equal(rej_g, rej[32 * L:])
