"""Parity of the CUDA path (through the C ABI) against the frozen reference outputs (tests/golden) and the
oracle, fp32 mode: <= 1e-5 relative on loss, tight on embeddings and gradients, bit-exact on masks/labels."""
import numpy as np
import pytest
import torch

from golden_util import CASES, check_grads, golden_masked, load_case, rebuild_inputs
from product_util import build_product, run_step

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference_fixture_fp32(name):
    from iisan_b200.precision import set_compute_mode
    set_compute_mode("fp32")
    z, meta = load_case(name)
    cfg, batch, params, pop = rebuild_inputs(meta)
    model = build_product(cfg, params, pop).eval()
    loss, grads = run_step(model, batch)
    np.testing.assert_allclose(loss, z["loss"], rtol=1e-5)
    # embeddings
    ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
    image = torch.from_numpy(batch["image"]).cuda(); text = torch.from_numpy(batch["text"]).cuda()
    with torch.no_grad():
        score = model.item_embeddings(image, text)
        cv, (tx, mm) = model.mm_encoder(image, text)
        E = cfg.embedding_dim
        prec = model.user_encoder(score.view(-1, 11, E)[:, :-1], torch.from_numpy(batch["log_mask"]).cuda(), "cuda")
    np.testing.assert_allclose(score.cpu().numpy(), z["score_embs"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(cv.cpu().numpy(), z["e_cv"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(tx.cpu().numpy(), z["e_text"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(mm.cpu().numpy(), z["e_mm"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(prec.reshape(-1, E).cpu().numpy(), z["prec_vec"], rtol=1e-4, atol=5e-6)
    worst = check_grads(z, grads, rtol=3e-4)
    print(f"{name}: loss {float(loss):.6f} (ref {float(z['loss']):.6f}), worst grad sample rel err {worst:.2e}")


@pytest.mark.parametrize("name", CASES)
def test_masks_and_labels_bit_exact(name):
    from iisan_b200.ops import inbatch_ce_masks
    from oracle import iisan_oracle as O
    z, meta = load_case(name)
    cfg, batch, _, _ = rebuild_inputs(meta)
    ids = torch.from_numpy(batch["ids"]).cuda(); lm = torch.from_numpy(batch["log_mask"]).cuda()
    bits = inbatch_ce_masks(ids, ids, lm, lm).cpu().numpy()
    L = cfg.max_seq_len
    rows = O.valid_rows(batch["log_mask"])
    assert np.array_equal(np.nonzero(bits[:, 0] & 8)[0], rows)                         # valid-row set
    masked = ((bits & 3) != 0)[rows]
    assert np.array_equal(masked, golden_masked(z))                                   # == reference's -1e4 pattern
    assert np.array_equal(np.argmax((bits & 4) != 0, axis=1)[rows], z["labels_valid"])  # label columns
    assert ((bits & 4) != 0).sum(axis=1).max() == 1
    # and against the oracle's integer restatement, bit for bit, for all rows
    assert np.array_equal((bits & 1) != 0, np.broadcast_to(~O.column_valid(batch["log_mask"]), bits.shape))
    assert np.array_equal((bits & 2) != 0, O.reject_mask(batch["ids"], batch["ids"], L))


def test_global_pool_masks_bit_exact():
    from iisan_b200.ops import inbatch_ce_masks
    from oracle import iisan_oracle as O
    from oracle.synthetic import PathConfig, make_ids
    cfg = PathConfig(item_num=50)
    ids, lm = make_ids(48, cfg, 77, "realistic")
    idc = torch.from_numpy(ids).cuda(); lmc = torch.from_numpy(lm).cuda()
    for off in (0, 16, 32):
        bits = inbatch_ce_masks(idc[off:off + 16], idc, lmc[off:off + 16], lmc, user_offset=off).cpu().numpy()
        assert np.array_equal((bits & 2) != 0, O.reject_mask(ids[off:off + 16], ids, 10, user_offset=off))
        lab = O.ce_labels(16, 10, off)
        assert np.array_equal(np.argmax((bits & 4) != 0, axis=1), lab)


def test_gather_states_bit_exact():
    from iisan_b200.ops import gather_states
    g = torch.Generator().manual_seed(3)
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        table = torch.randn(37, 13, 64, generator=g).to(dtype).cuda()
        ids = torch.randint(0, 37, (55,), generator=g).cuda()
        sel = [0, 2, 4, 12]
        out = gather_states(table, ids, sel)
        exp = table[ids][:, sel]
        exp[ids == 0] = 0
        assert torch.equal(out, exp)
