"""Parity of the CUDA path (through the C ABI) against the frozen reference outputs (tests/golden) and the
oracle, fp32 mode: <= 1e-5 relative on loss, tight on embeddings and gradients, bit-exact on masks/labels."""
import numpy as np
import pytest
import torch

from golden_util import CASES, check_grads, golden_masked, load_case, rebuild_inputs
from product_util import emulation_batch, fused_chain_route, build_product, run_step

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


# ------------------------------------------------------------------------------------------------------------------
# fast mode (IISAN_COMPUTE_BF16): tcgen05 GEMMs with bf16 operands, fp32 accumulation.
# Tolerances (BASELINE.json north_star): loss and embeddings <= 1e-2 relative; gradients are checked at 3e-2 of the
# largest reference entry per tensor (bf16 operands carry 2^-9 relative rounding through 7 chained stages).
# ------------------------------------------------------------------------------------------------------------------
BF16_LOSS_RTOL = 1e-2
BF16_EMB_RTOL = 1e-2
# Gradients of the fast mode are pinned against the rounding-point emulation (tests/bf16_emulation.py) in relative L2 per
# tensor.  Measured on B200: median 2e-3, dense batches <= 6e-3; the tail comes from single ReLU units whose
# pre-activation sits within one bf16 ulp of zero (padded item rows are identical, so such a unit flips for all of them
# at once) and from the scalar gate gradients, which are cancellation-heavy sums over N*d terms.
BF16_GRAD_L2 = 5e-2        # every tensor
BF16_GRAD_L2_MEDIAN = 2e-2 # median over tensors.  B=4/8 fixtures reach 1.2e-2 since the SASRec linears run as TF32 tiles: an fp32-ulp
                           # difference in summation order moves ~2e-4 of the operands across a TF32 rounding boundary (and, rarely, a
                           # ReLU unit across zero), which the 40-80 rows of these fixtures do not average out; every SAN gradient
                           # inherits that noise through d score_embs.  B=512: see test_gpu_user_encoder / scripts/dbg_ue_tc.py.
BF16_GATE_RTOL = 0.10      # the 21 scalar gate gradients, pooled into one vector (relative L2)
BF16_GRAD_COS = 0.97       # vs the fp32 reference: cosine of the per-tensor-normalised flattened gradient


def _bf16_round(a):
    return torch.from_numpy(a).bfloat16().float().numpy()


@pytest.mark.parametrize("state_dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("name", CASES)
def test_train_step_bf16_mode(name, state_dtype):
    from bf16_emulation import train_step_grads_emul
    from iisan_b200.precision import set_compute_mode
    from oracle import iisan_oracle as O
    z, meta = load_case(name)
    cfg, batch, params, pop = rebuild_inputs(meta)
    dt = getattr(torch, state_dtype)
    if dt == torch.bfloat16:                     # the references see exactly the stored (rounded) states
        batch = dict(batch, image=_bf16_round(batch["image"]), text=_bf16_round(batch["text"]))
    ref_out, ref_grads = O.train_step_grads(params, batch, pop, cfg)
    plan = O.stage_plan(cfg)
    fused = fused_chain_route(cfg, plan)          # whatever the stored dtype: fp32 / fp16 states are packed to bf16 first (ops.SanFn)
    emu_out, emu_grads = train_step_grads_emul(params, emulation_batch(batch, fused, dt), pop, cfg, ce_bf16=(cfg.embedding_dim == 64), fused_chain=fused)
    set_compute_mode("bf16")
    try:
        model = build_product(cfg, params, pop).eval()
        loss, grads = run_step(model, batch, dtype=dt)
        with torch.no_grad():
            score = model.item_embeddings(torch.from_numpy(batch["image"]).cuda().to(dt),
                                          torch.from_numpy(batch["text"]).cuda().to(dt)).cpu().numpy()
    finally:
        set_compute_mode(None)
    ref_loss, ref_score = float(ref_out["loss"]), ref_out["score_embs"]
    assert abs(float(loss) - ref_loss) <= BF16_LOSS_RTOL * abs(ref_loss), (float(loss), ref_loss)
    assert np.abs(score - ref_score).max() <= BF16_EMB_RTOL * np.abs(ref_score).max()
    assert abs(float(loss) - float(emu_out["loss"])) <= 3e-4 * abs(ref_loss)
    assert np.abs(score - emu_out["score_embs"]).max() <= 2e-3 * np.abs(ref_score).max()
    errs, dot, n1, n2 = [], 0.0, 0.0, 0.0
    gate_ref, gate_got = [], []
    for n, g in emu_grads.items():
        if g is None:
            assert grads[n] is None or not np.any(grads[n]), n
            continue
        err = float(np.linalg.norm((grads[n] - g).astype(np.float64)) / (np.linalg.norm(g.astype(np.float64)) + 1e-30))
        if g.size == 1:
            gate_ref.append(float(g.ravel()[0])); gate_got.append(float(grads[n].ravel()[0]))
        else:
            assert err <= BF16_GRAD_L2, f"{n}: {err}"
            errs.append(err)
        r = ref_grads[n].astype(np.float64).ravel(); o = grads[n].astype(np.float64).ravel()
        s = 1.0 / (np.linalg.norm(r) + 1e-30)        # per-tensor normalisation: every tensor weighs the same
        dot += float(np.dot(r, o)) * s * s; n1 += float(np.dot(r, r)) * s * s; n2 += float(np.dot(o, o)) * s * s
    worst = max(errs)
    gate_err = np.linalg.norm(np.array(gate_got) - np.array(gate_ref)) / np.linalg.norm(gate_ref)
    assert gate_err <= BF16_GATE_RTOL, gate_err
    assert float(np.median(errs)) <= BF16_GRAD_L2_MEDIAN, np.median(errs)
    cos = dot / np.sqrt(n1 * n2)
    assert cos >= BF16_GRAD_COS, cos
    print(f"{name}/{state_dtype}: loss {float(loss):.5f} (fp32 ref {ref_loss:.5f}), worst / median grad L2 err vs emulation {worst:.2e} / {np.median(errs):.2e}, "
          f"cosine vs fp32 reference {cos:.4f}")
