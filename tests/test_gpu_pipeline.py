"""Parity of the HOST-batch pipeline and of the captured step.

  * PipelinedTrainStep fed pinned host batches of the reference shapes [B, 11, 13, 768] (fp32 and bf16) -- the path behind
    bench.py's ``e2e`` number, Code_Cached/run.py:368-377 -- must give the loss of the plain device-tensor step, bit for bit,
    with every UNSELECTED layer of the device staging buffers poisoned (iisan_stage_states_h2d copies the selected layers only;
    nothing may read the others);
  * TrainStep(use_graph=True) == TrainStep(use_graph=False) without a store;
  * the whole train step at the BENCHMARKED size (BASELINE configs[1]: B = 512 users, item_num 19,246, dense and realistic ids)
    against oracle.model_forward: exact mode loss <= 1e-5 / embeddings 1e-4, fast mode <= 1e-2.
"""
import numpy as np
import pytest
import torch

from product_util import build_product

pytestmark = pytest.mark.gpu


def _setup(B, item_num, seed, mode="realistic"):
    from oracle.synthetic import PathConfig, make_ids, make_params, make_pop_prob
    cfg = PathConfig(item_num=item_num)
    ids, lm = make_ids(B, cfg, seed, mode)
    params = make_params(cfg, seed, perturb=True)
    pop = make_pop_prob(cfg, seed)
    return cfg, ids, lm, params, pop


def _host_batches(n, B, ids_np, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for k in range(n):
        ids = torch.from_numpy(np.roll(ids_np, k, axis=0).copy())
        image = torch.randn(B, 11, 13, 768, generator=g).to(dtype)
        text = torch.randn(B, 11, 13, 768, generator=g).to(dtype)
        pad = ids == 0
        image[pad] = 0
        text[pad] = 0
        out.append((ids, image, text))
    return out


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("state_dtype", [torch.float32, torch.bfloat16])
def test_host_batch_pipeline_equals_device_step(state_dtype, use_graph):
    from iisan_b200.engine import PipelinedTrainStep, TrainStep
    from iisan_b200.optim import FusedAdam
    from iisan_b200.precision import set_compute_mode
    B = 16
    cfg, ids_np, lm_np, params, pop = _setup(B, 150, 5)
    batches = _host_batches(3, B, ids_np, state_dtype, 7)
    lm = torch.from_numpy(lm_np)
    set_compute_mode("bf16")
    try:
        # ---- plain device-tensor steps (eager, no staging code involved) ----
        model = build_product(cfg, params, pop).eval()
        opt = FusedAdam(model.parameters(), lr=1e-3)
        plain = TrainStep(model, opt, use_graph=False)
        ref = [float(plain(i.view(-1).cuda(), im.cuda(), tx.cuda(), lm.cuda()).item()) for i, im, tx in batches]
        # ---- the pipelined host-batch runner ----
        model2 = build_product(cfg, params, pop).eval()
        opt2 = FusedAdam(model2.parameters(), lr=1e-3)
        pipe = PipelinedTrainStep(model2, opt2, use_graph=use_graph)
        sel_i = set(model2.mm_encoder.plan.layers_img_sel); sel_t = set(model2.mm_encoder.plan.layers_text_sel)
        assert len(sel_i) < 13 and len(sel_t) < 13
        pinned = [tuple(t.pin_memory() for t in (i.view(-1), im, tx, lm)) for i, im, tx in batches]
        got = []
        pipe.submit(*pinned[0])
        for k in range(len(pinned)):
            # poison every unselected layer of ALL device staging buffers allocated so far (NaN): the stage op must not have
            # to write them and no kernel may read them
            torch.cuda.synchronize()
            for buf in pipe.bufs:
                if buf is None:
                    continue
                for t, sel in ((buf[1], sel_i), (buf[2], sel_t)):
                    for l in range(13):
                        if l not in sel:
                            t[:, :, l] = float("nan")
            if k + 1 < len(pinned):
                pipe.submit(*pinned[k + 1])
            got.append(float(pipe.run().item()))
        assert np.isfinite(got).all(), got
        # step 1: same parameters, deterministic forward -> bit-equal loss.  Later steps differ only through the reduction order
        # of the gradient atomics of the step before (see test_gpu_store.py).
        assert got[0] == ref[0], (got, ref)
        assert np.allclose(got[1], ref[1], rtol=1e-4), (got, ref)
        assert np.allclose(got[2], ref[2], rtol=5e-3), (got, ref)
    finally:
        set_compute_mode(None)


def test_stage_states_h2d_copies_exactly_the_selected_layers():
    from iisan_b200.engine import stage_states_h2d
    g = torch.Generator().manual_seed(11)
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        src = torch.randn(5, 11, 13, 64, generator=g).to(dtype).pin_memory()
        for sel in ([0, 2, 4, 6, 8, 10, 12], [1, 2, 3, 9], [12], list(range(13))):
            dst = torch.full(src.shape, 7.0, dtype=dtype, device="cuda")
            stage_states_h2d(dst, src, sel)
            torch.cuda.synchronize()
            exp = torch.full(src.shape, 7.0, dtype=dtype)
            exp[:, :, sel] = src[:, :, sel]
            assert torch.equal(dst.cpu(), exp), (dtype, sel)


@pytest.mark.parametrize("state_dtype", [torch.float32, torch.bfloat16])
def test_captured_step_equals_eager_step(state_dtype):
    from iisan_b200.engine import TrainStep
    from iisan_b200.optim import FusedAdam
    from iisan_b200.precision import set_compute_mode
    B = 16
    cfg, ids_np, lm_np, params, pop = _setup(B, 150, 9)
    (ids, image, text), = _host_batches(1, B, ids_np, state_dtype, 3)
    dev = [t.cuda() for t in (ids.view(-1), image, text, torch.from_numpy(lm_np))]
    set_compute_mode("bf16")
    try:
        losses = {}
        finals = {}
        for kind in ("eager", "graph"):
            model = build_product(cfg, params, pop).eval()
            opt = FusedAdam(model.parameters(), lr=1e-3)
            step = TrainStep(model, opt, use_graph=(kind == "graph"))
            losses[kind] = [float(step(*dev).item()) for _ in range(4)]
            finals[kind] = {n: p.detach().clone() for n, p in model.named_parameters()}
        e, g = losses["eager"], losses["graph"]
        assert e[0] == g[0], losses                     # the capture's warm-up steps were rolled back: exactly ONE step applied
        assert np.allclose(e[1], g[1], rtol=1e-4), losses
        assert np.allclose(e, g, rtol=2e-2), losses
        assert e[-1] < e[0] and g[-1] < g[0], losses
        worst = max(float((finals["eager"][n] - finals["graph"][n]).norm() / (finals["eager"][n].norm() + 1e-12)) for n in finals["eager"])
        assert worst < 2e-2, worst
    finally:
        set_compute_mode(None)


# ----------------------------------------------------------------------------------------------------------------------
# whole step at the benchmarked size against the oracle (0.25 s per oracle forward on the host cores)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["dense", "realistic"])
def test_full_size_step_matches_oracle(mode):
    from iisan_b200.ops import inbatch_ce_masks
    from iisan_b200.precision import set_compute_mode
    from oracle import iisan_oracle as O
    B, item_num = 512, 19246
    cfg, ids_np, lm_np, params, pop = _setup(B, item_num, 21, mode)
    g = torch.Generator().manual_seed(21)
    image = torch.randn(B, 11, 13, 768, generator=g).bfloat16()
    text = torch.randn(B, 11, 13, 768, generator=g).bfloat16()
    pad = torch.from_numpy(ids_np) == 0
    image[pad] = 0
    text[pad] = 0
    batch = {"ids": ids_np, "log_mask": lm_np, "image": image.float().numpy(), "text": text.float().numpy()}
    with torch.no_grad():
        ref = O.model_forward(O.params_to_torch(params, requires_grad=False), batch, pop, cfg)
    ref_loss = float(ref["loss"])
    ids = torch.from_numpy(ids_np).cuda().view(-1)
    lm = torch.from_numpy(lm_np).cuda()
    E = cfg.embedding_dim
    # masks / labels / valid rows: bit-exact at the full size (both CE implementations)
    rows = O.valid_rows(lm_np)
    L = cfg.max_seq_len
    rej = O.reject_mask(ids_np, ids_np, L)
    colm = np.broadcast_to(~O.column_valid(lm_np), rej.shape)
    lab = O.ce_labels(B, L)
    bits = inbatch_ce_masks(ids.view(B, 11), ids.view(B, 11), lm, lm).cpu().numpy()            # exact-mode CE kernel
    assert np.array_equal(np.nonzero(bits[:, 0] & 8)[0], rows)
    assert np.array_equal((bits & 1) != 0, colm)
    assert np.array_equal((bits & 2) != 0, rej)
    assert np.array_equal(np.argmax((bits & 4) != 0, axis=1), lab)
    bits = inbatch_ce_masks(ids.view(B, 11), ids.view(B, 11), lm, lm, fast=True).cpu().numpy() # tensor-core CE kernel
    expect = rej | colm
    expect[np.arange(B * L), lab] = colm[np.arange(B * L), lab]
    assert np.array_equal(np.nonzero(bits[:, 0] & 8)[0], rows)
    assert np.array_equal((bits & 1) != 0, expect)
    assert np.array_equal(np.argmax((bits & 4) != 0, axis=1), lab)
    del bits, expect
    for cm, dt, loss_tol, emb_tol in (("fp32", torch.float32, 1e-5, 1e-4), ("bf16", torch.bfloat16, 1e-2, 1e-2)):
        set_compute_mode(cm)
        try:
            model = build_product(cfg, params, pop).eval()
            im = image.to(device="cuda", dtype=dt); tx = text.to(device="cuda", dtype=dt)
            with torch.no_grad():
                loss = float(model(ids, im, tx, lm, 0).item())
                score = model.item_embeddings(im, tx)
                prec = model.user_encoder(score.view(-1, 11, E)[:, :-1], lm, 0).reshape(-1, E)
            assert abs(loss - ref_loss) <= loss_tol * abs(ref_loss), (cm, loss, ref_loss)
            for name, a, b in (("score_embs", score, ref["score_embs"]), ("prec_vec", prec, ref["prec_vec"])):
                a = a.float().cpu().numpy(); b = b.numpy()
                err = np.abs(a - b).max() / np.abs(b).max()
                assert err <= emb_tol, (cm, name, err)
            print(f"[{mode}/{cm}] B=512 loss {loss:.6f} vs oracle {ref_loss:.6f}")
        finally:
            set_compute_mode(None)
