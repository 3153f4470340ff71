"""Third-generation path (resident-state chain forward san_chain3.cu + low-rank adjoint backward san_lr.cu) at the stage counts
and ragged row counts its eligibility rule admits but BASELINE's configurations do not exercise (A = 3 / 4 / 5 / 7 stages, N not a
multiple of the 128-row tile; the width is the 768 of the Code_Cached tree, san_lr_eligible), and the SAN alone against the
second generation at the benchmark size.

References: the fp32 oracle (oracle.train_step_grads; CC/model/model.py:300-349) for loss / embeddings, and its rounding-point
emulation (tests/bf16_emulation.py) for the gradients, with the same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,lists,B,seed", [(768, "1,3,5", 9, 11), (768, "2,5,8,11", 23, 12), (768, "1,3,5,7,9,11", 12, 13), (768, "0,4", 13, 14)])
def test_lowrank_path_other_widths_and_stage_counts(d, lists, B, seed):
    from bf16_emulation import lr_path, train_step_grads_emul
    from iisan_b200.precision import set_compute_mode
    from oracle import iisan_oracle as O
    from oracle.synthetic import PathConfig, make_batch, make_params, make_pop_prob
    from product_util import build_product, run_step
    cfg = PathConfig(item_num=300, d_img=d, d_text=d, vit_list=lists, bert_list=lists)
    assert lr_path(cfg, True)
    batch = make_batch(B, cfg, seed, "dense")
    params = make_params(cfg, seed, perturb=True)
    rng = np.random.default_rng(seed)
    for k in params:                       # adapters away from their 0.01 init, gates away from 0.5: the low-rank terms must matter
        if "adapter_list" in k and k.endswith("weight"):
            params[k] = (params[k] * 8.0).astype(np.float32)
        if "side_gate" in k:
            params[k] = rng.uniform(-0.12, 0.12, size=params[k].shape).astype(np.float32)
    pop = make_pop_prob(cfg, seed)
    rb = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    batch = dict(batch, image=rb(batch["image"]), text=rb(batch["text"]))
    ref_out, ref_grads = O.train_step_grads(params, batch, pop, cfg)
    emu_out, emu_grads = train_step_grads_emul(params, batch, pop, cfg, ce_bf16=True, fused_chain=True)
    set_compute_mode("bf16")
    try:
        model = build_product(cfg, params, pop).eval()
        loss, grads = run_step(model, batch, dtype=torch.bfloat16)
    finally:
        set_compute_mode(None)
    ref_loss = float(ref_out["loss"])
    assert abs(float(loss) - ref_loss) <= 1e-2 * abs(ref_loss), (float(loss), ref_loss)
    assert abs(float(loss) - float(emu_out["loss"])) <= 1e-3 * abs(ref_loss)
    errs, gate_ref, gate_got = [], [], []
    for n, g in emu_grads.items():
        if g is None:
            continue
        err = float(np.linalg.norm((grads[n] - g).astype(np.float64)) / (np.linalg.norm(g.astype(np.float64)) + 1e-30))
        if g.size == 1:
            gate_ref.append(float(g.ravel()[0])); gate_got.append(float(grads[n].ravel()[0]))
        else:
            assert err <= 5e-2, f"{n}: {err}"
            errs.append(err)
    gate_err = np.linalg.norm(np.array(gate_got) - np.array(gate_ref)) / np.linalg.norm(gate_ref)
    assert gate_err <= 0.10, gate_err
    assert float(np.median(errs)) <= 2e-2, np.median(errs)
    print(f"d={d} A={len(lists.split(',')) + 1} N={B * 11}: loss {float(loss):.5f} (fp32 {ref_loss:.5f}), worst / median grad err {max(errs):.2e} / {np.median(errs):.2e}, gates {gate_err:.2e}")


def test_lowrank_path_matches_generation_2_at_benchmark_size():
    """SAN alone, B = 512 (N = 5632 items, 44 row tiles x 3 towers): embeddings and every parameter gradient of generation 3 against
    generation 2 (which keeps the width-d stashes and differentiates stage by stage) on the same inputs."""
    import bench
    from iisan_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    model, _, _ = bench.build_model(dev, "bf16")
    san = model.mm_encoder.eval()
    g = torch.Generator(device=dev).manual_seed(5)
    with torch.no_grad():
        for n, p in san.named_parameters():
            if "side_gate" in n:
                p.copy_((torch.rand(p.shape, device=dev, generator=g) - 0.5) * 0.3)
            elif "adapter_list" in n and n.endswith("weight"):
                p.mul_(4.0)
            elif n.endswith("bias"):
                p.add_(torch.randn(p.shape, device=dev, generator=g) * 0.05)
    N = 5632
    img = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    txt = torch.randn(N, 13, 768, device=dev, generator=g).bfloat16()
    w = torch.randn(N, 192, device=dev, generator=g)
    res = {}
    prev = lib.iisan_debug_chain_generation(0)
    try:
        for gen in (2, 3):
            lib.iisan_debug_chain_generation(gen)
            san.zero_grad(set_to_none=True)
            out = san.embed(img, txt)
            (out * w).sum().backward()
            torch.cuda.synchronize()
            res[gen] = (out.detach().clone(), {n: p.grad.detach().clone() for n, p in san.named_parameters() if p.grad is not None})
    finally:
        lib.iisan_debug_chain_generation(prev)
    o2, g2 = res[2]; o3, g3 = res[3]
    assert torch.isfinite(o3).all()
    assert float((o2 - o3).norm() / o2.norm()) <= 5e-3
    gates2, gates3 = [], []
    for n in g2:
        assert n in g3, n
        if g2[n].numel() == 1:
            gates2.append(float(g2[n])); gates3.append(float(g3[n]))
        else:
            err = float((g2[n] - g3[n]).norm() / (g2[n].norm() + 1e-30))
            assert err <= 2e-2, (n, err)
    ge = np.linalg.norm(np.array(gates2) - np.array(gates3)) / np.linalg.norm(gates2)
    assert ge <= 2e-2, ge
