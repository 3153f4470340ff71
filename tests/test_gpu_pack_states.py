"""States stored as the reference writes them (fp32 .pt files, Code_Cached/preprocess_vectors.py:27-31; fp16 for the LLaMA / EVA
caches) on the fast path: iisan_pack_states selects the layers the towers read and rounds them to bf16 exactly like torch does,
and a chain-eligible configuration fed fp32 states then runs the very same fused kernels as one fed the pre-rounded bf16 states."""
import numpy as np
import pytest
import torch

from product_util import build_product

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,sel", [((37, 13, 768), [0, 2, 4, 6, 8, 10, 12]), ((5, 11, 4, 256), [1, 3]), ((3, 81, 8192), [0, 40, 80])])
def test_pack_states_bit_exact(dtype, shape, sel):
    from iisan_b200.ops import pack_states
    g = torch.Generator().manual_seed(len(sel) + shape[-1])
    x = (torch.randn(*shape, generator=g) * 3.0).to(dtype)
    x.view(-1)[::97] = 0.0
    got = pack_states(x.cuda(), sel).cpu()
    ref = x.reshape(-1, shape[-2], shape[-1])[:, sel].to(torch.bfloat16)          # round-to-nearest-even, as the kernel
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


def test_fp32_states_take_the_fused_chain():
    """bf16 mode, Code_Cached configuration (d = 768, r = 64, 7 stages): fp32-stored states == the same states pre-rounded to bf16,
    loss bit-equal, gradients equal up to the order of the atomic sums; and the library reports the route."""
    import ctypes as C
    from iisan_b200 import _lib
    from iisan_b200.precision import compute_mode, set_compute_mode
    from oracle.synthetic import PathConfig, make_batch, make_params, make_pop_prob
    cfg = PathConfig(item_num=200)
    params = make_params(cfg, 5, perturb=True)
    pop = make_pop_prob(cfg, 5)
    batch = make_batch(12, cfg, 9, "realistic")
    set_compute_mode("bf16")
    try:
        out = {}
        for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
            model = build_product(cfg, params, pop).eval()
            ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
            img = torch.from_numpy(batch["image"]).cuda().to(dt); txt = torch.from_numpy(batch["text"]).cuda().to(dt)
            lm = torch.from_numpy(batch["log_mask"]).cuda()
            if name == "fp32":
                binder = model.mm_encoder._bind()
                desc = binder.desc(img, txt, compute_mode(), False)
                assert _lib.load().iisan_san_fused_eligible(C.byref(desc)) == 1
            model.zero_grad(set_to_none=True)
            loss = model(ids, img, txt, lm, "cuda")
            loss.backward()
            out[name] = (loss.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()})
        assert torch.equal(out["fp32"][0], out["bf16"][0]), (out["fp32"][0].item(), out["bf16"][0].item())
        for n, g in out["bf16"][1].items():
            assert torch.allclose(out["fp32"][1][n], g, rtol=1e-3, atol=1e-6 + 1e-4 * g.abs().max().item()), n
    finally:
        set_compute_mode(None)
