"""Fused multi-tensor Adam (iisan_adam_step) against torch.optim.Adam on the same gradients, with per-group learning rates."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch():
    from iisan_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(768, 768), (64, 768), (768,), (1,), (256, 64), (10, 64), (5000,), (3, 5, 7)] * 25        # 200 tensors: two launches (IISAN_ADAM_MAX_TENSORS = 192)
    ref = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    ours = [p.detach().clone().requires_grad_(True) for p in ref]
    lrs = [2e-4, 1e-4, 5e-5]
    groups = lambda ps: [{"params": ps[i::3], "lr": lrs[i]} for i in range(3)]
    o_ref = torch.optim.Adam(groups(ref))
    o_ours = FusedAdam(groups(ours))
    for step in range(6):
        for a, b in zip(ref, ours):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (0.1 + step)
            a.grad = gr.clone(); b.grad = gr.clone()
        o_ref.step(); o_ours.step()
    for a, b in zip(ref, ours):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), (a - b).abs().max().item()
