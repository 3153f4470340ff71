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


def test_fused_adam_checkpoint_handover_with_torch_adam():
    """an optimizer state saved by torch.optim.Adam (the reference's epoch-N.pt) continues under FusedAdam and
    vice versa, with the same trajectory as an uninterrupted torch.optim.Adam run."""
    from iisan_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 192), (64,), (768, 64), (1,)]
    init = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    grads = [[torch.randn(s, device="cuda", generator=g) for s in shapes] for _ in range(6)]

    def run(opt, params, steps):
        for k in steps:
            for p, gr in zip(params, grads[k]):
                p.grad = gr.clone()
            opt.step()

    ref = [p.clone().requires_grad_(True) for p in init]
    o_ref = torch.optim.Adam(ref, lr=1e-3)
    run(o_ref, ref, range(6))
    # torch -> fused
    a = [p.clone().requires_grad_(True) for p in init]
    o_a = torch.optim.Adam(a, lr=1e-3)
    run(o_a, a, range(3))
    b = [p.detach().clone().requires_grad_(True) for p in a]
    o_b = FusedAdam(b, lr=1e-3)
    o_b.load_state_dict(o_a.state_dict())
    run(o_b, b, range(3, 6))
    for x, y in zip(ref, b):
        assert torch.allclose(x, y, rtol=2e-5, atol=2e-7)
    # fused -> torch
    c = [p.clone().requires_grad_(True) for p in init]
    o_c = FusedAdam(c, lr=1e-3)
    run(o_c, c, range(3))
    d = [p.detach().clone().requires_grad_(True) for p in c]
    o_d = torch.optim.Adam(d, lr=1e-3)
    o_d.load_state_dict(o_c.state_dict())
    run(o_d, d, range(3, 6))
    for x, y in zip(ref, d):
        assert torch.allclose(x, y, rtol=2e-5, atol=2e-7)
