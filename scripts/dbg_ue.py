import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_gpu_user_encoder import _encoder, _inputs
from iisan_b200.precision import set_compute_mode
set_compute_mode("fp32")
for p in (0.0, 0.25):
    enc = _encoder(p).train()
    x, lm = _inputs(8)
    te = enc.transformer_encoder
    w = torch.randn(8, 10, 64, device="cuda")
    def loss_at(xv):
        te._step_dev = torch.full((1,), 6, dtype=torch.int64, device="cuda")
        return (enc(xv[:, :-1], lm, "cuda") * w).sum()
    xs = x.clone().requires_grad_(True)
    loss_at(xs).backward()
    torch.manual_seed(5)
    for trial in range(3):
        d = torch.randn_like(x); d[:, -1] = 0
        ana = (xs.grad * d).sum().item()
        for eps in (3e-2, 1e-2, 3e-3, 1e-3):
            num = (loss_at(x + eps * d).item() - loss_at(x - eps * d).item()) / (2 * eps)
            print(f"p={p} trial {trial} eps {eps}: num {num:.4f} ana {ana:.4f}")
