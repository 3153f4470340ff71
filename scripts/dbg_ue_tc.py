"""Per-tensor error of the fused user encoder (exact FMA mode and TF32 tensor-core mode) against the oracle."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from test_gpu_user_encoder import _encoder, _inputs
from iisan_b200.precision import set_compute_mode
from oracle import iisan_oracle as O
from oracle.synthetic import PathConfig
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-12))
for B in (1, 5, 64, 512):
    for mode in ("fp32", "bf16"):
        set_compute_mode(mode)
        enc = _encoder(0.0).eval()
        x, lm = _inputs(B)
        xs = x.clone().requires_grad_(True)
        out = enc(xs[:, :-1], lm, "cuda")
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda()
        (out * w).sum().backward()
        P = {"user_encoder." + n: p.detach().cpu().clone().requires_grad_(True) for n, p in enc.named_parameters()}
        xr = x.cpu().clone().requires_grad_(True)
        if mode == "bf16" and os.environ.get("EMUL", "1") == "1":
            import bf16_emulation as EM
            ref = EM.user_encoder_forward_emul(P, xr[:, :-1], lm.cpu(), PathConfig())
        else:
            ref = O.user_encoder_forward(P, xr[:, :-1], lm.cpu(), PathConfig())
        (ref * w.cpu()).sum().backward()
        errs = {n.replace("transformer_encoder.", "").replace("transformer_blocks.", "b"): rel(p.grad.cpu(), P["user_encoder." + n].grad) for n, p in enc.named_parameters()}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"B={B} {mode}: out {rel(out.cpu(), ref):.2e} dx {rel(xs.grad.cpu(), xr.grad):.2e} worst {[(k, f'{v:.1e}') for k, v in worst]}")
set_compute_mode("fp32")
