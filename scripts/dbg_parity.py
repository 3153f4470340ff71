"""Per-tensor relative L2 error of the fast mode against the rounding-point emulation (with / without TF32 user encoder)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import bf16_emulation as EM
from golden_util import CASES, load_case, rebuild_inputs
from product_util import build_product, run_step
from iisan_b200.precision import set_compute_mode
from oracle import iisan_oracle as O
names = sys.argv[1:] or CASES[:2]
for name in names:
    z, meta = load_case(name)
    cfg, batch, params, pop = rebuild_inputs(meta)
    set_compute_mode("bf16")
    model = build_product(cfg, params, pop).eval()
    loss, grads = run_step(model, batch, dtype=torch.float32)
    set_compute_mode(None)
    for label, ue in (("tf32-emul", EM.user_encoder_forward_emul), ("fp32-ue", O.user_encoder_forward)):
        saved = EM.user_encoder_forward_emul
        EM.user_encoder_forward_emul = ue
        emu_out, emu_grads = EM.train_step_grads_emul(params, batch, pop, cfg, ce_bf16=(cfg.embedding_dim == 64), fused_chain=False)
        EM.user_encoder_forward_emul = saved
        errs = {}
        for n, g in emu_grads.items():
            if g is None or g.size == 1: continue
            errs[n] = float(np.linalg.norm((grads[n] - g).astype(np.float64)) / (np.linalg.norm(g.astype(np.float64)) + 1e-30))
        ue_e = [v for k, v in errs.items() if k.startswith("user_encoder")]
        san_e = [v for k, v in errs.items() if not k.startswith("user_encoder")]
        print(f"{name} {label}: loss {float(loss):.6f} vs {float(emu_out['loss']):.6f}; UE tensors median {np.median(ue_e):.2e} max {max(ue_e):.2e}; "
              f"others median {np.median(san_e):.2e} max {max(san_e):.2e}; all median {np.median(list(errs.values())):.2e}")
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
        print("   worst:", [(k[-45:], f"{v:.1e}") for k, v in worst])
