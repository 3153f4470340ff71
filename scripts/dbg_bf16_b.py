import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from product_util import build_product, run_step
from iisan_b200.precision import set_compute_mode
from oracle.synthetic import PathConfig, make_batch, make_params, make_pop_prob
from bf16_emulation import train_step_grads_emul
cfg = PathConfig(item_num=2000)
for B, mode in ((16, "realistic"), (64, "realistic"), (256, "realistic"), (64, "dense")):
    batch = make_batch(B, cfg, 5, mode); params = make_params(cfg, 5, perturb=True); pop = make_pop_prob(cfg, 5)
    emu_out, emu_grads = train_step_grads_emul(params, batch, pop, cfg)
    set_compute_mode("bf16")
    model = build_product(cfg, params, pop).eval()
    loss, grads = run_step(model, batch)
    errs = {}; l2 = {}
    for n, g in emu_grads.items():
        if g is None: continue
        errs[n] = float(np.abs(grads[n] - g).max() / (np.abs(g).max() + 1e-30))
        l2[n] = float(np.linalg.norm(grads[n] - g) / (np.linalg.norm(g) + 1e-30))
    v = np.array(sorted(errs.values())); w = np.array(sorted(l2.values()))
    print(B, mode, "loss", float(loss), float(emu_out["loss"]), "max-norm: median %.2e p90 %.2e max %.2e | l2: median %.2e p90 %.2e max %.2e" % (np.median(v), v[int(0.9*len(v))], v[-1], np.median(w), w[int(0.9*len(w))], w[-1]))
    for n, e in sorted(errs.items(), key=lambda kv: -kv[1])[:5]: print("   %.3e l2 %.3e %s" % (e, l2[n], n))
