import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from golden_util import CASES, load_case, rebuild_inputs
from product_util import build_product, run_step
from iisan_b200.precision import set_compute_mode
from oracle import iisan_oracle as O
from bf16_emulation import train_step_grads_emul
for name in CASES:
    z, meta = load_case(name)
    cfg, batch, params, pop = rebuild_inputs(meta)
    emu_out, emu_grads = train_step_grads_emul(params, batch, pop, cfg)
    set_compute_mode("bf16")
    model = build_product(cfg, params, pop).eval()
    loss, grads = run_step(model, batch)
    errs = {}
    for n, g in emu_grads.items():
        if g is None: continue
        errs[n] = float(np.abs(grads[n] - g).max() / (np.abs(g).max() + 1e-30))
    v = np.array(sorted(errs.values()))
    print(name, "loss", float(loss), float(emu_out["loss"]), "median err %.2e p90 %.2e max %.2e  n>2e-2: %d of %d" % (np.median(v), v[int(0.9*len(v))], v[-1], (v > 2e-2).sum(), len(v)))
    for n, e in sorted(errs.items(), key=lambda kv: -kv[1])[:8]: print("   %.3e %s" % (e, n))
