"""Helpers shared by the parity tests: load the frozen reference outputs (tests/golden/*.npz, made by
oracle/make_golden.py from the reference's own model package) and compare gradients to their digests."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ALL = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("eval_", "data_")))   # eval_*: tests/test_eval.py, data_*: tests/test_store_cpu.py
ORACLE_ONLY_CASES = [c for c in _ALL if c.startswith("oracle_")]  # pin the oracle only (CPU suite): options whose product parity is tested against the oracle
_ALL = [c for c in _ALL if not c.startswith("oracle_")]
CASES = [c for c in _ALL if not c.startswith("versa_")]          # small widths: every parity test runs on all of them
VERSA_CASES = [c for c in _ALL if c.startswith("versa_")]         # BASELINE configs[3]/[4] at their real widths / layer counts


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def rebuild_inputs(meta):
    from oracle.synthetic import PathConfig, make_batch, make_params, make_pop_prob
    cfg = PathConfig(**meta["cfg"])
    batch = make_batch(meta["B"], cfg, meta["seed"], meta["mode"])
    params = make_params(cfg, meta["seed"], perturb=True)
    pop = make_pop_prob(cfg, meta["seed"])
    return cfg, batch, params, pop


def golden_masked(z):
    shape = tuple(int(v) for v in z["masked_shape"])
    bits = np.unpackbits(z["masked_bits"])[: shape[0] * shape[1]].astype(bool)
    return bits.reshape(shape)


def check_grads(z, grads, rtol, atol_scale=1e-6, names=None):
    """grads: name -> np.ndarray | None.  Compares against the digest (norm, sum, strided sample)."""
    worst = 0.0
    for key in z.files:
        if key.startswith("gradnone/"):
            n = key[len("gradnone/"):]
            assert grads.get(n) is None or not np.any(grads[n]), n
            continue
        if not key.endswith("/sample"):
            continue
        n = key[len("grad/"):-len("/sample")]
        if names is not None and n not in names:
            continue
        g = grads[n]
        assert g is not None, f"missing grad for {n}"
        flat = np.asarray(g, dtype=np.float32).reshape(-1)
        step = int(z[f"grad/{n}/step"]); ref = z[key]
        got = flat[::step]
        norm = float(z[f"grad/{n}/norm"])
        scale = max(norm / np.sqrt(max(flat.size, 1)), 1e-12)       # rms of the reference gradient
        err = np.abs(got - ref).max() / (np.abs(ref).max() + atol_scale * scale + 1e-30)
        worst = max(worst, float(err))
        assert np.allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max() + 1e-12), \
            f"{n}: max abs err {np.abs(got - ref).max():.3e} vs max |ref| {np.abs(ref).max():.3e}"
        got_norm = float(np.linalg.norm(flat.astype(np.float64)))
        assert abs(got_norm - norm) <= rtol * max(norm, 1e-12) * 4 + 1e-12, f"{n}: norm {got_norm} vs {norm}"
    return worst
