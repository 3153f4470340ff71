"""The low-rank adjoint of the side-adapter network (tests/lowrank_reference.py, the algebra of the third-generation chain
path) against autograd of the oracle's own forward (oracle.san_forward, CC/model/model.py:300-349), in float64."""
import numpy as np
import pytest
import torch

from oracle import iisan_oracle as O
from oracle.synthetic import PathConfig, make_params
from lowrank_reference import tower_backward


@pytest.mark.parametrize("seed,scale", [(1, 1.0), (2, 30.0)])
def test_lowrank_adjoint_matches_autograd(seed, scale):
    cfg = PathConfig(item_num=50, d_img=32, d_text=32, r_cv=8, r_bert=8, embedding_dim=8)
    Pn = make_params(cfg, seed)
    rng = np.random.default_rng(seed)
    for k in Pn:                                  # adapters far from their N(0, 0.01^2) init, gates away from 0.5
        if "adapter_list" in k and k.endswith("weight"):
            Pn[k] = Pn[k] * scale
        if "side_gate" in k:
            Pn[k] = rng.uniform(-0.15, 0.15, size=Pn[k].shape).astype(np.float32)
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in Pn.items()}
    N = 37
    image = torch.tensor(rng.standard_normal((N, 13, 32)).astype(np.float32))      # the oracle reads the states as fp32
    text = torch.tensor(rng.standard_normal((N, 13, 32)).astype(np.float32))
    e_cv, e_tx, e_mm = O.san_forward(P, image, text, cfg)
    wts = [torch.tensor(rng.standard_normal((N, 8))) for _ in range(3)]
    loss = (e_cv * wts[0]).sum() + (e_tx * wts[1]).sum() + (e_mm * wts[2]).sum()
    loss.backward()

    plan = O.stage_plan(cfg)
    image, text = image.double(), text.double()
    m = "mm_encoder."
    towers = {
        "cv": ("intra", "cv_adapter_list", "side_gate_params_cv", "fc_cv", "cv_pre_fc", e_cv, wts[0]),
        "text": ("intra", "bert_adapter_list", "side_gate_params_text", "fc_bert", "bert_pre_fc", e_tx, wts[1]),
        "mm": ("mm", "mm_adapter_list", "side_gate_params_mm", "fc_mm", "fc_mm_down", e_mm, wts[2]),
    }
    for name, (kind, ad, gate, fc, pre, y_ref, e) in towers.items():
        with torch.no_grad():
            idx = [(p[0], p[1], p[2], p[3], p[4]) for p in plan]
            if name == "text":
                h = [text[:, p[1]] for p in idx]; h2 = None; a_ix = [p[0] for p in idx]
            elif name == "cv":
                h = [image[:, p[3]] for p in idx]; h2 = None; a_ix = [p[2] for p in idx]
            else:
                h = [image[:, p[3]] for p in idx]; h2 = [text[:, p[1]] for p in idx]; a_ix = [p[4] for p in idx]
            g = lambda k: P[m + k].detach()
            y, G = tower_backward(kind, h, h2, [g(f"{gate}.{a}") for a in a_ix],
                                  [g(f"{ad}.{a}.fc_down.weight") for a in a_ix], [g(f"{ad}.{a}.fc_down.bias") for a in a_ix],
                                  [g(f"{ad}.{a}.fc_up.weight") for a in a_ix], [g(f"{ad}.{a}.fc_up.bias") for a in a_ix],
                                  g(fc + ".weight"), g(fc + ".bias"), g(pre + ".weight"), g(pre + ".bias"), e)
        assert torch.allclose(y, y_ref.detach(), rtol=1e-10, atol=1e-10)

        def chk(key, val):
            ref = P[m + key].grad
            err = (val.reshape(ref.shape) - ref).abs().max() / (ref.abs().max() + 1e-30)
            assert err < 1e-8, (name, key, float(err))
        for s, a in enumerate(a_ix):
            chk(f"{ad}.{a}.fc_down.weight", G["Wd"][s]); chk(f"{ad}.{a}.fc_down.bias", G["bd"][s])
            chk(f"{ad}.{a}.fc_up.weight", G["Wu"][s]); chk(f"{ad}.{a}.fc_up.bias", G["bu"][s])
            chk(f"{gate}.{a}", G["gate"][s])
        chk(fc + ".weight", G["W_fc"]); chk(fc + ".bias", G["b_fc"])
        chk(pre + ".weight", G["W_pre"]); chk(pre + ".bias", G["b_pre"])
