"""TEST INFRASTRUCTURE -- freeze outputs of the *reference's own* model package as golden fixtures.

Run in the build container only (``/root/reference`` is not present on the GPU box):

    python -m oracle.make_golden            # writes tests/golden/*.npz

For every case it imports ``/root/reference/Code_Cached{,_Asym}/model`` (one tree per subprocess,
because both own the top-level module name ``model``), instantiates ``ModelMM`` +
``IISANAdaptedMModel`` exactly as ``Code_Cached/run.py:138,182-183`` does (stub ``image_net`` with a
``classifier`` Linear, ``bert_model = nn.Identity()``), loads the seeded parameters from
``oracle.synthetic.make_params``, runs forward + backward on the seeded batch on CPU in fp32 and
stores: loss, item/user embeddings, the masked-logit pattern and labels seen by the criterion, and
for every parameter gradient its l2 norm, sum and a strided sample (full tensor when small).
Nothing from the reference is copied; only numbers it computed are stored.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (tree, PathConfig kwargs, B, batch mode, seed)
    "cc_dense_b4": ("Code_Cached", {}, 4, "dense", 101),
    "cc_real_b8": ("Code_Cached", {}, 8, "realistic", 202),
    "cc_real_b16_smallcat": ("Code_Cached", {"item_num": 40}, 16, "realistic", 303),
    "asym_text_wide": ("Code_Cached_Asym", dict(asym=True, d_text=96, d_img=64, layers_text=9, layers_img=5,
                                                bert_list="1,3,5,7", vit_list="1,3", r_cv=16, r_bert=24,
                                                embedding_dim=32, item_num=500), 6, "realistic", 404),
    "asym_img_wide": ("Code_Cached_Asym", dict(asym=True, d_text=64, d_img=128, layers_text=5, layers_img=7,
                                               bert_list="1,3", vit_list="0,2,3,5", r_cv=24, r_bert=16,
                                               embedding_dim=32, item_num=500), 6, "realistic", 505),
    "asym_equal": ("Code_Cached_Asym", dict(asym=True, d_text=64, d_img=64, layers_text=7, layers_img=7,
                                            bert_list="1,3,5", vit_list="1,3,5", r_cv=16, r_bert=16,
                                            embedding_dim=32, item_num=500), 5, "dense", 606),
    # --- args.adapter_activation == "GELU" (CC/model/modules.py:104-107; parameters.py:72 defaults to RELU), both trees ---
    "cc_gelu_b6": ("Code_Cached", {"adapter_activation": "GELU"}, 6, "realistic", 909),
    "asym_gelu_text_wide": ("Code_Cached_Asym", dict(asym=True, d_text=96, d_img=64, layers_text=9, layers_img=5,
                                                     bert_list="1,3,5,7", vit_list="1,3", r_cv=16, r_bert=24, embedding_dim=32,
                                                     item_num=500, adapter_activation="GELU"), 5, "realistic", 919),
    # --- remove_first == "TRUE" (CC/model/model.py:264-265, 305-308: the towers start from cached layer 0 and -- quirk Q1 -- the
    #     cv layer list is read from side_adapter_bert_list; CA/model/model.py:265-270 reads each modality's own string).  The
    #     "oracle_" prefix keeps these out of the GPU case lists (tests/golden_util.py): they pin the ORACLE's restatement of the
    #     option; the product is compared with the oracle on it in tests/test_gpu_store.py ---
    "oracle_cc_remove_first": ("Code_Cached", {"remove_first": "TRUE", "bert_list": "0,2,4,6,8,10", "vit_list": "1,3,5,7,9,11"},
                               5, "realistic", 929),
    "oracle_asym_remove_first": ("Code_Cached_Asym", dict(asym=True, remove_first="TRUE", d_text=96, d_img=64, layers_text=9, layers_img=5,
                                                          bert_list="1,3,5,7", vit_list="0,2", r_cv=16, r_bert=24, embedding_dim=32,
                                                          item_num=500), 5, "realistic", 939),
    # --- BASELINE.json configs[3] / configs[4] at their REAL widths and layer counts (small B: the widths, layer pitches,
    #     stage plans and the dim-alignment GEMM are what these cases pin; tests/test_gpu_versa_shapes.py) ---
    # BERT-large text + ViT-large image, group layer-drop 13 text vs 7 image adapters: 6 text-only stages, then 7 paired
    # (list options CA/script/run_IISAN.py:48-53)
    "versa_bertlarge_vitlarge": ("Code_Cached_Asym", dict(asym=True, d_text=1024, d_img=1024, layers_text=25, layers_img=25,
                                                          bert_list="1,3,5,7,9,11,13,15,17,19,21,23", vit_list="1,3,5,7,9,11",
                                                          r_cv=64, r_bert=64, embedding_dim=64, item_num=22785), 3, "realistic", 707),
    # the same widths with equal adapter counts (list option "3,7,11,15,19,23", CA/script/run_IISAN.py:49): with bf16 states this is
    # the configuration the fused chain kernels serve at d = 1024
    "versa_large_sym": ("Code_Cached_Asym", dict(asym=True, d_text=1024, d_img=1024, layers_text=25, layers_img=25,
                                                 bert_list="3,7,11,15,19,23", vit_list="3,7,11,15,19,23",
                                                 r_cv=64, r_bert=64, embedding_dim=64, item_num=22785), 4, "realistic", 717),
    # BERT-large text + ViT-base image: group layer-drop AND down_project 1024 -> 768 (CA/model/model.py:406-411)
    "versa_bertlarge_vitbase": ("Code_Cached_Asym", dict(asym=True, d_text=1024, d_img=768, layers_text=25, layers_img=13,
                                                         bert_list="1,3,5,7,9,11,13,15,17,19,21,23", vit_list="1,3,5,7,9,11",
                                                         r_cv=64, r_bert=64, embedding_dim=64, item_num=22785), 3, "realistic", 808),
    # LLaMA-3-70B-shaped text [81, 8192] + EVA-CLIP-18B-shaped image [49, 5120], down_project 8192 -> 5120
    # (CA/script/run_IISAN_eva.py:56-65)
    "versa_llama70b_evaclip": ("Code_Cached_Asym", dict(asym=True, d_text=8192, d_img=5120, layers_text=81, layers_img=49,
                                                        bert_list="4,19,34,49,64,79", vit_list="2,11,20,29,38,47",
                                                        r_cv=64, r_bert=64, embedding_dim=64, item_num=22785), 2, "dense", 909),
}

SAMPLE_MAX = 2048


def grad_digest(g: np.ndarray):
    flat = g.reshape(-1).astype(np.float32)
    step = max(1, flat.size // SAMPLE_MAX)
    return {"norm": np.float64(np.linalg.norm(flat.astype(np.float64))),
            "sum": np.float64(flat.astype(np.float64).sum()),
            "step": np.int64(step), "sample": flat[::step].copy()}


def run_case(name: str):
    sys.dont_write_bytecode = True
    tree, kw, B, mode, seed = CASES[name]
    sys.path.insert(0, ROOT)
    from oracle.synthetic import PathConfig, make_args, make_batch, make_params, make_pop_prob
    cfg = PathConfig(**kw)
    sys.path.insert(0, os.path.join("/root/reference", tree))
    import torch
    from torch import nn
    from model.model import ModelMM, IISANAdaptedMModel          # the reference's own code

    torch.manual_seed(0)
    args = make_args(cfg)

    class ImgStub(nn.Module):                                    # stands in for ViTForImageClassification
        def __init__(self):
            super().__init__()
            self.classifier = nn.Linear(cfg.d_img, cfg.embedding_dim)

    pop = make_pop_prob(cfg, seed)
    m = ModelMM(args, cfg.item_num, True, ImgStub(), nn.Identity(), pop)
    m.mm_encoder = IISANAdaptedMModel(m.mm_encoder, args)       # Code_Cached/run.py:182-183
    params = make_params(cfg, seed, perturb=True)
    ref_names = [n for n, _ in m.named_parameters()]
    assert ref_names == list(params.keys()), (set(ref_names) ^ set(params.keys()), ref_names[:5], list(params)[:5])
    with torch.no_grad():
        for n, p in m.named_parameters():
            assert tuple(p.shape) == params[n].shape, (n, p.shape, params[n].shape)
            p.copy_(torch.from_numpy(params[n]))
    m.eval()                                                     # dropout off (SAN has none; SASRec p=drop_rate)

    seen = {}
    m.criterion.register_forward_hook(
        lambda mod, i, o: seen.update(logits=i[0].detach().clone(), labels=i[1].detach().clone()))

    # capture score_embs / prec_vec through forward hooks on the sub-modules
    cap = {}
    m.com_dense.register_forward_hook(lambda mod, i, o: cap.__setitem__("score_embs", o.detach().clone()))
    m.user_encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("prec_vec", o.detach().clone()))
    m.mm_encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("san", (o[0].detach().clone(), o[1][0].detach().clone(), o[1][1].detach().clone())))

    b = make_batch(B, cfg, seed, mode)
    ids = torch.from_numpy(b["ids"]).view(-1)
    loss = m(ids, torch.from_numpy(b["image"]), torch.from_numpy(b["text"]), torch.from_numpy(b["log_mask"]), "cpu")
    loss.backward()

    out = {"loss": np.float32(loss.item()),
           "score_embs": cap["score_embs"].numpy(),
           "prec_vec": cap["prec_vec"].reshape(-1, cfg.embedding_dim).numpy(),
           "e_cv": cap["san"][0].numpy(), "e_text": cap["san"][1].numpy(), "e_mm": cap["san"][2].numpy(),
           "masked_bits": np.packbits((seen["logits"] == -1e4).numpy().reshape(-1)),
           "masked_shape": np.array(seen["logits"].shape, dtype=np.int64),
           "labels_valid": seen["labels"].numpy().astype(np.int64),
           "logits_valid_rowsum": seen["logits"].double().sum(1).numpy(),
           "meta": np.frombuffer(json.dumps({"tree": tree, "cfg": cfg.to_dict(), "B": B, "mode": mode,
                                             "seed": seed, "torch": torch.__version__}).encode(), dtype=np.uint8)}
    n_none = 0
    for n, p in m.named_parameters():
        if p.grad is None:
            n_none += 1
            out[f"gradnone/{n}"] = np.int64(1)
            continue
        d = grad_digest(p.grad.numpy())
        for k, v in d.items():
            out[f"grad/{n}/{k}"] = v
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: loss={loss.item():.6f} params={len(ref_names)} unused={n_none} "
          f"masked={int((seen['logits'] == -1e4).sum())}/{seen['logits'].numel()}")


def main():
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        return
    for name in CASES:
        subprocess.run([sys.executable, "-m", "oracle.make_golden", name], cwd=ROOT, check=True)


if __name__ == "__main__":
    main()
