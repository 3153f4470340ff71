"""TEST INFRASTRUCTURE -- seeded synthetic batches / parameters for the IISAN(Cached) hot path.

Deterministic numpy generators (PCG64) so the same inputs are rebuilt in the build container (where
the reference is imported to make golden fixtures) and on the GPU box (where only this repo exists).

Shapes follow the reference's train batch (Code_Cached/data_utils/dataset.py:65-92 and
Code_Cached/run.py:368-377): per user 11 item slots, left padded with id 0 / zero hidden states,
``log_mask = [0]*mask_len + [1]*(seq_len-1)``.
"""
from __future__ import annotations

import argparse
import math
from dataclasses import dataclass, field, asdict

import numpy as np

# Train-sequence length histogram of Dataset/Scientific ([probe]).
SCIENTIFIC_LEN_HIST = {3: 4892, 4: 2652, 5: 1465, 6: 895, 7: 609, 8: 356, 9: 300, 10: 208, 11: 699}


@dataclass
class PathConfig:
    """Static description of one IISAN(Cached) / IISAN-Versa configuration."""
    item_num: int = 20314           # Scientific catalogue size
    max_seq_len: int = 10           # parameters.py:34 ; run.py:373 hard-codes 11 slots
    embedding_dim: int = 64
    d_img: int = 768
    d_text: int = 768
    layers_img: int = 13            # cached states per item (n_layers + 1)
    layers_text: int = 13
    vit_list: str = "1,3,5,7,9,11"
    bert_list: str = "1,3,5,7,9,11"
    r_cv: int = 64
    r_bert: int = 64
    heads: int = 2
    blocks: int = 2
    drop_rate: float = 0.0
    asym: bool = False              # Code_Cached_Asym semantics
    remove_first: str = "None"
    adapter_activation: str = "RELU"

    @property
    def slots(self) -> int:
        return self.max_seq_len + 1

    def img_layers_selected(self):
        """Hidden-state indices the image tower reads.

        Code_Cached/model/model.py:263-268 (quirk Q1: without remove_first BOTH lists come from
        side_adapter_vit_list; with remove_first == "TRUE" the cv list comes from the bert string).
        Code_Cached_Asym/model/model.py:265-270 uses the matching string for each modality.
        """
        if self.asym:
            src = self.vit_list
        else:
            src = self.bert_list if self.remove_first == "TRUE" else self.vit_list
        base = [int(i) + 1 for i in src.split(",")]
        return base if self.remove_first == "TRUE" else [0] + base

    def text_layers_selected(self):
        src = self.bert_list if self.asym else self.vit_list
        base = [int(i) + 1 for i in src.split(",")]
        return base if self.remove_first == "TRUE" else [0] + base

    def to_dict(self):
        return asdict(self)


def make_args(cfg: PathConfig) -> argparse.Namespace:
    """The ``args`` namespace the reference constructors read."""
    ns = argparse.Namespace(
        max_seq_len=cfg.max_seq_len, l2_weight=0, embedding_dim=cfg.embedding_dim,
        num_attention_heads=cfg.heads, drop_rate=cfg.drop_rate, transformer_block=cfg.blocks,
        modality="intra_inter", news_attributes=["title"], num_words_title=30,
        num_words_abstract=50, num_words_body=50, word_embedding_dim=cfg.d_text,
        remove_first=cfg.remove_first, side_adapter_vit_list=cfg.vit_list,
        side_adapter_bert_list=cfg.bert_list, cv_adapter_down_size=cfg.r_cv,
        bert_adapter_down_size=cfg.r_bert, adapter_dropout_rate=0.1,
        adapter_activation=cfg.adapter_activation, fusion_method="gated")
    if cfg.asym:
        ns.text_embedding_dim = cfg.d_text
        ns.image_embedding_dim = cfg.d_img
        ns.text_layers = cfg.layers_text - 1
        ns.image_layers = cfg.layers_img - 1
    return ns


# ----------------------------------------------------------------------------------------------
# batches
# ----------------------------------------------------------------------------------------------

def make_ids(B: int, cfg: PathConfig, seed: int, mode: str = "dense", dup_prob: float = 0.2,
             cross_dup_prob: float = 0.3):
    """ids int64 [B,11] (left padded with 0) and log_mask float32 [B,10].

    ``dense``: all 11 slots valid.  ``realistic``: lengths from the Scientific histogram, some users
    get a duplicated item inside their own sequence and some share items with other users (both
    exercise the reject mask, Code_Cached/model/model.py:92-100).
    """
    rng = np.random.default_rng(seed)
    S = cfg.slots
    ids = np.zeros((B, S), dtype=np.int64)
    log_mask = np.zeros((B, S - 1), dtype=np.float32)
    if mode == "dense":
        lens = np.full(B, S, dtype=np.int64)
    elif mode == "realistic":
        ks = np.array(sorted(SCIENTIFIC_LEN_HIST)); ps = np.array([SCIENTIFIC_LEN_HIST[k] for k in ks], dtype=np.float64)
        lens = rng.choice(ks, size=B, p=ps / ps.sum())
        lens = np.minimum(lens, S)
    else:
        raise ValueError(mode)
    for u in range(B):
        n = int(lens[u])
        seq = rng.integers(1, cfg.item_num + 1, size=n)
        if mode == "realistic":
            if n >= 3 and rng.random() < dup_prob:
                a, b = rng.choice(n, size=2, replace=False)
                seq[a] = seq[b]
            if u > 0 and rng.random() < cross_dup_prob:
                v = int(rng.integers(0, u)); nv = int(lens[v])
                seq[int(rng.integers(0, n))] = ids[v, S - nv + int(rng.integers(0, nv))]
        ids[u, S - n:] = seq
        log_mask[u, S - n:] = 1.0          # positions mask_len .. 9  (n-1 ones)
    return ids, log_mask


def make_states(ids: np.ndarray, cfg: PathConfig, seed: int):
    """image [B,11,layers_img,d_img], text [B,11,layers_text,d_text] fp32 ~ N(0,1); padded slots zero."""
    rng = np.random.default_rng(seed + 7919)
    B, S = ids.shape
    image = rng.standard_normal((B, S, cfg.layers_img, cfg.d_img), dtype=np.float32)
    text = rng.standard_normal((B, S, cfg.layers_text, cfg.d_text), dtype=np.float32)
    pad = (ids == 0)
    image[pad] = 0.0
    text[pad] = 0.0
    return image, text


def make_pop_prob(cfg: PathConfig, seed: int) -> np.ndarray:
    """``[1] + normalised train counts`` (Code_Cached/data_utils/preprocess.py:77-82); counts >= 1."""
    rng = np.random.default_rng(seed + 104729)
    counts = np.floor(1.0 + rng.pareto(1.2, size=cfg.item_num) * 3.0).astype(np.float64)
    p = counts / counts.sum()
    return np.concatenate([[1.0], p]).astype(np.float32)


def make_batch(B: int, cfg: PathConfig, seed: int, mode: str = "dense"):
    ids, log_mask = make_ids(B, cfg, seed, mode)
    image, text = make_states(ids, cfg, seed)
    return {"ids": ids, "log_mask": log_mask, "image": image, "text": text}


# ----------------------------------------------------------------------------------------------
# parameters (names and shapes = reference state_dict)
# ----------------------------------------------------------------------------------------------

def _adapter_counts(cfg: PathConfig):
    a_img = len(cfg.img_layers_selected())
    a_text = len(cfg.text_layers_selected())
    if cfg.asym:
        # CA/model/model.py:279-287: mm adapters follow the *narrower* modality's list
        if cfg.d_text > cfg.d_img:
            a_mm = a_img
        else:
            a_mm = a_text
    else:
        a_mm = a_img
    return a_text, a_img, a_mm


def param_shapes(cfg: PathConfig):
    """Ordered ``name -> shape`` of every trainable tensor of ModelMM + IISANAdaptedMModel."""
    E, H = cfg.embedding_dim, cfg.embedding_dim * 4
    shp = {}
    te = "user_encoder.transformer_encoder."
    shp[te + "position_embedding.weight"] = (cfg.max_seq_len, E)
    shp[te + "layer_norm.weight"] = (E,); shp[te + "layer_norm.bias"] = (E,)
    for b in range(cfg.blocks):
        p = f"{te}transformer_blocks.{b}."
        for w in ("w_Q", "w_K", "w_V", "fc"):
            shp[f"{p}multi_head_attention.{w}.weight"] = (E, E)
        shp[f"{p}multi_head_attention.layer_norm.weight"] = (E,)
        shp[f"{p}multi_head_attention.layer_norm.bias"] = (E,)
        shp[f"{p}feed_forward.w_1.weight"] = (H, E); shp[f"{p}feed_forward.w_1.bias"] = (H,)
        shp[f"{p}feed_forward.w_2.weight"] = (E, H); shp[f"{p}feed_forward.w_2.bias"] = (E,)
        shp[f"{p}feed_forward.layer_norm.weight"] = (E,); shp[f"{p}feed_forward.layer_norm.bias"] = (E,)
    a_text, a_img, a_mm = _adapter_counts(cfg)
    d_mm = min(cfg.d_text, cfg.d_img) if cfg.asym else cfg.d_text
    r_mm = cfg.r_bert
    if cfg.asym and cfg.d_text > cfg.d_img:
        r_mm = cfg.r_cv
    m = "mm_encoder."
    if cfg.asym:
        shp[m + "cv_pre_fc.weight"] = (E, E); shp[m + "cv_pre_fc.bias"] = (E,)
        shp[m + "bert_pre_fc.weight"] = (E, E); shp[m + "bert_pre_fc.bias"] = (E,)
    else:
        shp[m + "cv_pre_fc.weight"] = (E, cfg.d_img); shp[m + "cv_pre_fc.bias"] = (E,)
        shp[m + "bert_pre_fc.weight"] = (E, cfg.d_text); shp[m + "bert_pre_fc.bias"] = (E,)

    def adapters(name, n, d, r):
        for i in range(n):
            shp[f"{m}{name}.{i}.fc_down.weight"] = (r, d); shp[f"{m}{name}.{i}.fc_down.bias"] = (r,)
            shp[f"{m}{name}.{i}.fc_up.weight"] = (d, r); shp[f"{m}{name}.{i}.fc_up.bias"] = (d,)
    adapters("cv_adapter_list", a_img, cfg.d_img, cfg.r_cv)
    adapters("bert_adapter_list", a_text, cfg.d_text, cfg.r_bert)
    if cfg.asym and cfg.d_text != cfg.d_img:
        n_dp = a_img if cfg.d_text > cfg.d_img else a_text
        for i in range(n_dp):
            shp[f"{m}down_project_list.{i}.weight"] = (d_mm, max(cfg.d_text, cfg.d_img))
            shp[f"{m}down_project_list.{i}.bias"] = (d_mm,)
    adapters("mm_adapter_list", a_mm, d_mm, r_mm)
    if cfg.asym:
        shp[m + "fc_bert.weight"] = (E, cfg.d_text); shp[m + "fc_bert.bias"] = (E,)
        shp[m + "fc_cv.weight"] = (E, cfg.d_img); shp[m + "fc_cv.bias"] = (E,)
    else:
        shp[m + "fc_bert.weight"] = (cfg.d_text, cfg.d_text); shp[m + "fc_bert.bias"] = (cfg.d_text,)
        shp[m + "fc_cv.weight"] = (cfg.d_img, cfg.d_img); shp[m + "fc_cv.bias"] = (cfg.d_img,)
    shp[m + "fc_mm.weight"] = (d_mm, d_mm); shp[m + "fc_mm.bias"] = (d_mm,)
    shp[m + "fc_mm_down.weight"] = (E, d_mm); shp[m + "fc_mm_down.bias"] = (E,)
    n_gate_mm = min(a_img, a_text) if cfg.asym else a_img
    n_gate_text = a_text if cfg.asym else a_img
    for i in range(n_gate_text):
        shp[f"{m}side_gate_params_text.{i}"] = (1,)
    for i in range(a_img):
        shp[f"{m}side_gate_params_cv.{i}"] = (1,)
    for i in range(n_gate_mm):
        shp[f"{m}side_gate_params_mm.{i}"] = (1,)
    shp["com_dense.weight"] = (E, 3 * E); shp["com_dense.bias"] = (E,)
    return shp


def make_params(cfg: PathConfig, seed: int, perturb: bool = True):
    """Seeded fp32 parameters following the reference's init distributions.

    adapters N(0, 0.01^2) / zero bias (modules.py:102-110); user encoder xavier-normal / zero bias,
    LayerNorm 1/0 (encoders.py:45-51); other Linear layers U(+-1/sqrt(fan_in)) (torch default);
    gates 0 (model.py:284-296).  ``perturb`` adds N(0, 0.05^2) to biases, gates and LayerNorm
    affine terms so that zero inputs, gates and biases are all exercised.
    """
    rng = np.random.default_rng(seed + 15485863)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if "adapter_list" in name:
            v = rng.standard_normal(shape) * 0.01 if name.endswith("weight") else np.zeros(shape)
        elif name.startswith("user_encoder"):
            if "layer_norm" in name:
                v = np.ones(shape) if name.endswith("weight") else np.zeros(shape)
            elif name.endswith("bias"):
                v = np.zeros(shape)
            else:
                std = math.sqrt(2.0 / (shape[0] + shape[1]))
                v = rng.standard_normal(shape) * std
        elif "side_gate" in name:
            v = np.zeros(shape)
        else:
            fan_in = shape[1] if len(shape) == 2 else None
            if fan_in is None:
                # bias: bound from the matching weight's fan_in
                wshape = out[name[:-4] + "weight"].shape
                fan_in = wshape[1]
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, size=shape)
        if perturb and (name.endswith("bias") or "side_gate" in name or "layer_norm" in name):
            v = v + rng.standard_normal(shape) * 0.05
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# cache files of the data loader (Code_Cached/preprocess_vectors.py:27-31 writes them, data_utils/dataset.py:29-34 reads them)
# ----------------------------------------------------------------------------------------------

def make_cache_case(seed: int = 4242, item_num: int = 12, layers: int = 13, d: int = 16):
    """Seeded stand-in for one dataset: ``item_id_to_keys`` (bytes ASINs, as the lmdb key table holds them), per-item
    [layers, d] fp32 states for both modalities and train sequences (``u2seq``: user -> item ids, lengths 2..11, one user
    with a repeated item)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    keys = {i: f"B{seed % 1000:03d}{i:06d}".encode() for i in range(1, item_num + 1)}
    bert = {i: torch.randn(layers, d, generator=g) for i in range(1, item_num + 1)}
    vit = {i: torch.randn(layers, d, generator=g) for i in range(1, item_num + 1)}
    rng = np.random.default_rng(seed)
    u2seq = {}
    for u, n in enumerate([2, 3, 5, 11, 7, 11, 4, 9]):
        seq = [int(v) for v in rng.integers(1, item_num + 1, size=n)]
        if u == 4:
            seq[2] = seq[0]
        u2seq[u] = seq
    return keys, bert, vit, u2seq
