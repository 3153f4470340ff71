"""Host-side logic of the hot path: layer selection, the stage plan of the side-adapter network and the
binders that turn nn.Parameters into the pointer tables of the C ABI.

The stage plan restates the control flow of the reference forward:
  Code_Cached/model/model.py:263-268, 318-338       (every tower runs every stage; quirk Q1: both layer
                                                    lists come from side_adapter_vit_list)
  Code_Cached_Asym/model/model.py:265-270, 353-417  (group layer-drop: the tower with more adapters runs
                                                    `diff` solo stages first, then the paired stages)
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass

import torch

from . import _lib as L


def _parse_list(s: str):
    return [int(tok) + 1 for tok in str(s).split(",")]


@dataclass
class SanPlan:
    asym: bool
    remove_first: bool
    d_text: int
    d_img: int
    r_text: int
    r_img: int
    r_mm: int
    emb: int
    layers_text_sel: list
    layers_img_sel: list
    n_text: int
    n_img: int
    n_mm: int
    n_gate_text: int
    n_gate_img: int
    n_gate_mm: int
    n_down_project: int
    stages: list          # (text_adapter, text_layer, img_adapter, img_layer, mm_index) with -1 = idle
    activation: int = 0   # iisan_activation: 0 ReLU, 1 exact GELU (args.adapter_activation, CC/model/modules.py:104-107)

    @property
    def d_mm(self):
        return min(self.d_text, self.d_img)

    # Layers the kernels READ: the selected ones plus, with remove_first, cached layer 0 -- the towers then start from the
    # embedding output instead of zero (CC/model/model.py:305-308) although layer 0 feeds no adapter stage.  Stores and
    # partial host-to-device copies must carry these, not just the selected layers.
    @property
    def layers_text_read(self):
        return sorted(set(self.layers_text_sel) | ({0} if self.remove_first else set()))

    @property
    def layers_img_read(self):
        return sorted(set(self.layers_img_sel) | ({0} if self.remove_first else set()))


def make_plan(args, asym: bool) -> SanPlan:
    """Validate the configuration and lay out the stages.  Unsupported reference options are rejected at
    construction time (SURVEY.md 8b 'Errors')."""
    if getattr(args, "fusion_method", "gated") != "gated":
        raise NotImplementedError("iisan_b200 builds the gated fusion only (fusion_method='gated')")
    if "intra" not in args.modality or "inter" not in args.modality:
        raise NotImplementedError("iisan_b200 builds modality='intra_inter' only (the only configuration the "
                                  "reference Code_Cached forward can run, SURVEY.md Q2)")
    # CC/model/modules.py:104-107: nn.GELU() iff the string is exactly "GELU", nn.ReLU() for anything else
    activation = 1 if getattr(args, "adapter_activation", "RELU") == "GELU" else 0
    rf = (args.remove_first == "TRUE")
    E = int(args.embedding_dim)
    if asym:
        d_text, d_img = int(args.text_embedding_dim), int(args.image_embedding_dim)
        t_sel = _parse_list(args.side_adapter_bert_list)
        i_sel = _parse_list(args.side_adapter_vit_list)
    else:
        d_text = d_img = 768                       # hard-coded in the reference (model.py:260,301-302)
        if int(args.word_embedding_dim) != 768:
            raise NotImplementedError("Code_Cached semantics fix the hidden width to 768; use iisan_b200.model_asym")
        t_sel = _parse_list(args.side_adapter_vit_list)
        i_sel = _parse_list(args.side_adapter_bert_list if rf else args.side_adapter_vit_list)
    if not rf:
        t_sel, i_sel = [0] + t_sel, [0] + i_sel
    n_text, n_img = len(t_sel), len(i_sel)
    r_img, r_text = int(args.cv_adapter_down_size), int(args.bert_adapter_down_size)
    if asym:
        if d_text > d_img:
            n_mm, r_mm, n_dp = n_img, r_img, n_img
        elif d_text < d_img:
            n_mm, r_mm, n_dp = n_text, r_text, n_text
        else:
            n_mm, r_mm, n_dp = n_text, r_text, 0
        n_gate_text, n_gate_img, n_gate_mm = n_text, n_img, min(n_text, n_img)
    else:
        if n_text != n_img:
            raise NotImplementedError("Code_Cached semantics need equal adapter counts")
        n_mm, r_mm, n_dp = n_img, r_text, 0
        n_gate_text = n_gate_img = n_gate_mm = n_img
    stages = []
    if asym:
        solo_t, solo_i = max(0, n_text - n_img), max(0, n_img - n_text)
        for s in range(solo_t):
            stages.append((s, t_sel[s], -1, -1, -1))
        for s in range(solo_i):
            stages.append((-1, -1, s, i_sel[s], -1))
        for s in range(min(n_text, n_img)):
            stages.append((s + solo_t, t_sel[s + solo_t], s + solo_i, i_sel[s + solo_i], s))
    else:
        for s in range(n_img):
            stages.append((s, t_sel[s], s, i_sel[s], s))
    if len(stages) > L.MAX_STAGES:
        raise NotImplementedError(f"more than {L.MAX_STAGES} stages")
    return SanPlan(asym, rf, d_text, d_img, r_text, r_img, r_mm, E, t_sel, i_sel, n_text, n_img, n_mm,
                   n_gate_text, n_gate_img, n_gate_mm, n_dp, stages, activation)


# --------------------------------------------------------------------------------------------------
# binders
# --------------------------------------------------------------------------------------------------
_ADAPTER_RE = re.compile(r"^(cv|bert|mm)_adapter_list\.(\d+)\.(fc_down|fc_up)\.(weight|bias)$")
_DP_RE = re.compile(r"^down_project_list\.(\d+)\.(weight|bias)$")
_GATE_RE = re.compile(r"^side_gate_params_(text|cv|mm)\.(\d+)$")
_HEADS = {"fc_bert": "fc_text", "fc_cv": "fc_img", "fc_mm": "fc_mm", "bert_pre_fc": "pre_text", "cv_pre_fc": "pre_img",
          "fc_mm_down": "mm_down"}


def _san_slot(table: L.SanParams, name: str, ptr: int):
    m = _ADAPTER_RE.match(name)
    if m:
        tower = {"cv": table.img, "bert": table.text, "mm": table.mm}[m.group(1)]
        field = {"fc_down": "down", "fc_up": "up"}[m.group(3)]
        setattr(tower[int(m.group(2))], ("w_" if m.group(4) == "weight" else "b_") + field, ptr)
        return
    m = _DP_RE.match(name)
    if m:
        setattr(table.down_project[int(m.group(1))], "w" if m.group(2) == "weight" else "b", ptr)
        return
    m = _GATE_RE.match(name)
    if m:
        {"text": table.gate_text, "cv": table.gate_img, "mm": table.gate_mm}[m.group(1)][int(m.group(2))] = ptr
        return
    head, _, kind = name.rpartition(".")
    if head in _HEADS:
        setattr(getattr(table, _HEADS[head]), "w" if kind == "weight" else "b", ptr)
        return
    raise KeyError(f"unexpected SAN parameter {name}")


class _BinderBase:
    """Caches the pointer table of a parameter tuple (parameters are updated in place by the optimizer, so
    the table only changes when the module is moved / re-materialised)."""

    def __init__(self):
        self._key = None
        self._table = None

    def _build(self, params):
        raise NotImplementedError

    def param_table(self, params):
        key = tuple(p.data_ptr() for p in params)
        if key != self._key:
            self._table = self._build(params)
            self._key = key
        return self._table


class SanBinder(_BinderBase):
    def __init__(self, plan: SanPlan, names):
        super().__init__()
        self.plan = plan
        self.names = list(names)
        used = set()
        for (ta, _, ia, _, mi) in plan.stages:
            if ta >= 0:
                used.add(f"bert_adapter_list.{ta}."); used.add(f"side_gate_params_text.{ta}")
            if ia >= 0:
                used.add(f"cv_adapter_list.{ia}."); used.add(f"side_gate_params_cv.{ia}")
            if mi >= 0:
                used.add(f"mm_adapter_list.{mi}."); used.add(f"side_gate_params_mm.{mi}"); used.add(f"down_project_list.{mi}.")
        self.used = [any(n == u or (u.endswith(".") and n.startswith(u)) for u in used) or n.split(".")[0] in _HEADS
                     for n in self.names]
        self._desc_cache = {}
        self._sel_dev = {}

    def read_layers_dev(self, device):
        """(image, text) layer lists the towers read, as int32 tensors on ``device`` (cached: created outside any graph capture)."""
        key = str(device)
        if key not in self._sel_dev:
            self._sel_dev[key] = (torch.as_tensor(self.plan.layers_img_read, dtype=torch.int32, device=device),
                                  torch.as_tensor(self.plan.layers_text_read, dtype=torch.int32, device=device))
        return self._sel_dev[key]

    def _build(self, params):
        t = L.SanParams()
        for n, p in zip(self.names, params):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise L.IisanLibraryError(f"parameter {n} must be contiguous fp32")
            _san_slot(t, n, p.data_ptr())
        return t

    def grad_table(self, params, device):
        sizes = [p.numel() if u else 0 for p, u in zip(params, self.used)]
        # 64-element alignment keeps every slice 256 B aligned
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot); tot += (s + 63) // 64 * 64
        from .ops import grad_zeros
        flat = grad_zeros(max(tot, 1), device)
        t = L.SanParams()
        views = []
        base = flat.data_ptr()
        for n, p, u, o, s in zip(self.names, params, self.used, offs, sizes):
            if not u:
                views.append(None)
                continue
            views.append(flat[o:o + s].view(p.shape))
            _san_slot(t, n, base + 4 * o)
        return flat, views, t

    def desc(self, image, text, compute, packed=False):
        """``packed``: the tensors hold ONLY the selected layers, in increasing layer order (iisan_b200.store): layer index l of
        the plan is addressed at its rank among the selected layers."""
        pl = self.plan
        if image.dim() not in (3, 4) or text.dim() not in (3, 4):
            raise L.IisanLibraryError("hidden states must be [B,11,layers,d] or [b,layers,d]")
        if image.dtype != text.dtype:
            raise L.IisanLibraryError("image and text hidden states must share a dtype")
        li, di = image.shape[-2], image.shape[-1]
        lt, dt = text.shape[-2], text.shape[-1]
        n = image.numel() // (li * di)
        if text.numel() // (lt * dt) != n:
            raise L.IisanLibraryError("image and text batches disagree on the number of items")
        if di != pl.d_img or dt != pl.d_text:
            raise L.IisanLibraryError(f"hidden widths ({dt},{di}) do not match the module ({pl.d_text},{pl.d_img})")
        rank_i = {l: k for k, l in enumerate(pl.layers_img_read)}        # with remove_first layer 0 is packed too (rank 0)
        rank_t = {l: k for k, l in enumerate(pl.layers_text_read)}
        if packed:
            if li != len(rank_i) or lt != len(rank_t):
                raise L.IisanLibraryError(f"packed states must hold exactly the layers the towers read ({len(rank_i)} image, {len(rank_t)} text)")
        elif max(pl.layers_img_sel) >= li or max(pl.layers_text_sel) >= lt:
            raise L.IisanLibraryError("a selected layer index is outside the cached states")
        key = (n, li, lt, image.dtype, compute, bool(packed))
        d = self._desc_cache.get(key)
        if d is None:
            d = L.SanDesc()
            d.n_items, d.d_text, d.d_img, d.d_mm = n, pl.d_text, pl.d_img, pl.d_mm
            d.layers_text, d.layers_img = lt, li
            d.r_text, d.r_img, d.r_mm, d.emb = pl.r_text, pl.r_img, pl.r_mm, pl.emb
            d.n_stages = len(pl.stages)
            for s, (ta, tl, ia, il, mi) in enumerate(pl.stages):
                d.text_adapter[s], d.text_layer[s] = ta, (rank_t[tl] if (packed and ta >= 0) else tl)
                d.img_adapter[s], d.img_layer[s] = ia, (rank_i[il] if (packed and ia >= 0) else il)
                d.mm_index[s] = mi
            d.asym, d.remove_first = int(pl.asym), int(pl.remove_first)
            d.activation = int(pl.activation)
            d.state_dtype = L.torch_dtype_code(image.dtype)
            d.compute = compute
            d.out_ld = 3 * pl.emb
            self._desc_cache[key] = d
        return d


_UE_BLOCK_RE = re.compile(r"^transformer_encoder\.transformer_blocks\.(\d+)\.(.+)$")
_UE_BLOCK_FIELDS = {
    "multi_head_attention.w_Q.weight": "w_q", "multi_head_attention.w_K.weight": "w_k",
    "multi_head_attention.w_V.weight": "w_v", "multi_head_attention.fc.weight": "w_fc",
    "multi_head_attention.layer_norm.weight": "ln1_w", "multi_head_attention.layer_norm.bias": "ln1_b",
    "feed_forward.w_1.weight": "w1", "feed_forward.w_1.bias": "b1",
    "feed_forward.w_2.weight": "w2", "feed_forward.w_2.bias": "b2",
    "feed_forward.layer_norm.weight": "ln2_w", "feed_forward.layer_norm.bias": "ln2_b",
}
_UE_TOP = {"transformer_encoder.position_embedding.weight": "pos_emb", "transformer_encoder.layer_norm.weight": "ln_w",
           "transformer_encoder.layer_norm.bias": "ln_b"}


def _ue_slot(table: L.UeParams, name: str, ptr: int):
    if name in _UE_TOP:
        setattr(table, _UE_TOP[name], ptr)
        return
    m = _UE_BLOCK_RE.match(name)
    if m and m.group(2) in _UE_BLOCK_FIELDS:
        setattr(table.blocks[int(m.group(1))], _UE_BLOCK_FIELDS[m.group(2)], ptr)
        return
    raise KeyError(f"unexpected user-encoder parameter {name}")


class UserEncoderBinder(_BinderBase):
    def __init__(self, names, emb, heads, n_blocks, dropout):
        super().__init__()
        self.names = list(names)
        self.emb, self.heads, self.n_blocks, self.dropout = emb, heads, n_blocks, float(dropout)
        if n_blocks > L.MAX_BLOCKS:
            raise NotImplementedError(f"more than {L.MAX_BLOCKS} transformer blocks")

    def _build(self, params):
        t = L.UeParams()
        for n, p in zip(self.names, params):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise L.IisanLibraryError(f"parameter {n} must be contiguous fp32")
            _ue_slot(t, n, p.data_ptr())
        return t

    def grad_table(self, params, device):
        offs, tot = [], 0
        for p in params:
            offs.append(tot); tot += (p.numel() + 63) // 64 * 64
        from .ops import grad_zeros
        flat = grad_zeros(tot, device)
        t = L.UeParams()
        views = []
        base = flat.data_ptr()
        for n, p, o in zip(self.names, params, offs):
            views.append(flat[o:o + p.numel()].view(p.shape))
            _ue_slot(t, n, base + 4 * o)
        return flat, views, t

    def desc(self, users, seq_len, training, seed, offset, compute, offset_dev=None):
        d = L.UeDesc()
        d.users, d.seq_len, d.emb, d.heads, d.n_blocks = users, seq_len, self.emb, self.heads, self.n_blocks
        d.training = int(bool(training) and self.dropout > 0)
        d.dropout_p = self.dropout
        d.seed, d.offset = int(seed), int(offset)
        d.offset_dev = offset_dev
        d.compute = compute
        return d
