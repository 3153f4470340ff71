"""Parameter containers of the reference `model.modules` (Code_Cached/model/modules.py), same class names,
constructor arguments, parameter names and initialisation -- the compute lives in libiisan_b200.so.

The SASRec sub-blocks (PositionwiseFeedForward, MultiHeadedAttention, TransformerBlock) only own
parameters: their arithmetic is executed by the fused user-encoder entry points, driven from
TransformerEncoder.forward.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib as L
from ..ops import LinearFn, UserEncoderFn
from ..plan import UserEncoderBinder
from ..precision import compute_mode


class PositionwiseFeedForward(nn.Module):
    """Code_Cached/model/modules.py:6-18 (parameters only)."""

    def __init__(self, d_model, d_inner, dropout):
        super().__init__()
        self.w_1 = nn.Linear(d_model, d_inner)
        self.w_2 = nn.Linear(d_inner, d_model)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.dropout = nn.Dropout(dropout)
        self.activate = nn.ReLU()


class MultiHeadedAttention(nn.Module):
    """Code_Cached/model/modules.py:35-64 (parameters only; Q/K/V/fc are bias-free)."""

    def __init__(self, n_heads, d_model, dropout):
        super().__init__()
        if d_model % n_heads:
            raise ValueError("d_model must be divisible by n_heads")
        self.d_model, self.n_heads = d_model, n_heads
        self.d_k = self.d_v = d_model // n_heads
        for name in ("w_Q", "w_K", "w_V", "fc"):
            setattr(self, name, nn.Linear(d_model, d_model, bias=False))
        self.dropout = nn.Dropout(p=dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class TransformerBlock(nn.Module):
    """Code_Cached/model/modules.py:67-76 (parameters only)."""

    def __init__(self, d_model, n_heads, d_inner, dropout):
        super().__init__()
        self.multi_head_attention = MultiHeadedAttention(n_heads=n_heads, d_model=d_model, dropout=dropout)
        self.feed_forward = PositionwiseFeedForward(d_model=d_model, d_inner=d_inner, dropout=dropout)


class TransformerEncoder(nn.Module):
    """Code_Cached/model/modules.py:79-96.  forward() runs the whole encoder in the CUDA library."""

    def __init__(self, n_vocab, n_position, d_model, n_heads, dropout, n_layers):
        super().__init__()
        self.position_embedding = nn.Embedding(n_position, d_model)
        self.dropout = nn.Dropout(p=dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.transformer_blocks = nn.ModuleList(
            TransformerBlock(d_model=d_model, n_heads=n_heads, d_inner=d_model * 4, dropout=dropout) for _ in range(n_layers))
        self._cfg = (d_model, n_heads, n_layers, float(dropout))
        self._binder = None
        self._step_dev = None
        self.dropout_seed = None         # derived on first use from torch's seed and the rank (see _dropout_stream)

    def _bind(self):
        if self._binder is None:
            names = ["transformer_encoder." + n for n, _ in self.named_parameters()]
            d_model, n_heads, n_layers, p = self._cfg
            self._binder = UserEncoderBinder(names, d_model, n_heads, n_layers, p)
        return self._binder

    # ---- dropout stream: Philox keyed by (seed, step counter) ----
    def _dropout_stream(self, device):
        """Seed = torch's global seed (torch.manual_seed, i.e. the reference's setup_seed, run.py:465-472) mixed with the rank, so
        that seeding the run seeds the masks and the ranks of a data-parallel job draw different masks.  The step counter lives on
        the device (advanced by a stream-ordered add: valid eagerly and under CUDA-graph replay)."""
        if self.dropout_seed is None:
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()
            self.dropout_seed = (0x5EED1154 ^ int(torch.initial_seed()) ^ (rank * 0x9E3779B97F4A7C15)) & ((1 << 62) - 1)
        if self._step_dev is None or self._step_dev.device != device:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=device)
        return self._step_dev

    def dropout_state(self):
        """(seed, step) of the dropout stream, for checkpoints (kept OUT of state_dict(): its keys are the reference's)."""
        return {"seed": self.dropout_seed, "step": None if self._step_dev is None else int(self._step_dev.item())}

    def load_dropout_state(self, st, device=None):
        self.dropout_seed = st.get("seed")
        if st.get("step") is not None:
            dev = device if device is not None else next(self.parameters()).device
            self._step_dev = torch.full((1,), int(st["step"]), dtype=torch.int64, device=dev)

    def forward(self, input_embs, log_mask, att_mask=None, seq_len=None):
        # att_mask is implied by log_mask (causal + key padding, encoders.py:54-57) and rebuilt in-kernel.
        # seq_len: input_embs is a whole [B, S, E] block of which only the first seq_len slots are the input (ops.UserEncoderFn).
        params = tuple(self.parameters())
        d_model, n_heads, n_layers, p = self._cfg
        offset, seed = 0, 0
        if self.training and p > 0:
            # Each forward gets its OWN copy of the advanced counter; the backward regenerates the masks from that copy, so a
            # second training forward before the first backward (gradient accumulation, two model calls per step) cannot change
            # the masks of the first (the clone is a stream-ordered device copy: capturable).
            ctr = self._dropout_stream(input_embs.device)
            ctr += 1
            offset = ctr.clone()
            seed = self.dropout_seed
        return UserEncoderFn.apply(self._bind(), input_embs, log_mask, self.training, seed, offset,
                                   compute_mode(), seq_len, *params)


class AdapterBlock(nn.Module):
    """Code_Cached/model/modules.py:98-116: fc_up(act(fc_down(x))) + x.  The Dropout is constructed but,
    as in the reference forward, never applied.  Inside IISANAdaptedMModel the whole stage chain is
    fused; this standalone forward exists for API parity and runs on the same dense-layer kernels."""

    def __init__(self, args, input_size, down_size, dropout=0.1):
        super().__init__()
        self.fc_down = nn.Linear(input_size, down_size)
        self.fc_up = nn.Linear(down_size, input_size)
        for lin in (self.fc_down, self.fc_up):
            nn.init.normal_(lin.weight, std=1e-2)
            nn.init.zeros_(lin.bias)
        self.activate = nn.GELU() if getattr(args, "adapter_activation", "RELU") == "GELU" else nn.ReLU()     # modules.py:104-107
        self.dropout = nn.Dropout(dropout)

    def forward(self, input_embs):
        c = compute_mode()
        z = self.activate(LinearFn.apply(input_embs, self.fc_down.weight, self.fc_down.bias, c))
        return LinearFn.apply(z, self.fc_up.weight, self.fc_up.bias, c) + input_embs


class FusedLinear(nn.Linear):
    """nn.Linear whose forward/backward run through iisan_linear_* (used for com_dense)."""

    def forward(self, x):
        return LinearFn.apply(x, self.weight, self.bias, compute_mode())
