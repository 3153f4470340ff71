"""Arithmetic-mode selection for the hot path.

fp32  : fp32 FMA everywhere -- matches the reference run without autocast to <= 1e-5 relative.
bf16  : bf16 tensor-core GEMMs with fp32 accumulation -- the analogue of the reference's
        torch.cuda.amp.autocast() region (Code_Cached/run.py:380), <= 1e-2 relative.

Default: follow the caller like the reference does -- inside an autocast region use the fast mode,
otherwise fp32.  ``set_compute_mode('fp32'|'bf16'|None)`` or IISAN_B200_COMPUTE overrides it.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

_forced = os.environ.get("IISAN_B200_COMPUTE") or None


def set_compute_mode(mode):
    global _forced
    if mode not in (None, "fp32", "bf16"):
        raise ValueError(mode)
    _forced = mode


def compute_mode() -> int:
    mode = _forced
    if mode is None:
        mode = "bf16" if torch.is_autocast_enabled() else "fp32"
    return L.COMPUTE_BF16 if mode == "bf16" else L.COMPUTE_FP32
