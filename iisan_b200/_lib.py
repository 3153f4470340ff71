"""ctypes binding of libiisan_b200.so (the C ABI declared in include/iisan_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails
this module raises.  Build it with ``python -m iisan_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

MAX_STAGES = 64
MAX_BLOCKS = 8
ABI_VERSION = 3

F32, BF16, F16 = 0, 1, 2
K_STREAM, K_GEMM, K_USER, K_CE, K_MISC, K_CHAIN, K_CHAIN_BWD = range(7)
KERNEL_CLASSES = ("stream", "gemm", "user", "ce", "misc", "chain", "chain_bwd", "wgrad_stream")
COMPUTE_FP32, COMPUTE_BF16 = 0, 1

# IISAN_B200_LIB: load an instrumented build variant instead (iisan_b200/build.py; debugging aid, same ABI)
_LIB_PATH = os.environ.get("IISAN_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libiisan_b200.so")

vp = C.c_void_p


class AdapterPtrs(C.Structure):
    _fields_ = [("w_down", vp), ("b_down", vp), ("w_up", vp), ("b_up", vp)]


class LinearPtrs(C.Structure):
    _fields_ = [("w", vp), ("b", vp)]


class SanParams(C.Structure):
    _fields_ = [("text", AdapterPtrs * MAX_STAGES), ("img", AdapterPtrs * MAX_STAGES), ("mm", AdapterPtrs * MAX_STAGES),
                ("down_project", LinearPtrs * MAX_STAGES),
                ("gate_text", vp * MAX_STAGES), ("gate_img", vp * MAX_STAGES), ("gate_mm", vp * MAX_STAGES),
                ("fc_text", LinearPtrs), ("fc_img", LinearPtrs), ("fc_mm", LinearPtrs),
                ("pre_text", LinearPtrs), ("pre_img", LinearPtrs), ("mm_down", LinearPtrs)]


i32 = C.c_int32


class SanDesc(C.Structure):
    _fields_ = [("n_items", i32), ("d_text", i32), ("d_img", i32), ("d_mm", i32),
                ("layers_text", i32), ("layers_img", i32),
                ("r_text", i32), ("r_img", i32), ("r_mm", i32), ("emb", i32), ("n_stages", i32),
                ("text_adapter", i32 * MAX_STAGES), ("text_layer", i32 * MAX_STAGES),
                ("img_adapter", i32 * MAX_STAGES), ("img_layer", i32 * MAX_STAGES),
                ("mm_index", i32 * MAX_STAGES),
                ("asym", i32), ("remove_first", i32), ("state_dtype", i32), ("compute", i32), ("out_ld", i32),
                ("activation", i32)]


class UeBlockPtrs(C.Structure):
    _fields_ = [(n, vp) for n in ("w_q", "w_k", "w_v", "w_fc", "ln1_w", "ln1_b", "w1", "b1", "w2", "b2", "ln2_w", "ln2_b")]


class UeParams(C.Structure):
    _fields_ = [("pos_emb", vp), ("ln_w", vp), ("ln_b", vp), ("blocks", UeBlockPtrs * MAX_BLOCKS)]


class UeDesc(C.Structure):
    _fields_ = [("users", i32), ("seq_len", i32), ("emb", i32), ("heads", i32), ("n_blocks", i32), ("training", i32),
                ("dropout_p", C.c_float), ("seed", C.c_uint64), ("offset", C.c_uint64), ("compute", i32), ("reserved", i32),
                ("offset_dev", vp)]


class AdamTensor(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("numel", C.c_int64), ("lr", C.c_float),
                ("reserved", i32)]


class CeDesc(C.Structure):
    _fields_ = [("row_users", i32), ("col_users", i32), ("seq_len", i32), ("emb", i32), ("user_offset", C.c_int64),
                ("compute", i32), ("reserved", i32)]


_SIGNATURES = {
    "iisan_abi_version": (C.c_int, []),
    "iisan_status_string": (C.c_char_p, [C.c_int]),
    "iisan_last_cuda_error": (C.c_int, []),
    "iisan_last_cuda_error_string": (C.c_char_p, []),
    "iisan_sizeof": (C.c_size_t, [C.c_int]),
    "iisan_launch_count": (C.c_int64, [C.c_int]),
    "iisan_timing_enable": (C.c_int, [C.c_int]),
    "iisan_timing_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "iisan_san_workspace_bytes": (C.c_size_t, [C.POINTER(SanDesc)]),
    "iisan_san_forward": (C.c_int, [C.POINTER(SanDesc), C.POINTER(SanParams), vp, vp, vp, C.c_size_t, vp, vp]),
    "iisan_san_backward": (C.c_int, [C.POINTER(SanDesc), C.POINTER(SanParams), C.POINTER(SanParams), vp, vp, vp, C.c_size_t, vp, vp]),
    "iisan_linear_forward": (C.c_int, [i32, i32, i32, vp, C.c_int64, vp, vp, vp, C.c_int64, i32, vp]),
    "iisan_linear_backward": (C.c_int, [i32, i32, i32, vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int64, vp, vp, i32, vp]),
    "iisan_gemm_bf16": (C.c_int, [i32, i32, i32, vp, C.c_int64, i32, vp, C.c_int64, i32, vp, C.c_int64, vp, C.c_int64, vp, i32, i32, vp]),
    "iisan_user_encoder_workspace_bytes": (C.c_size_t, [C.POINTER(UeDesc)]),
    "iisan_user_encoder_forward": (C.c_int, [C.POINTER(UeDesc), C.POINTER(UeParams), vp, C.c_int64, vp, vp, C.c_size_t, vp, vp]),
    "iisan_user_encoder_backward": (C.c_int, [C.POINTER(UeDesc), C.POINTER(UeParams), C.POINTER(UeParams), vp, C.c_int64, vp, vp,
                                              C.c_size_t, vp, vp, vp]),
    "iisan_inbatch_ce_workspace_bytes": (C.c_size_t, [C.POINTER(CeDesc)]),
    "iisan_inbatch_ce_forward": (C.c_int, [C.POINTER(CeDesc), vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp, vp, vp]),
    "iisan_inbatch_ce_backward": (C.c_int, [C.POINTER(CeDesc), vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp, vp, vp, vp, vp]),
    "iisan_inbatch_ce_masks": (C.c_int, [C.POINTER(CeDesc), vp, vp, vp, vp, vp, vp]),
    "iisan_inbatch_ce_masks_fast": (C.c_int, [C.POINTER(CeDesc), vp, vp, vp, vp, vp, C.c_size_t, vp, vp]),
    "iisan_adam_step": (C.c_int, [C.POINTER(AdamTensor), i32, C.c_float, C.c_float, C.c_float, vp, i32, vp]),
    "iisan_stage_states_h2d": (C.c_int, [vp, vp, C.c_int64, i32, i32, i32, C.POINTER(i32), i32, vp]),
    "iisan_gather_states": (C.c_int, [vp, i32, C.c_int64, i32, i32, vp, i32, vp, i32, vp, vp]),
    "iisan_pack_states": (C.c_int, [vp, i32, C.c_int64, i32, i32, vp, i32, vp, vp]),
    "iisan_san_fused_eligible": (C.c_int, [C.POINTER(SanDesc)]),
    "iisan_eval_ranks": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "iisan_probe_tile_stream": (C.c_int, [vp, C.c_int64, i32, i32, C.POINTER(i32), i32, i32, i32, i32, vp, vp]),
    "iisan_debug_chain_generation": (C.c_int, [i32]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class IisanLibraryError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once).  Raises IisanLibraryError when it is missing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise IisanLibraryError(
            f"{_LIB_PATH} not found: the CUDA extension is not built. Run `python -m iisan_b200.build` "
            "(there is no CPU / PyTorch fallback for the IISAN hot path).")
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing -> loud
        fn.restype = res
        fn.argtypes = args
    v = lib.iisan_abi_version()
    if v != ABI_VERSION:
        raise IisanLibraryError(f"ABI version mismatch: library {v}, binding {ABI_VERSION}")
    for which, st in enumerate((SanDesc, SanParams, UeDesc, UeParams, CeDesc, AdamTensor)):
        if lib.iisan_sizeof(which) != C.sizeof(st):
            raise IisanLibraryError(f"struct layout mismatch for {st.__name__}: C {lib.iisan_sizeof(which)} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status == 0:
        return
    lib = load()
    msg = lib.iisan_status_string(status).decode()
    if status == 2:
        msg += ": " + lib.iisan_last_cuda_error_string().decode()
    raise IisanLibraryError(f"{what} failed: {msg}")


def torch_dtype_code(dt) -> int:
    import torch
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return F16
    raise IisanLibraryError(f"unsupported hidden-state dtype {dt}")


def require_cuda(t, name: str):
    if not t.is_cuda:
        raise IisanLibraryError(
            f"{name} is on {t.device}: the iisan_b200 hot path runs on a CUDA device only (no CPU fallback)")
