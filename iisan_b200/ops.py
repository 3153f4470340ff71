"""torch.autograd glue over the C ABI (include/iisan_b200.h).

PyTorch is only plumbing here: it owns device memory, the current stream and the autograd graph;
every FLOP of the hot path happens inside libiisan_b200.so.  Nothing in this file computes on tensors
with torch ops, and there is no fallback: tensors that are not on a CUDA device raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class GradArena:
    """One persistent fp32 buffer that the backward functions carve their zero-initialised parameter gradients from (instead of a
    fresh torch.zeros each): one memset per step, and the data-parallel step averages ALL gradients with ONE collective over
    ``used()``.  Activated by iisan_b200.engine.TrainStep; without an active arena the functions allocate as before."""

    active = None

    def __init__(self, numel, device):
        self.buf = torch.zeros(int(numel), dtype=torch.float32, device=device)
        self.off = 0
        self.high = 0                       # high-water mark of the slices handed out so far

    def reset(self):
        # one memset per step (stream-ordered, capturable); after the first step only the part the last step used
        (self.buf if self.high == 0 else self.buf[:self.high]).zero_()
        self.off = 0

    def take(self, numel):
        n = (int(numel) + 63) // 64 * 64    # 256-byte aligned slices
        if self.off + n > self.buf.numel():
            raise L.IisanLibraryError("gradient arena too small")
        v = self.buf[self.off:self.off + int(numel)]
        self.off += n
        self.high = max(self.high, self.off)
        return v

    def used(self):
        return self.buf[:self.off]


def grad_zeros(numel, device):
    """Zero-initialised flat fp32 gradient buffer: a slice of the active arena, else a fresh tensor."""
    a = GradArena.active
    if a is not None and a.buf.device == torch.device(device):
        return a.take(numel)
    return torch.zeros(max(int(numel), 1), dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------------------
# dense layer (com_dense)
# --------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b through iisan_linear_forward / iisan_linear_backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, compute):
        L.require_cuda(x, "linear input")
        lib = L.load()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        x2 = x2.float()
        rows, k = x2.shape
        n = weight.shape[0]
        y = torch.empty(rows, n, dtype=torch.float32, device=x.device)
        L.check(lib.iisan_linear_forward(rows, n, k, _p(x2), x2.stride(0), _p(weight), _p(bias), _p(y), n, compute, _stream()),
                "iisan_linear_forward")
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.compute = compute
        ctx.lead = lead
        return y.view(*lead, n)

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x2, weight = ctx.saved_tensors
        rows, k = x2.shape
        n = weight.shape[0]
        dy2 = dy.reshape(rows, n)
        if dy2.stride(-1) != 1 or dy2.dtype != torch.float32:
            dy2 = dy2.contiguous().float()
        dx = torch.empty(rows, k, dtype=torch.float32, device=dy.device) if ctx.needs_input_grad[0] else None
        dw = grad_zeros(weight.numel(), dy.device).view(weight.shape) if ctx.needs_input_grad[1] else None
        db = grad_zeros(n, dy.device)[:n] if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        L.check(lib.iisan_linear_backward(rows, n, k, _p(x2), x2.stride(0), _p(weight), _p(dy2), dy2.stride(0), _p(dx), k,
                                          _p(dw), _p(db), ctx.compute, _stream()), "iisan_linear_backward")
        return (None if dx is None else dx.view(*ctx.lead, k)), dw, db, None


# --------------------------------------------------------------------------------------------------
# side-adapter network
# --------------------------------------------------------------------------------------------------
class SanFn(torch.autograd.Function):
    """out[N, 3E] = (cv | text | mm) embeddings.  `binder` builds descriptors and pointer tables."""

    @staticmethod
    def forward(ctx, binder, image, text, compute, packed, *params):
        L.require_cuda(image, "image hidden states")
        L.require_cuda(text, "text hidden states")
        lib = L.load()
        image = image.contiguous()
        text = text.contiguous()
        desc = binder.desc(image, text, compute, packed)
        if (compute == L.COMPUTE_BF16 and not packed and image.dtype != torch.bfloat16
                and lib.iisan_san_fused_eligible(C.byref(desc))):
            # states stored as the reference writes them (fp32 .pt files; fp16 for the LLaMA / EVA caches): select the layers the
            # towers read and round them to bf16 ONCE (iisan_pack_states), then run the fused chain kernels on the packed batch
            # instead of the layered path (which reads the stored dtype in place, three launches per stage)
            sel_i, sel_t = binder.read_layers_dev(image.device)        # device int32 tensors, made once (no H2D copy under capture)
            image = pack_states(image, sel_i)
            text = pack_states(text, sel_t)
            packed = True
            desc = binder.desc(image, text, compute, True)
        ptrs = binder.param_table(params)
        n, e = desc.n_items, desc.emb
        out = torch.empty(n, 3 * e, dtype=torch.float32, device=image.device)
        nbytes = lib.iisan_san_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise L.IisanLibraryError("iisan_san_workspace_bytes rejected the descriptor")
        ws = _workspace(nbytes, image.device)
        L.check(lib.iisan_san_forward(C.byref(desc), C.byref(ptrs), _p(image), _p(text), _p(ws), ws.numel(), _p(out), _stream()),
                "iisan_san_forward")
        ctx.binder, ctx.desc, ctx.ptrs, ctx.ws = binder, desc, ptrs, ws
        ctx.image, ctx.text = image, text
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = L.load()
        d_out = d_out.contiguous().float()
        flat, views, gptrs = ctx.binder.grad_table(ctx.params, d_out.device)
        L.check(lib.iisan_san_backward(C.byref(ctx.desc), C.byref(ctx.ptrs), C.byref(gptrs), _p(ctx.image), _p(ctx.text),
                                       _p(ctx.ws), ctx.ws.numel(), _p(d_out), _stream()), "iisan_san_backward")
        ctx.ws = None
        return (None, None, None, None, None, *views)


# --------------------------------------------------------------------------------------------------
# SASRec user encoder
# --------------------------------------------------------------------------------------------------
class UserEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, binder, embs, log_mask, training, seed, offset, compute, use_len, *params):
        """``use_len``: None, or the number of leading sequence slots of ``embs`` [B, S, E] that form the input (the model
        passes the whole [B, L+1, E] item-embedding block with use_len = L: the gradient then comes back as ONE [B, S, E]
        tensor with a zero last slot, instead of through autograd's slice backward = a fill and a strided copy per step)."""
        L.require_cuda(embs, "user-encoder input")
        lib = L.load()
        if embs.dtype != torch.float32 or embs.stride(-1) != 1 or embs.stride(1) != embs.shape[2]:
            embs = embs.float().contiguous()
        b, l, e = embs.shape
        ctx.full = use_len is not None and int(use_len) < l
        if ctx.full:
            if embs.stride(0) != l * e:
                embs = embs.contiguous()
            l = int(use_len)
        log_mask = log_mask.to(device=embs.device, dtype=torch.float32).contiguous()
        offset_dev = None
        if torch.is_tensor(offset):                       # device-side step counter (CUDA-graph safe dropout stream)
            ctx.offset_t = offset
            offset_dev, offset = offset.data_ptr(), 0
        desc = binder.desc(b, l, training, seed, offset, compute, offset_dev)
        ptrs = binder.param_table(params)
        nbytes = lib.iisan_user_encoder_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise L.IisanLibraryError("iisan_user_encoder_workspace_bytes rejected the descriptor")
        ws = _workspace(nbytes, embs.device)
        out = torch.empty(b, l, e, dtype=torch.float32, device=embs.device)
        L.check(lib.iisan_user_encoder_forward(C.byref(desc), C.byref(ptrs), _p(embs), embs.stride(0), _p(log_mask), _p(ws),
                                               ws.numel(), _p(out), _stream()), "iisan_user_encoder_forward")
        ctx.binder, ctx.desc, ctx.ptrs, ctx.ws = binder, desc, ptrs, ws
        ctx.embs, ctx.log_mask, ctx.params = embs, log_mask, params
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = L.load()
        d_out = d_out.contiguous().float()
        b, l, e = d_out.shape
        flat, views, gptrs = ctx.binder.grad_table(ctx.params, d_out.device)
        ld_user = ctx.embs.stride(0)                      # d_embs shares the input's user stride
        d_full = torch.zeros(b, ld_user // e, e, dtype=torch.float32, device=d_out.device)
        d_embs = d_full[:, :l]
        L.check(lib.iisan_user_encoder_backward(C.byref(ctx.desc), C.byref(ctx.ptrs), C.byref(gptrs), _p(ctx.embs),
                                                ctx.embs.stride(0), _p(ctx.log_mask), _p(ctx.ws), ctx.ws.numel(), _p(d_out),
                                                _p(d_embs), _stream()), "iisan_user_encoder_backward")
        ctx.ws = None
        return (None, d_full if ctx.full else d_embs, None, None, None, None, None, None, *views)


# --------------------------------------------------------------------------------------------------
# in-batch cross-entropy
# --------------------------------------------------------------------------------------------------
def make_ce_desc(row_users, col_users, seq_len, emb, user_offset, compute):
    d = L.CeDesc()
    d.row_users, d.col_users, d.seq_len, d.emb = row_users, col_users, seq_len, emb
    d.user_offset, d.compute = user_offset, compute
    return d


class InBatchCeFn(torch.autograd.Function):
    """(loss_sum, n_valid, loss) of the local rows against the column pool.

    prec [B*L, E]; score [C_users*(L+1), E]; ids int64; log_mask fp32 [.., L]; pop_prob fp32 table.
    ``loss`` = loss_sum / n_valid is the reference's per-rank mean (CC/model/model.py:104); ``loss_sum``
    and ``n_valid`` feed the global-negative normalisation.
    """

    @staticmethod
    def forward(ctx, prec, score, ids_rows, ids_cols, lm_rows, lm_cols, pop_prob, user_offset, compute):
        L.require_cuda(prec, "prec_vec")
        lib = L.load()
        prec = prec.contiguous().float()
        score = score.contiguous().float()
        ids_rows = ids_rows.contiguous().view(-1)
        ids_cols = ids_cols.contiguous().view(-1)
        lm_rows = lm_rows.contiguous().float()
        lm_cols = lm_cols.contiguous().float()
        b, l = lm_rows.shape
        bc = lm_cols.shape[0]
        e = prec.shape[1]
        desc = make_ce_desc(b, bc, l, e, user_offset, compute)
        nbytes = lib.iisan_inbatch_ce_workspace_bytes(C.byref(desc))
        if nbytes == 0:
            raise L.IisanLibraryError("iisan_inbatch_ce_workspace_bytes rejected the descriptor")
        ws = _workspace(nbytes, prec.device)
        res = torch.empty(2, dtype=torch.float32, device=prec.device)       # loss_sum, loss
        n_valid = torch.empty(1, dtype=torch.int32, device=prec.device)
        L.check(lib.iisan_inbatch_ce_forward(C.byref(desc), _p(prec), _p(score), _p(ids_rows), _p(ids_cols), _p(lm_rows),
                                             _p(lm_cols), _p(pop_prob), _p(ws), ws.numel(), _p(res[0:1]), _p(n_valid),
                                             _p(res[1:2]), _stream()), "iisan_inbatch_ce_forward")
        ctx.desc, ctx.ws = desc, ws
        ctx.tensors = (prec, score, ids_rows, ids_cols, lm_rows, lm_cols, pop_prob)
        ctx.n_valid = n_valid
        ctx.mark_non_differentiable(n_valid)
        ctx.set_materialize_grads(False)          # an unused output (loss_sum or loss) arrives as None, not as a zero-filled tensor
        return res[0], n_valid, res[1]

    @staticmethod
    def backward(ctx, g_sum, _g_n, g_loss):
        lib = L.load()
        prec, score, ids_rows, ids_cols, lm_rows, lm_cols, pop_prob = ctx.tensors
        g_sum = None if g_sum is None else g_sum.reshape(1).float().contiguous()
        g_loss = None if g_loss is None else g_loss.reshape(1).float().contiguous()
        d_prec = torch.empty_like(prec)
        d_score = torch.empty_like(score)
        L.check(lib.iisan_inbatch_ce_backward(C.byref(ctx.desc), _p(prec), _p(score), _p(ids_rows), _p(ids_cols), _p(lm_rows),
                                              _p(lm_cols), _p(pop_prob), _p(ctx.ws), ctx.ws.numel(), _p(g_sum), _p(g_loss),
                                              _p(ctx.n_valid), _p(d_prec), _p(d_score), _stream()), "iisan_inbatch_ce_backward")
        return d_prec, d_score, None, None, None, None, None, None, None


def inbatch_ce_masks(ids_rows, ids_cols, lm_rows, lm_cols, user_offset=0, fast=False):
    """uint8 [B*L, C] mask bits straight from the CUDA path (parity probe).  ``fast``: the bit-per-(user, column) masks of
    the tensor-core CE (bit0 = masked incl. the label exception, bit2 = label, bit3 = row valid)."""
    lib = L.load()
    L.require_cuda(ids_rows, "ids")
    b, l = lm_rows.shape
    bc = lm_cols.shape[0]
    if fast:
        desc = make_ce_desc(b, bc, l, 64, user_offset, L.COMPUTE_BF16)
        ws = _workspace(lib.iisan_inbatch_ce_workspace_bytes(C.byref(desc)), ids_rows.device)
        out = torch.empty(b * l, bc * (l + 1), dtype=torch.uint8, device=ids_rows.device)
        L.check(lib.iisan_inbatch_ce_masks_fast(C.byref(desc), _p(ids_rows.contiguous().view(-1)), _p(ids_cols.contiguous().view(-1)),
                                                _p(lm_rows.contiguous().float()), _p(lm_cols.contiguous().float()), _p(ws), ws.numel(),
                                                _p(out), _stream()), "iisan_inbatch_ce_masks_fast")
        return out
    desc = make_ce_desc(b, bc, l, 32, user_offset, 0)
    out = torch.empty(b * l, bc * (l + 1), dtype=torch.uint8, device=ids_rows.device)
    L.check(lib.iisan_inbatch_ce_masks(C.byref(desc), _p(ids_rows.contiguous().view(-1)), _p(ids_cols.contiguous().view(-1)),
                                       _p(lm_rows.contiguous().float()), _p(lm_cols.contiguous().float()), _p(out), _stream()),
            "iisan_inbatch_ce_masks")
    return out


def pack_states(states, sel):
    """out[n, len(sel), d] bf16 = round(states[..., sel, :]) through iisan_pack_states (states [..., layers, d], any stored dtype)."""
    lib = L.load()
    L.require_cuda(states, "hidden states")
    states = states.contiguous()
    layers, d = states.shape[-2], states.shape[-1]
    n = states.numel() // (layers * d)
    sel_t = sel if torch.is_tensor(sel) else torch.as_tensor(list(sel), dtype=torch.int32, device=states.device)
    out = torch.empty(n, len(sel), d, dtype=torch.bfloat16, device=states.device)
    L.check(lib.iisan_pack_states(_p(states), L.torch_dtype_code(states.dtype), n, layers, d, _p(sel_t), len(sel), _p(out), _stream()),
            "iisan_pack_states")
    return out


def gather_states(table, ids, sel, out=None):
    """out[n, len(sel), d] = table[ids, sel, :] (zeros for id 0) through iisan_gather_states.  ``out``: write into this buffer
    (the pipelined runner gathers the next batch into the static inputs of a captured step)."""
    lib = L.load()
    L.require_cuda(ids, "ids")
    n_items, layers, d = table.shape
    ids = ids.contiguous().view(-1)
    # a device int32 tensor is used as is (no H2D copy: legal during CUDA-graph capture)
    sel_t = sel if torch.is_tensor(sel) else torch.as_tensor(sel, dtype=torch.int32, device=ids.device)
    if out is None:
        out = torch.empty(ids.numel(), len(sel), d, dtype=table.dtype, device=ids.device)
    elif out.shape != (ids.numel(), len(sel), d) or out.dtype != table.dtype or not out.is_contiguous() or out.device != table.device:
        raise ValueError("gather_states: `out` must be a contiguous [n, len(sel), d] tensor of the table's dtype and device")
    L.check(lib.iisan_gather_states(_p(table), L.torch_dtype_code(table.dtype), n_items, layers, d, _p(ids), ids.numel(),
                                    _p(sel_t), len(sel), _p(out), _stream()), "iisan_gather_states")
    return out
