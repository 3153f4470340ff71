"""TEST INFRASTRUCTURE -- torch emulation of the fast mode's rounding points.

The fast mode (IISAN_COMPUTE_BF16) rounds GEMM operands and the stage stash to bf16 and accumulates in fp32.  Against
the fp32 oracle its *gradients* differ by ReLU activations that flip under the rounding noise (a per-element effect that
only averages out over large batches), so the backward kernels are pinned against this emulation instead: the oracle's
forward (oracle/iisan_oracle.py, same reference citations) with a straight-through bf16 rounding inserted exactly where
the CUDA path rounds.  Loss / embeddings of the fast mode are still compared with the fp32 reference itself.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import iisan_oracle as O


def rb(x):
    """bf16 rounding with a straight-through gradient."""
    return x + (x.detach().bfloat16().float() - x.detach())


def rt(x):
    """TF32 rounding (cvt.rna.tf32.f32: nearest, ties away from zero, 10 explicit mantissa bits), straight-through gradient."""
    bits = x.detach().contiguous().view(torch.int32)
    r = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return x + (r - x.detach())


def _lin(x, w, b=None):
    return F.linear(x, rb(w), b)


def user_encoder_forward_emul(P, embs, log_mask, cfg, prefix="user_encoder.transformer_encoder."):
    """oracle.user_encoder_forward (CC/model/encoders.py:53-58, CC/model/modules.py:14-18, 54-64, 89-96) with the operands of
    every linear layer rounded to TF32, as the fused fast-mode kernel does (user_encoder_fused.cu, mma.sync TF32 tiles)."""
    B, L, E = embs.shape
    H = cfg.heads; dk = E // H
    keep = torch.tril((log_mask != 0)[:, None, None, :].expand(B, 1, L, L))
    att_mask = torch.where(keep, 0.0, O.ATT_NEG)
    lin = lambda x, w, b=None: F.linear(rt(x), rt(P[w]), None if b is None else P[b])
    x = O._ln(embs + P[prefix + "position_embedding.weight"][None, :L], P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"])
    for b in range(cfg.blocks):
        p = f"{prefix}transformer_blocks.{b}."
        a = p + "multi_head_attention."
        q = lin(x, a + "w_Q.weight").view(B, L, H, dk).transpose(1, 2)
        k = lin(x, a + "w_K.weight").view(B, L, H, dk).transpose(1, 2)
        v = lin(x, a + "w_V.weight").view(B, L, H, dk).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-2, -1)) / (dk ** 0.5) + att_mask
        ctx = torch.matmul(torch.softmax(att, dim=-1), v).transpose(1, 2).reshape(B, L, E)
        x = O._ln(x + lin(ctx, a + "fc.weight"), P[a + "layer_norm.weight"], P[a + "layer_norm.bias"])
        f = p + "feed_forward."
        y = lin(F.relu(lin(x, f + "w_1.weight", f + "w_1.bias")), f + "w_2.weight", f + "w_2.bias")
        x = O._ln(x + y, P[f + "layer_norm.weight"], P[f + "layer_norm.bias"])
    return x


def lr_path(cfg, fused_chain):
    """True where the third-generation path runs (san_lr_eligible, san_lr.cu): symmetric chain, d == 768, E == 64, CC heads."""
    return bool(fused_chain and not cfg.asym and cfg.d_text == 768 and cfg.embedding_dim == 64)


def san_forward_emul(P, image, text, cfg, prefix="mm_encoder.", fused_chain=False):
    """``fused_chain``: the fused chain kernel keeps last_s in fp32 registers when it fuses the next stage (only the stash
    and the final stage, which feeds the heads, are rounded).  On the third-generation path (``lr_path``) the two head layers
    are applied as ONE matrix M = bf16(bf16(W_pre) bf16(W_fc)) with the bias W_pre b_fc + b_pre in fp32."""
    h_cv = image.reshape(-1, image.shape[-2], image.shape[-1]).float()
    h_tx = text.reshape(-1, text.shape[-2], text.shape[-1]).float()
    N = h_cv.shape[0]
    dev = h_cv.device
    d_mm = min(cfg.d_text, cfg.d_img) if cfg.asym else cfg.d_text
    if cfg.remove_first == "TRUE":
        last_cv, last_tx = h_cv[:, 0], h_tx[:, 0]
    else:
        last_cv = torch.zeros(N, cfg.d_img, device=dev); last_tx = torch.zeros(N, cfg.d_text, device=dev)
    last_mm = torch.zeros(N, d_mm, device=dev)

    plan = O.stage_plan(cfg)

    act = F.gelu if getattr(cfg, "adapter_activation", "RELU") == "GELU" else F.relu      # GELU: exact erf form on the fp32 pre-activation

    def adapter(pfx, x, final):
        z = rb(act(_lin(x, P[pfx + ".fc_down.weight"], P[pfx + ".fc_down.bias"])))
        y = _lin(z, P[pfx + ".fc_up.weight"], P[pfx + ".fc_up.bias"]) + x
        return y if (fused_chain and not final) else rb(y)

    for si, (ta, tl, ia, il, mi) in enumerate(plan):
        final = si == len(plan) - 1
        if ia is not None:
            g = O._gate(P[f"{prefix}side_gate_params_cv.{ia}"])
            x_cv = rb(g * h_cv[:, il] + (1 - g) * last_cv)
        if ta is not None:
            g = O._gate(P[f"{prefix}side_gate_params_text.{ta}"])
            x_tx = rb(g * h_tx[:, tl] + (1 - g) * last_tx)
        if ta is not None:
            last_tx = adapter(f"{prefix}bert_adapter_list.{ta}", x_tx, final)
        if ia is not None:
            last_cv = adapter(f"{prefix}cv_adapter_list.{ia}", x_cv, final)
        if mi is not None:
            mm_tx, mm_cv = h_tx[:, tl], h_cv[:, il]
            if cfg.asym and cfg.d_text > cfg.d_img:
                mm_tx = _lin(rb(mm_tx), P[f"{prefix}down_project_list.{mi}.weight"], P[f"{prefix}down_project_list.{mi}.bias"])
            elif cfg.asym and cfg.d_img > cfg.d_text:
                mm_cv = _lin(rb(mm_cv), P[f"{prefix}down_project_list.{mi}.weight"], P[f"{prefix}down_project_list.{mi}.bias"])
            g = O._gate(P[f"{prefix}side_gate_params_mm.{mi}"])
            x_mm = rb(last_mm + g * mm_cv + (1 - g) * mm_tx)
            last_mm = adapter(f"{prefix}mm_adapter_list.{mi}", x_mm, final)
    if lr_path(cfg, fused_chain):
        def head(fc, pre, last):
            M = rb(rb(P[f"{prefix}{pre}.weight"]) @ rb(P[f"{prefix}{fc}.weight"]))
            c = P[f"{prefix}{pre}.weight"] @ P[f"{prefix}{fc}.bias"] + P[f"{prefix}{pre}.bias"]
            return last @ M.T + c
        return head("fc_cv", "cv_pre_fc", last_cv), head("fc_bert", "bert_pre_fc", last_tx), head("fc_mm", "fc_mm_down", last_mm)
    lin = lambda n, x: _lin(x, P[f"{prefix}{n}.weight"], P[f"{prefix}{n}.bias"])
    e_tx = lin("bert_pre_fc", rb(lin("fc_bert", last_tx)))
    e_cv = lin("cv_pre_fc", rb(lin("fc_cv", last_cv)))
    e_mm = lin("fc_mm_down", rb(lin("fc_mm", last_mm)))
    return e_cv, e_tx, e_mm


def train_step_grads_emul(params_np, batch, pop_prob, cfg, ce_bf16=False, fused_chain=False):
    """Like oracle.train_step_grads with the fast mode's rounding points (CPU fp32 torch)."""
    P = O.params_to_torch(params_np)
    ids, lm = batch["ids"], batch["log_mask"]
    B, S = ids.shape
    image = torch.as_tensor(batch["image"]); text = torch.as_tensor(batch["text"])
    debias = torch.log(torch.from_numpy(pop_prob)[torch.from_numpy(ids.reshape(-1))])
    e_cv, e_tx, e_mm = san_forward_emul(P, image, text, cfg, fused_chain=fused_chain)
    cat = torch.cat([e_cv, e_tx, e_mm], dim=1)
    if cat.shape[1] % 16 == 0 and cfg.embedding_dim % 16 == 0:      # linear_tf32_supported: com_dense runs as TF32 tiles
        score = F.linear(rt(cat), rt(P["com_dense.weight"]), P["com_dense.bias"])
    else:
        score = F.linear(cat, P["com_dense.weight"], P["com_dense.bias"])
    embs = score.view(B, S, cfg.embedding_dim)
    tf32_ue = cfg.embedding_dim == 64 and S - 1 == 10 and cfg.heads <= 4          # ue_fused_supported (user_encoder_fused.cu)
    ue = user_encoder_forward_emul if tf32_ue else O.user_encoder_forward
    prec = ue(P, embs[:, :-1], torch.from_numpy(lm), cfg).reshape(-1, cfg.embedding_dim)
    if ce_bf16:
        loss, _ = O.inbatch_ce(rb(prec), rb(score), debias, ids, lm, ids, lm)
    else:
        loss, _ = O.inbatch_ce(prec, score, debias, ids, lm, ids, lm)
    loss.backward()
    grads = {k: (None if v.grad is None else v.grad.detach().numpy()) for k, v in P.items()}
    return {"loss": loss.detach().numpy(), "score_embs": score.detach().numpy()}, grads
