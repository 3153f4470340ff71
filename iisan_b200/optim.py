"""Optimizer step of the hot path (Code_Cached/run.py:260-307, :383-385).

``param_groups`` reproduces the reference's name-routed learning-rate groups:
  names containing 'cv'  : plain 'fc'/'classifier' (not 'fc_') -> recsys lr ; adapters -> adapter_cv_lr ;
                           everything else (fc_cv, side_gate_params_cv) -> fine_tune_lr_image
  names containing 'bert': plain 'fc' (not 'fc_') -> recsys lr ; adapters -> adapter_bert_lr ;
                           everything else (fc_bert) -> fine_tune_lr_text
  'mm_adapter'           : adapter_cv_lr
  everything else        : lr
"""
from __future__ import annotations

import ctypes as C

import torch


def param_groups(model, args):
    text_enc, image_net, recsys, ad_cv, ad_text = [], [], [], [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if "cv" in name:
            if ("fc" in name and "fc_" not in name) or "classifier" in name or "decoder_pred" in name:
                recsys.append(p)
            elif "adapter" not in name and "lora" not in name:
                image_net.append(p)
            else:
                ad_cv.append(p)
        elif "bert" in name:
            if "fc" in name and "fc_" not in name:
                recsys.append(p)
            elif "adapter" not in name and "lora" not in name:
                text_enc.append(p)
            else:
                ad_text.append(p)
        elif "mm_adapter" in name:
            ad_cv.append(p)
        else:
            recsys.append(p)
    return [
        {"params": text_enc, "lr": args.fine_tune_lr_text},
        {"params": image_net, "lr": args.fine_tune_lr_image},
        {"params": recsys, "lr": args.lr},
        {"params": ad_cv, "lr": args.adapter_cv_lr},
        {"params": ad_text, "lr": args.adapter_bert_lr},
    ]


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (betas, eps; weight_decay 0, amsgrad off -- what the reference uses, Code_Cached/run.py:301-307)
    through iisan_adam_step: one launch for the 146 tensors of the base model instead of one multi-tensor launch per
    learning-rate group, step counter on the device (CUDA-graph capturable).  ``param_groups`` as for torch.optim.Adam."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        # the full key set of torch.optim.Adam's param groups, so that a saved state loads into either optimizer
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                                      capturable=True, differentiable=False, fused=None, decoupled_weight_decay=False))
        self._table = None
        self._key = None
        self._step_dev = None

    def _build(self):
        from . import _lib as L
        entries = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                    raise L.IisanLibraryError("FusedAdam needs contiguous fp32 CUDA parameters")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                entries.append((p, grad, st["exp_avg"], st["exp_avg_sq"], float(g["lr"])))
        key = tuple((p.data_ptr(), gr.data_ptr(), lr) for p, gr, _, _, lr in entries)
        if key != self._key:
            arr = (L.AdamTensor * len(entries))()
            for i, (p, gr, m, v, lr) in enumerate(entries):
                arr[i].param, arr[i].grad, arr[i].exp_avg, arr[i].exp_avg_sq = p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr()
                arr[i].numel, arr[i].lr = p.numel(), lr
            self._table, self._key, self._keep = arr, key, [e[1] for e in entries]
        return len(entries)

    # ---- checkpoint hand-over with torch.optim.Adam (Code_Cached/run.py:234-243 restores `optimizer` from epoch-N.pt,
    #      data_utils/utils.py:104-110 saves it): same state_dict layout -- exp_avg, exp_avg_sq and a per-parameter `step` ----
    def state_dict(self):
        sd = super().state_dict()
        step = float(self._step_dev.item()) if self._step_dev is not None else 0.0
        for st in sd["state"].values():
            st["step"] = torch.tensor(step, dtype=torch.float32)
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        step = None
        for p, st in self.state.items():
            if "step" in st:
                v = st.pop("step")
                step = float(v.item() if torch.is_tensor(v) else v) if step is None else step
        if step is not None:
            dev = self.param_groups[0]["params"][0].device
            self._step_dev = torch.full((1,), step, dtype=torch.float32, device=dev)
        self._key = None                    # rebuild the pointer table (the moment tensors were replaced)

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib as L
        if closure is not None:
            raise NotImplementedError("FusedAdam does not take a closure")
        n = self._build()
        if n == 0:
            return None
        lib = L.load()
        dev = self.param_groups[0]["params"][0].device
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        b1, b2 = self.param_groups[0]["betas"]
        for g in self.param_groups:          # ONE launch for every group: only the learning rate may differ between groups
            if tuple(g["betas"]) != (b1, b2) or g["eps"] != self.param_groups[0]["eps"]:
                raise ValueError("FusedAdam: all parameter groups must share betas and eps (only lr is per group)")
            if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
                raise ValueError("FusedAdam implements plain Adam: weight_decay, amsgrad and maximize are not supported")
        L.check(lib.iisan_adam_step(self._table, n, b1, b2, self.param_groups[0]["eps"], C.c_void_p(self._step_dev.data_ptr()), 1,
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "iisan_adam_step")
        return None
