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
