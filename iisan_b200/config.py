"""Reference hyper-parameters of the shipped IISAN(Cached) run as the ``args``
namespace the model constructors read (Code_Cached/parameters.py + Code_Cached/scripts/run_IISAN.py)."""
from __future__ import annotations

import argparse


def default_args(**overrides) -> argparse.Namespace:
    ns = argparse.Namespace(
        max_seq_len=10, min_seq_len=5, l2_weight=0, embedding_dim=64, num_attention_heads=2, drop_rate=0.1,
        transformer_block=2, modality="intra_inter", fusion_method="gated", remove_first="None",
        news_attributes=["title"], num_words_title=30, num_words_abstract=50, num_words_body=50,
        word_embedding_dim=768, side_adapter_vit_list="1,3,5,7,9,11", side_adapter_bert_list="1,3,5,7,9,11",
        cv_adapter_down_size=64, bert_adapter_down_size=64, adapter_dropout_rate=0.1, adapter_activation="RELU",
        lr=2e-4, adapter_cv_lr=1e-4, adapter_bert_lr=1e-4, fine_tune_lr_image=1e-4, fine_tune_lr_text=5e-5,
        adding_adapter_to="all", use_scale="half", batch_size=64)
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns
