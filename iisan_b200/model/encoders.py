"""Mirror of the reference `model.encoders` (Code_Cached/model/encoders.py) for the Cached path.

User_Encoder is on the hot path (its forward runs in the CUDA library).  The ViT/BERT wrappers exist
only so that ModelMM / IISANAdaptedMModel can be constructed exactly as Code_Cached/run.py:138,182-183
does: in the Cached path the frozen backbones never run (run.py discards them), so their forward raises.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.init import constant_, xavier_normal_

from .modules import TransformerEncoder


class User_Encoder(nn.Module):
    """Code_Cached/model/encoders.py:37-58."""

    def __init__(self, item_num, max_seq_len, item_dim, num_attention_heads, dropout, n_layers):
        super().__init__()
        self.transformer_encoder = TransformerEncoder(n_vocab=item_num, n_position=max_seq_len, d_model=item_dim,
                                                      n_heads=num_attention_heads, dropout=dropout, n_layers=n_layers)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(module):
        # xavier-normal Linear/Embedding weights, zero biases (encoders.py:45-51)
        if isinstance(module, (nn.Embedding, nn.Linear)):
            xavier_normal_(module.weight.data)
            if getattr(module, "bias", None) is not None:
                constant_(module.bias.data, 0)

    def forward(self, input_embs, log_mask, local_rank=None, seq_len=None):
        """Reference call: ``user_encoder(input_embs[:, :-1, :], log_mask, local_rank)`` (CC/model/model.py:76-79).  ``seq_len``
        (extension): pass the unsliced [B, S, E] block and the number of leading slots to use; same result, the slice and its
        autograd backward (a zero fill and a strided copy) disappear from the step."""
        return self.transformer_encoder(input_embs, log_mask, None, seq_len=seq_len)


def _off_path(what):
    raise NotImplementedError(
        f"{what} belongs to the uncached backbone path, which IISAN(Cached) never executes "
        "(hidden states are precomputed; Code_Cached/run.py:182 replaces mm_encoder by the side-adapter network)")


class Vit_Encoder(nn.Module):
    def __init__(self, image_net):
        super().__init__()
        self.image_net = image_net
        self.activate = nn.GELU()

    def forward(self, item_content):
        _off_path("Vit_Encoder.forward")


class Text_Encoder(nn.Module):
    def __init__(self, bert_model, item_embedding_dim, word_embedding_dim):
        super().__init__()
        self.bert_model = bert_model
        self.fc = nn.Linear(word_embedding_dim, item_embedding_dim)
        self.activate = nn.GELU()

    def forward(self, text):
        _off_path("Text_Encoder.forward")


class Bert_Encoder(nn.Module):
    def __init__(self, args, bert_model):
        super().__init__()
        self.args = args
        if len(args.news_attributes) == 0:
            raise ValueError("news_attributes must not be empty")
        self.text_encoders = nn.ModuleDict({"title": Text_Encoder(bert_model, args.embedding_dim, args.word_embedding_dim)})

    def forward(self, news):
        _off_path("Bert_Encoder.forward")


class MM_Encoder(nn.Module):
    """Holder of the two backbone wrappers (encoders.py:6-14); IISANAdaptedMModel takes their heads."""

    def __init__(self, args, image_net, bert_model):
        super().__init__()
        self.cv_encoder = Vit_Encoder(image_net=image_net)
        self.bert_encoder = Bert_Encoder(args=args, bert_model=bert_model)

    def forward(self, sample_items_images, sample_items_text):
        _off_path("MM_Encoder.forward")
