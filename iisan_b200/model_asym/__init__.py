"""`model` package of Code_Cached_Asym (IISAN-Versa): per-modality widths / depths, group layer-drop and
down_project dimension alignment (Code_Cached_Asym/model/model.py:257-429).

Drop-in:  sys.modules['model'] = iisan_b200.model_asym
"""
from ..model.encoders import Bert_Encoder, MM_Encoder, Text_Encoder, User_Encoder, Vit_Encoder
from ..model.model import IISANAdaptedMModel as _SAN
from ..model.model import Model, ModelMM
from ..model.modules import (AdapterBlock, FusedLinear, MultiHeadedAttention, PositionwiseFeedForward, TransformerBlock,
                             TransformerEncoder)


class IISANAdaptedMModel(_SAN):
    def __init__(self, mm_model, args):
        super().__init__(mm_model, args, asym=True)


__all__ = ["ModelMM", "Model", "IISANAdaptedMModel", "AdapterBlock", "User_Encoder", "MM_Encoder", "Vit_Encoder",
           "Bert_Encoder", "Text_Encoder", "TransformerEncoder", "TransformerBlock", "MultiHeadedAttention",
           "PositionwiseFeedForward", "FusedLinear"]
