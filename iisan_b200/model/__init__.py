"""`model` package of the reference, rebuilt on the B200-native kernels (Code_Cached semantics).

Drop-in:  sys.modules['model'] = iisan_b200.model   (before `from model import *` in run.py)
"""
from .encoders import Bert_Encoder, MM_Encoder, Text_Encoder, User_Encoder, Vit_Encoder
from .model import IISANAdaptedMModel, Model, ModelMM
from .modules import (AdapterBlock, FusedLinear, MultiHeadedAttention, PositionwiseFeedForward, TransformerBlock,
                      TransformerEncoder)

__all__ = ["ModelMM", "Model", "IISANAdaptedMModel", "AdapterBlock", "User_Encoder", "MM_Encoder", "Vit_Encoder",
           "Bert_Encoder", "Text_Encoder", "TransformerEncoder", "TransformerBlock", "MultiHeadedAttention",
           "PositionwiseFeedForward", "FusedLinear"]
