"""avsep_b200 -- B200-native (sm_100a) forward path of the AV-Separation-Transformer.

Host-side mirror of the reference interface (``model.py``) over the C ABI of libavsep.so.
"""
from .model import (AudioEncoder, VisualEncoder, CrossModalFusion, CrossAttentionLayer, SeparationDecoder,
                    AVSeparationTransformer, PositionalEncoding)
from .engine import Engine, EngineConfig, STAGE_NAMES
from .dataset import SyntheticAVDataset, evaluate_separation

__all__ = ["AudioEncoder", "VisualEncoder", "CrossModalFusion", "CrossAttentionLayer", "SeparationDecoder",
           "AVSeparationTransformer", "PositionalEncoding", "Engine", "EngineConfig", "STAGE_NAMES", "SyntheticAVDataset", "evaluate_separation"]
