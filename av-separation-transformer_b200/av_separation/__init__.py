"""Drop-in import name of the reference package (src/av_separation/__init__.py:6-22) for the forward path:
``from av_separation import AVSeparationTransformer`` resolves to the B200-native implementation, and
``SyntheticAVDataset`` to its GPU-backed mirror.  ``av_separation.losses`` exists for import compatibility only
(training objectives are out of scope)."""
from avsep_b200 import (AudioEncoder, VisualEncoder, CrossModalFusion, SeparationDecoder,  # noqa: F401
                        AVSeparationTransformer, SyntheticAVDataset)

__all__ = ["AudioEncoder", "VisualEncoder", "CrossModalFusion", "SeparationDecoder", "AVSeparationTransformer",
           "SyntheticAVDataset"]
