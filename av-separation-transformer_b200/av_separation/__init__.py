"""Drop-in import name of the reference package (src/av_separation/__init__.py:6-22) for the forward path:
``from av_separation import AVSeparationTransformer`` resolves to the B200-native implementation.
Dataset, losses and training utilities of the reference are out of scope."""
from avsep_b200 import (AudioEncoder, VisualEncoder, CrossModalFusion, SeparationDecoder,  # noqa: F401
                        AVSeparationTransformer)

__all__ = ["AudioEncoder", "VisualEncoder", "CrossModalFusion", "SeparationDecoder", "AVSeparationTransformer"]
