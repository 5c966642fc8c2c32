"""``from av_separation.model import ...`` as the reference's tests do (tests/test_model.py:13-20)."""
from avsep_b200.model import (AudioEncoder, VisualEncoder, CrossModalFusion, CrossAttentionLayer,  # noqa: F401
                              SeparationDecoder, AVSeparationTransformer, PositionalEncoding)
