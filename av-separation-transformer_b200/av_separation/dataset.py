"""``from av_separation.dataset import SyntheticAVDataset`` (reference src/av_separation/dataset.py:20-151) resolves
to the GPU-backed mirror: same constructor, ``__len__`` / ``__getitem__`` contract and item keys; items are device
tensors produced by ``avsep_synth_batch``."""
from avsep_b200.dataset import SyntheticAVDataset, evaluate_separation  # noqa: F401
