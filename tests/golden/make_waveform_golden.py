"""Generates tests/golden/waveform_stft.npz from the REAL reference (run in the build container only):
SyntheticAVDataset._stft (reference src/av_separation/dataset.py:122-135) on seeded waveforms, so that
oracle/waveform_oracle.stft_complex's magnitudes are pinned against the reference's own analysis."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/src")
from av_separation.dataset import SyntheticAVDataset   # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261018)
out = {}
for name, (sr, dur, n_fft, hop) in {"default": (8000, 1.0, 512, 128), "short": (4000, 0.25, 128, 32),
                                    "ragged": (8000, 0.3371, 256, 100)}.items():
    ds = SyntheticAVDataset(num_samples=1, sample_rate=sr, duration=dur, n_fft=n_fft, hop_length=hop)
    x = (0.3 * rng.standard_normal(ds.num_samples_audio)).astype(np.float32)
    out[name + "_wave"] = x
    out[name + "_mag"] = ds._stft(x)
    out[name + "_geom"] = np.array([n_fft, hop], dtype=np.int64)
np.savez_compressed(os.path.join(here, "waveform_stft.npz"), **out)
print({k: v.shape for k, v in out.items()})
