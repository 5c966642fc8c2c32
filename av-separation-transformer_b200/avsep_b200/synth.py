"""SyntheticAVDataset-shaped inputs generated directly as torch tensors (any device).

Shapes and value ranges follow the reference's dataset (src/av_separation/dataset.py:33-45,70-151):
``mixed_spec`` (B, F, T) fp32 >= 0 and heavy-tailed -- each utterance is two narrow spectral lines (the two
speakers' sinusoids through a Hann STFT) peaking around 30-110 over a ~1e-3 floor; ``lip_frames`` (B, N, H, W)
fp32 in [0, 1], zero outside the centre half of the frame, brightness + N(0, 0.05) noise inside.
This is a data generator for benchmarks; it is not part of the forward path.
"""
from __future__ import annotations

import torch


def shapes_for(sample_rate: int = 8000, duration: float = 1.0, n_fft: int = 512, hop: int = 128,
               frames_per_second: int = 25, num_speakers: int = 2):
    """(F, T, N) as the dataset derives them (dataset.py:63-65,114)."""
    n = int(sample_rate * duration)
    return n_fft // 2 + 1, 1 + n // hop, num_speakers * int(round(frames_per_second * duration))


def synthetic_batch(B: int, F: int = 257, T: int = 63, N: int = 50, Hh: int = 32, Ww: int = 32, seed: int = 0,
                    device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    f = torch.arange(F, dtype=torch.float32).view(1, F, 1)
    mixed = torch.zeros(B, F, T)
    for _ in range(2):
        centre = (0.04 + 0.26 * torch.rand(B, 1, 1, generator=g)) * F
        amp = 30.0 + 80.0 * torch.rand(B, 1, 1, generator=g)
        width = 0.8 + 0.8 * torch.rand(B, 1, 1, generator=g)
        env = 1.0 + 0.05 * torch.randn(B, 1, T, generator=g)
        mixed += amp * torch.exp(-0.5 * ((f - centre) / width) ** 2) * env
    mixed += torch.randn(B, F, T, generator=g).abs() * 1e-3
    frames = torch.zeros(B, N, Hh, Ww)
    h0, h1, w0, w1 = Hh // 4, 3 * Hh // 4, Ww // 4, 3 * Ww // 4
    bright = torch.rand(B, N, 1, 1, generator=g)
    noise = 0.05 * torch.randn(B, N, h1 - h0, w1 - w0, generator=g)
    frames[:, :, h0:h1, w0:w1] = (bright + noise).clamp_(0.0, 1.0)
    return mixed.to(device), frames.to(device)
