"""GPU-backed mirror of the reference's SyntheticAVDataset (src/av_separation/dataset.py).

Same constructor arguments, same ``__len__`` / ``__getitem__`` contract and item keys, plus ``batch(indices)`` which
is how it is meant to be used: the reference's per-item random draws are consumed on the host from numpy's
``default_rng(idx)`` in the reference's order (dataset.py:71-83,145 -- a few hundred numbers per item), and the
expensive part (waveforms, 3 x T windowed FFTs, energies, frame painting: a Python loop in the reference) runs in
``avsep_synth_batch`` on the device, producing tensors that are already where ``avsep_forward`` wants them.
There is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .engine import Engine, EngineConfig


class SyntheticAVDataset:
    def __init__(self, num_samples: int = 1000, sample_rate: int = 8000, duration: float = 1.0, n_fft: int = 512,
                 hop_length: int = 128, num_frames: int = 25, frame_h: int = 32, frame_w: int = 32,
                 speaker_freqs: tuple = (220.0, 440.0), seed: int = 42, device: int = 0, engine: Engine | None = None):
        self.num_samples = num_samples
        self.sample_rate = sample_rate
        self.duration = duration
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.num_frames = num_frames
        self.frame_h = frame_h
        self.frame_w = frame_w
        self.speaker_freqs = tuple(speaker_freqs)
        self.num_speakers = len(self.speaker_freqs)
        self.num_samples_audio = int(sample_rate * duration)          # dataset.py:59
        self.freq_bins = n_fft // 2 + 1                               # dataset.py:63
        self.T = 1 + self.num_samples_audio // hop_length             # dataset.py:65
        self.device = int(device)
        self._engine = engine
        self._geom = dict(num_samples_audio=self.num_samples_audio, duration=duration, n_fft=n_fft,
                          hop_length=hop_length, num_frames=num_frames, frame_h=frame_h, frame_w=frame_w,
                          num_speakers=self.num_speakers)

    @property
    def engine(self) -> Engine:
        if self._engine is None:   # any engine handle will do: synthesis does not touch model weights
            self._engine = Engine(EngineConfig(self.freq_bins, 64, 4, 1, 1, self.num_speakers, "bf16"), self.device)
        return self._engine

    def __len__(self) -> int:
        return self.num_samples

    def draws(self, idx: int):
        """Host side of one item: the reference's RNG stream, in its order (dataset.py:71-83, 95-104, 145)."""
        rng = np.random.default_rng(idx)
        S = self.num_speakers
        amps = rng.uniform(0.3, 1.0, size=S)
        freqs = np.empty(S)
        phases = np.empty(S)
        for i, f in enumerate(self.speaker_freqs):
            freqs[i] = f * rng.uniform(0.95, 1.05)
            phases[i] = rng.uniform(0, 2 * math.pi)
        ph = 3 * self.frame_h // 4 - self.frame_h // 4
        pw = 3 * self.frame_w // 4 - self.frame_w // 4
        noise = rng.normal(0, 0.05, (S, self.num_frames, ph, pw)).astype(np.float32)
        return amps, freqs, phases, noise

    def batch(self, indices, want_clean: bool = True):
        """dict of device tensors: mixed_spec (B,F,T), lip_frames (B,S*nf,H,W), clean_specs (B,S,F,T)."""
        d = [self.draws(int(i)) for i in indices]
        dev = torch.device("cuda", self.device)
        up = lambda k, dt: torch.from_numpy(np.stack([x[k] for x in d], 0)).to(dt).pin_memory().to(dev, non_blocking=True)
        amps, freqs, phases = up(0, torch.float64), up(1, torch.float64), up(2, torch.float64)
        noise = up(3, torch.float32)
        mixed, frames, clean = self.engine.synth_batch(self._geom, amps, freqs, phases, noise, want_clean)
        out = {"mixed_spec": mixed, "lip_frames": frames}
        if want_clean:
            out["clean_specs"] = clean
        return out

    def __getitem__(self, idx: int):
        b = self.batch([idx])
        return {k: v[0] for k, v in b.items()}


def evaluate_separation(model, dataset: SyntheticAVDataset, num_eval: int = 20, batch_size: int = 64):
    """Mirror of demo.py:evaluate_separation (demo.py:32-64) with synthesis, forward and the SNR reduction on the
    device: returns (mean input SNR, mean best-permutation output SNR) in dB."""
    n = min(num_eval, len(dataset))
    ins, outs = [], []
    for lo in range(0, n, batch_size):
        b = dataset.batch(range(lo, min(lo + batch_size, n)))
        separated, _ = model(b["mixed_spec"], b["lip_frames"])
        i_snr, o_snr, _, _ = model.engine.eval_snr(separated.contiguous(), b["clean_specs"], b["mixed_spec"])
        ins.append(i_snr.flatten())
        outs.append(o_snr)
    return float(torch.cat(ins).mean()), float(torch.cat(outs).mean())
