"""CPU restatement of the steps either side of the forward path.  TEST INFRASTRUCTURE ONLY:
imported by tests/, tests/golden/make_synth_golden.py and nothing on the product path.

  * input synthesis   SyntheticAVDataset.__getitem__ / _stft / _make_lip_frame   (reference dataset.py:70-151)
  * SNR evaluation    snr_db / evaluate_separation / _permutation_snr            (reference demo.py:25-80)
                      si_snr                                                     (reference losses.py:14-42)

Pinned against the reference itself: tests/golden/synth_*.npz hold items produced by the real SyntheticAVDataset and
scalars produced by the real demo.py / losses.py functions (generator: tests/golden/make_synth_golden.py);
tests/test_synth_oracle.py checks this file against them.

The random draws are separated from the arithmetic (``draw_item``) because the GPU kernels take the drawn values as
inputs: numpy's PCG64 + ziggurat stream is consumed on the host, in the reference's order, and everything that
costs time (waveforms, 3 x 63 windowed FFTs per item, energies, frame painting) runs on the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from itertools import permutations

import numpy as np


@dataclass(frozen=True)
class SynthConfig:                       # dataset.py:33-45 (constructor defaults)
    sample_rate: int = 8000
    duration: float = 1.0
    n_fft: int = 512
    hop_length: int = 128
    num_frames: int = 25
    frame_h: int = 32
    frame_w: int = 32
    speaker_freqs: tuple = (220.0, 440.0)

    @property
    def num_speakers(self):
        return len(self.speaker_freqs)

    @property
    def num_samples_audio(self):        # dataset.py:59
        return int(self.sample_rate * self.duration)

    @property
    def freq_bins(self):                # dataset.py:63
        return self.n_fft // 2 + 1

    @property
    def T(self):                        # dataset.py:65
        return 1 + self.num_samples_audio // self.hop_length

    @property
    def patch(self):                    # dataset.py:141-142 (lip region = centre 50 %)
        h0, h1 = self.frame_h // 4, 3 * self.frame_h // 4
        w0, w1 = self.frame_w // 4, 3 * self.frame_w // 4
        return h0, h1, w0, w1


def draw_item(cfg: SynthConfig, idx: int):
    """The reference's random draws for item ``idx`` in the reference's order (dataset.py:71-98,145).

    Returns amps (S,) f64, freqs (S,) f64 (jittered), phases (S,) f64, noise (S, num_frames, ph, pw) f32.
    """
    rng = np.random.default_rng(idx)                                   # dataset.py:71
    S = cfg.num_speakers
    amps = rng.uniform(0.3, 1.0, size=S)                               # dataset.py:74
    freqs = np.empty(S)
    phases = np.empty(S)
    for i, f in enumerate(cfg.speaker_freqs):                          # dataset.py:78-83
        freqs[i] = f * rng.uniform(0.95, 1.05)
        phases[i] = rng.uniform(0, 2 * math.pi)
    h0, h1, w0, w1 = cfg.patch
    noise = np.empty((S, cfg.num_frames, h1 - h0, w1 - w0), np.float32)
    for s in range(S):                                                 # dataset.py:95-104: speaker-major, frame-minor
        for fi in range(cfg.num_frames):
            noise[s, fi] = rng.normal(0, 0.05, (h1 - h0, w1 - w0)).astype(np.float32)   # dataset.py:145
    return amps, freqs, phases, noise


def waveforms(cfg: SynthConfig, amps, freqs, phases):
    """clean (S, n) f32 and mixed (n,) f32 (dataset.py:60,78-85)."""
    n = cfg.num_samples_audio
    t = np.linspace(0, cfg.duration, n, endpoint=False)
    clean = [(a * np.sin(2 * math.pi * f * t + p)).astype(np.float32) for a, f, p in zip(amps, freqs, phases)]
    mixed = sum(clean).astype(np.float32)
    return np.stack(clean, 0), mixed


def stft_mag(cfg: SynthConfig, audio: np.ndarray) -> np.ndarray:
    """Hann-windowed magnitude STFT, frames zero-padded at the tail (dataset.py:122-135).  (freq_bins, T) f32."""
    window = np.hanning(cfg.n_fft)
    out = np.empty((cfg.freq_bins, cfg.T), np.float32)
    for i in range(cfg.T):
        start = i * cfg.hop_length
        frame = np.zeros(cfg.n_fft, np.float32)
        chunk = audio[start:start + cfg.n_fft]
        frame[:len(chunk)] = chunk
        frame *= window
        out[:, i] = np.abs(np.fft.rfft(frame))
    return out


def lip_frames(cfg: SynthConfig, clean: np.ndarray, noise: np.ndarray) -> np.ndarray:
    """(S * num_frames, H, W) f32: brightness = min(1, 20 * mean(x^2)) over the frame's audio span, plus noise,
    clipped to [0, 1], painted into the centre patch (dataset.py:91-111,137-147)."""
    S, n = clean.shape
    step = n // cfg.num_frames                                         # dataset.py:93
    h0, h1, w0, w1 = cfg.patch
    out = np.zeros((S, cfg.num_frames, cfg.frame_h, cfg.frame_w), np.float32)
    for s in range(S):
        for fi in range(cfg.num_frames):
            a = fi * step
            b = min(a + step, n)
            energy = float(np.mean(clean[s, a:b] ** 2))
            brightness = min(1.0, energy * 20.0)
            out[s, fi, h0:h1, w0:w1] = np.clip(brightness + noise[s, fi], 0, 1)
    return out.reshape(S * cfg.num_frames, cfg.frame_h, cfg.frame_w)


def synth_item(cfg: SynthConfig, idx: int):
    """One dataset item: mixed_spec (F,T), lip_frames (S*nf,H,W), clean_specs (S,F,T)  (dataset.py:70-119)."""
    amps, freqs, phases, noise = draw_item(cfg, idx)
    clean, mixed = waveforms(cfg, amps, freqs, phases)
    return {
        "mixed_spec": stft_mag(cfg, mixed),
        "lip_frames": lip_frames(cfg, clean, noise),
        "clean_specs": np.stack([stft_mag(cfg, c) for c in clean], 0),
    }


# ---------------------------------------------------------------------------------------------------------------
# SNR evaluation after the path
# ---------------------------------------------------------------------------------------------------------------
def snr_db(signal: np.ndarray, noise: np.ndarray, eps: float = 1e-8) -> float:          # demo.py:25-29
    sig_power = np.mean(signal ** 2)
    noise_power = np.mean(noise ** 2)
    return 10 * math.log10(sig_power / (noise_power + eps) + eps)


def permutation_snr(separated: np.ndarray, targets: np.ndarray) -> float:               # demo.py:67-80
    S = separated.shape[0]
    best = -1e9
    for perm in permutations(range(S)):
        val = np.mean([snr_db(targets[t], separated[s] - targets[t]) for s, t in zip(perm, range(S))])
        if val > best:
            best = val
    return float(best)


def input_snrs(mixed: np.ndarray, targets: np.ndarray):                                  # demo.py:55-58
    return [snr_db(targets[s], mixed - targets[s]) for s in range(targets.shape[0])]


def si_snr_rows(estimate: np.ndarray, target: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """Per-row SI-SNR in dB (losses.py:14-42 before the final .mean()); rows = leading axis, the rest flattened."""
    est = estimate.reshape(estimate.shape[0], -1).astype(np.float32)
    tgt = target.reshape(target.shape[0], -1).astype(np.float32)
    est = est - est.mean(-1, keepdims=True)
    tgt = tgt - tgt.mean(-1, keepdims=True)
    dot = (est * tgt).sum(-1, keepdims=True)
    tgt_energy = (tgt * tgt).sum(-1, keepdims=True) + np.float32(eps)
    proj = dot / tgt_energy * tgt
    noise = est - proj
    return 10 * np.log10((proj * proj).sum(-1) / ((noise * noise).sum(-1) + np.float32(eps)) + np.float32(eps))
