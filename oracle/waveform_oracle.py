"""CPU restatement of the waveform side of the path.  TEST INFRASTRUCTURE ONLY: imported by tests/
and tests/golden/make_waveform_golden.py, never by the product path.

  * stft_complex   SyntheticAVDataset._stft (reference dataset.py:122-135) with `np.abs` left off: frame i starts at
                   i*hop (no centring), zero-padded past the end, float32 `frame *= np.hanning(n_fft)`, np.fft.rfft.
                   PINNED: |stft_complex(x)| is checked against the real `_stft` of the reference
                   (tests/golden/waveform_stft.npz, generator tests/golden/make_waveform_golden.py).
  * istft_masked   weighted overlap-add inverse, y[n] = sum_i w[n-i*hop] irfft(m_i X_i)[n-i*hop] / sum_i w^2[n-i*hop].
                   The reference has no inverse (README.md:140 lists phase / iSTFT reconstruction as missing), so
                   this is PINNED on the named third-party algorithm instead: scipy.signal.istft with the reference's
                   analysis window / hop and boundary=False (tests/test_waveform_oracle.py::
                   test_inverse_is_scipy_signal_istft, <= 1e-9 relative), plus the identities istft(stft(x)) == x,
                   linearity in the masks, and a direct O(n_fft^2) inverse DFT (`irfft_direct`) for small sizes.
"""
from __future__ import annotations

import numpy as np


def num_frames(L: int, hop: int) -> int:            # dataset.py:65
    return 1 + L // hop


def stft_complex(audio: np.ndarray, n_fft: int = 512, hop: int = 128) -> np.ndarray:
    """(L,) float32 -> (F, T) complex64, the loop of dataset.py:125-134 keeping the phase (numpy >= 2 transforms a
    float32 frame in single precision, so the reference's spectrum is complex64 before `np.abs`)."""
    audio = np.asarray(audio, dtype=np.float32)
    window = np.hanning(n_fft)                       # dataset.py:124 (float64)
    T = num_frames(len(audio), hop)
    cols = []
    for i in range(T):
        start = i * hop
        frame = np.zeros(n_fft, dtype=np.float32)
        chunk = audio[start:start + n_fft]
        frame[:len(chunk)] = chunk
        frame *= window                              # float32 in-place product, dataset.py:131
        cols.append(np.fft.rfft(frame))
    return np.stack(cols, axis=-1)


def window_sum_squares(L: int, n_fft: int, hop: int) -> np.ndarray:
    """sum_i w^2[n - i*hop] for n < L: the overlap-add denominator.  Where it is tiny (the first samples of a Hann
    analysis) the inverse amplifies the fp32 rounding of the spectrum by 1/wss."""
    w2 = np.hanning(n_fft) ** 2
    T = num_frames(L, hop)
    wss = np.zeros((T - 1) * hop + n_fft)
    for i in range(T):
        wss[i * hop:i * hop + n_fft] += w2
    return wss[:L]


def irfft_direct(half: np.ndarray, n_fft: int) -> np.ndarray:
    """numpy irfft semantics by the definition (imaginary parts of DC / Nyquist ignored); O(n_fft^2), small sizes."""
    half = np.asarray(half, dtype=np.complex128)
    k = np.arange(1, n_fft // 2)
    n = np.arange(n_fft)
    ph = np.exp(2j * np.pi * np.outer(n, k) / n_fft)
    y = half[0].real + half[n_fft // 2].real * np.cos(np.pi * n) + 2.0 * (ph @ half[1:n_fft // 2]).real
    return y / n_fft


def istft_masked(spec: np.ndarray, masks: np.ndarray | None, L: int, n_fft: int = 512, hop: int = 128,
                 direct: bool = False) -> np.ndarray:
    """spec (F, T) complex, masks (S, F, T) or None -> (S, L) float64."""
    spec = np.asarray(spec, dtype=np.complex128)
    F, T = spec.shape
    m = np.ones((1, F, T)) if masks is None else np.asarray(masks, dtype=np.float64)
    window = np.hanning(n_fft)
    span = (T - 1) * hop + n_fft
    assert L <= span
    out = np.zeros((m.shape[0], L))
    for s in range(m.shape[0]):
        acc = np.zeros(span)
        wss = np.zeros(span)
        for i in range(T):
            col = m[s, :, i] * spec[:, i]
            frame = irfft_direct(col, n_fft) if direct else np.fft.irfft(col, n_fft)
            acc[i * hop:i * hop + n_fft] += window * frame
            wss[i * hop:i * hop + n_fft] += window * window
        ok = wss[:L] > 1e-11
        out[s, ok] = acc[:L][ok] / wss[:L][ok]
    return out
