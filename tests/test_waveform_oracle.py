"""The waveform oracle (oracle/waveform_oracle.py) against the reference's own analysis (tests/golden/waveform_stft.npz,
made from the real SyntheticAVDataset._stft) and against the identities that anchor the inverse, which the reference
does not have (README.md:140)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import waveform_oracle as wo   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "waveform_stft.npz")
CASES = ("default", "short", "ragged")


@pytest.mark.parametrize("name", CASES)
def test_stft_magnitude_is_the_reference_stft(name):
    z = np.load(GOLD)
    n_fft, hop = (int(v) for v in z[name + "_geom"])
    spec = wo.stft_complex(z[name + "_wave"], n_fft, hop)
    assert spec.shape == z[name + "_mag"].shape
    np.testing.assert_array_equal(np.abs(spec).astype(np.float32), z[name + "_mag"])   # same operations: bit-exact


@pytest.mark.parametrize("name", CASES)
def test_round_trip_recovers_the_waveform(name):
    z = np.load(GOLD)
    n_fft, hop = (int(v) for v in z[name + "_geom"])
    x = z[name + "_wave"]
    y = wo.istft_masked(wo.stft_complex(x, n_fft, hop), None, len(x), n_fft, hop)[0]
    assert y[0] == 0.0                                    # np.hanning(M)[0] == 0: no frame reaches sample 0
    wss = wo.window_sum_squares(len(x), n_fft, hop)
    solid = wss > 1e-3                                    # elsewhere 1/wss amplifies the complex64 rounding of the spectrum
    np.testing.assert_allclose(y[solid], x[solid], rtol=0, atol=2e-6)
    np.testing.assert_allclose(y[1:], x[1:], rtol=0, atol=1e-2)


def test_masks_summing_to_one_split_the_mixture():
    z = np.load(GOLD)
    n_fft, hop = (int(v) for v in z["short_geom"])
    x = z["short_wave"]
    spec = wo.stft_complex(x, n_fft, hop)
    rng = np.random.default_rng(5)
    m0 = rng.uniform(0, 1, (1,) + spec.shape)
    y = wo.istft_masked(spec, np.concatenate([m0, 1 - m0]), len(x), n_fft, hop)
    solid = wo.window_sum_squares(len(x), n_fft, hop) > 1e-3
    np.testing.assert_allclose(y.sum(0)[solid], x[solid], rtol=0, atol=2e-6)


def test_fast_inverse_agrees_with_the_direct_dft():
    rng = np.random.default_rng(6)
    n_fft, hop, T = 32, 8, 9
    spec = rng.standard_normal((n_fft // 2 + 1, T)) + 1j * rng.standard_normal((n_fft // 2 + 1, T))   # non-zero imag at DC / Nyquist
    m = rng.uniform(0, 1, (2, n_fft // 2 + 1, T))
    L = (T - 1) * hop + n_fft
    a = wo.istft_masked(spec, m, L, n_fft, hop)
    b = wo.istft_masked(spec, m, L, n_fft, hop, direct=True)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)
    assert a[:, 0].tolist() == [0.0, 0.0] and a[:, L - 1].tolist() == [0.0, 0.0]   # both window end points are 0
