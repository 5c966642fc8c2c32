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


@pytest.mark.parametrize("name", CASES)
def test_inverse_is_scipy_signal_istft(name):
    """Pin of the inverse on a named third-party algorithm (the reference has none, README.md:140):
    scipy.signal.istft with the reference's analysis window (np.hanning(n_fft), dataset.py:124), hop and no boundary
    extension is the same weighted overlap-add -- irfft of every column, times the window, summed at i*hop, divided by
    the summed squared window where that is > 1e-10.  scipy's transform pair carries a 1/sum(window) scaling
    (scaling='spectrum'), which is undone on the input."""
    from scipy import signal
    z = np.load(GOLD)
    n_fft, hop = (int(v) for v in z[name + "_geom"])
    x = z[name + "_wave"]
    spec = wo.stft_complex(x, n_fft, hop)
    rng = np.random.default_rng(11)
    masks = rng.uniform(0, 1, (2,) + spec.shape)
    win = np.hanning(n_fft)
    L = len(x)
    ours = wo.istft_masked(spec, masks, L, n_fft, hop)
    import warnings
    for s in range(2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")              # NOLA warning: np.hanning's end points are 0 (sample 0 has no mass)
            _, ref = signal.istft((masks[s] * spec.astype(np.complex128)) / win.sum(), window=win, nperseg=n_fft,
                                  noverlap=n_fft - hop, nfft=n_fft, input_onesided=True, boundary=False)
        ref = ref[:L]
        wss = wo.window_sum_squares(L, n_fft, hop)
        live = wss > 1e-10                               # scipy leaves samples with no window mass un-normalised (= 0 here)
        np.testing.assert_allclose(ours[s][live], ref[live], rtol=0, atol=1e-9 * max(1.0, np.abs(ref[live]).max()))
        assert np.all(ours[s][~live] == 0.0)
