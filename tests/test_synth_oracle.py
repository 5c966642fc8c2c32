"""CPU: the synthesis / SNR oracle against the fixtures produced by the real reference, and the host-side draws."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from oracle import synth_oracle as so   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WIDE = so.SynthConfig(sample_rate=16000, duration=0.5, n_fft=256, hop_length=64, num_frames=10, frame_h=16, frame_w=24,
                      speaker_freqs=(200.0, 330.0, 512.0))
CASES = [("default", so.SynthConfig(), (0, 1, 7, 123)), ("wide", WIDE, (3, 4))]


def test_oracle_items_match_reference_dataset():
    z = np.load(os.path.join(GOLD, "synth_items.npz"))
    for cname, cfg, idxs in CASES:
        for idx in idxs:
            item = so.synth_item(cfg, idx)
            for k in ("mixed_spec", "lip_frames", "clean_specs"):
                ref = z[f"{cname}_{idx}_{k}"]
                assert item[k].shape == ref.shape and item[k].dtype == np.float32
                assert np.abs(item[k] - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (cname, idx, k)


def test_oracle_snr_matches_reference_functions():
    with open(os.path.join(GOLD, "synth_snr.json")) as f:
        gold = json.load(f)["snr"]
    cfg = so.SynthConfig()
    for idx_s, e in gold.items():
        idx = int(idx_s)
        item = so.synth_item(cfg, idx)
        tg, mixed = item["clean_specs"], item["mixed_spec"]
        rng = np.random.default_rng(100 + idx)
        sep = (tg[::-1] * rng.uniform(0.7, 1.1, tg.shape) + rng.normal(0, 0.3, tg.shape)).astype(np.float32)
        assert np.allclose(so.input_snrs(mixed, tg), e["input_snr"], atol=1e-6)
        assert abs(so.permutation_snr(sep, tg) - e["perm_snr"]) < 1e-6
        assert abs(float(so.si_snr_rows(sep, tg).mean()) - e["si_snr_mean"]) < 1e-3


def test_host_draws_follow_the_reference_rng_order():
    """avsep_b200.dataset.SyntheticAVDataset.draws (product host logic) == the oracle's draws, bit for bit."""
    from avsep_b200.dataset import SyntheticAVDataset
    for cfg, idxs in ((so.SynthConfig(), (0, 5, 999)), (WIDE, (3,))):
        ds = SyntheticAVDataset(sample_rate=cfg.sample_rate, duration=cfg.duration, n_fft=cfg.n_fft,
                                hop_length=cfg.hop_length, num_frames=cfg.num_frames, frame_h=cfg.frame_h,
                                frame_w=cfg.frame_w, speaker_freqs=cfg.speaker_freqs)
        assert (ds.freq_bins, ds.T, len(ds)) == (cfg.freq_bins, cfg.T, 1000)
        for idx in idxs:
            a, f, p, nz = ds.draws(idx)
            a2, f2, p2, nz2 = so.draw_item(cfg, idx)
            assert np.array_equal(a, a2) and np.array_equal(f, f2) and np.array_equal(p, p2) and np.array_equal(nz, nz2)
