"""Golden vectors for the rows either side of the forward path, from the REAL reference (build container only).

    python tests/golden/make_synth_golden.py      # writes tests/golden/synth_items.npz, synth_snr.json

Items come from the reference's SyntheticAVDataset (dataset.py), SNR scalars from the reference's demo.py / losses.py
functions.  The oracle restatement (oracle/synth_oracle.py) is checked against them here; the fixture is only written
when it agrees.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth_oracle as so   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ds_mod = load("/root/reference/src/av_separation/dataset.py", "ref_dataset")
    loss_mod = load("/root/reference/src/av_separation/losses.py", "ref_losses")
    # demo.py imports the package at module scope; its three SNR helpers are plain numpy -- exec just those.
    src = open("/root/reference/demo.py").read()
    a = src.index("def snr_db")
    b = src.index("def evaluate_separation")
    c = src.index("def _permutation_snr")
    d = src.index("def quick_train")
    ns = {"np": np, "math": __import__("math")}
    exec(src[a:b] + src[c:d], ns)          # reference functions, executed unmodified
    snr_db, perm_snr = ns["snr_db"], ns["_permutation_snr"]

    out = {}
    report = {}
    # default dataset (demo default: 25 frames/speaker -> N=50) and a second geometry (16 kHz, 2 s, 3 speakers)
    cases = {
        "default": (dict(), so.SynthConfig(), [0, 1, 7, 123]),
        "wide": (dict(sample_rate=16000, duration=0.5, n_fft=256, hop_length=64, num_frames=10, frame_h=16, frame_w=24,
                      speaker_freqs=(200.0, 330.0, 512.0)),
                 so.SynthConfig(sample_rate=16000, duration=0.5, n_fft=256, hop_length=64, num_frames=10, frame_h=16,
                                frame_w=24, speaker_freqs=(200.0, 330.0, 512.0)), [3, 4]),
    }
    for cname, (kw, cfg, idxs) in cases.items():
        ds = ds_mod.SyntheticAVDataset(num_samples=1000, **kw)
        for idx in idxs:
            item = ds[idx]
            mine = so.synth_item(cfg, idx)
            for k in ("mixed_spec", "lip_frames", "clean_specs"):
                ref = item[k].numpy()
                err = float(np.abs(ref - mine[k]).max())
                report[f"{cname}/{idx}/{k}"] = err
                assert ref.shape == mine[k].shape and err <= 1e-4 * max(1.0, float(np.abs(ref).max())), (cname, idx, k, err)
                out[f"{cname}_{idx}_{k}"] = ref
    np.savez_compressed(os.path.join(HERE, "synth_items.npz"), **out)

    # SNR scalars on dataset items with a perturbed "separated" (deterministic)
    snr = {}
    ds = ds_mod.SyntheticAVDataset(num_samples=10)
    for idx in (0, 5):
        item = ds[idx]
        tg = item["clean_specs"].numpy()
        mixed = item["mixed_spec"].numpy()
        rng = np.random.default_rng(100 + idx)
        sep = (tg[::-1] * rng.uniform(0.7, 1.1, tg.shape) + rng.normal(0, 0.3, tg.shape)).astype(np.float32)   # swapped speakers
        e = {
            "input_snr": [snr_db(tg[s], mixed - tg[s]) for s in range(tg.shape[0])],
            "perm_snr": perm_snr(sep, tg),
            "si_snr_mean": float(loss_mod.si_snr(torch.from_numpy(sep.copy()), torch.from_numpy(tg.copy()))),
        }
        mine_in = so.input_snrs(mixed, tg)
        assert np.allclose(mine_in, e["input_snr"], atol=1e-9)
        assert abs(so.permutation_snr(sep, tg) - e["perm_snr"]) < 1e-9
        assert abs(float(so.si_snr_rows(sep, tg).mean()) - e["si_snr_mean"]) < 1e-4
        snr[str(idx)] = e
    with open(os.path.join(HERE, "synth_snr.json"), "w") as f:
        json.dump({"snr": snr, "oracle_vs_reference_max_abs_err": report}, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
