"""Generate golden input/output vectors from the REAL reference (build container only).

    python tests/golden/make_golden.py [case ...]  # writes tests/golden/*.npz (all cases, or the named ones)

The reference is a Python package at /root/reference/src (read-only, absent on the GPU box).
It is imported under an alias, loaded with ``oracle.weights.make_state_dict`` (numpy RNG,
reproducible anywhere), run in eval mode under no_grad on CPU fp32, and the outputs plus
per-stage probes (forward hooks) are stored.  Weights and inputs are NOT stored: tests
regenerate them from the recorded seeds.  Both oracle restatements are checked against the
reference here, so a fixture is only written when oracle == reference to ~1e-5.
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

from oracle import avsep_oracle as onp            # noqa: E402
from oracle import avsep_oracle_torch as otorch   # noqa: E402
from oracle.weights import CONFIGS, make_inputs, make_state_dict, num_parameters, ModelConfig  # noqa: E402

REF_SRC = "/root/reference/src/av_separation"

# name -> (config name, B, T, N, Hh, Ww, input kind, weight seed, input seed, gain, store stride (f,t))
CASES = {
    # the reference tests' own shapes (tests/test_model.py:29-36): F=65, hd=16, 16x16 frames, N=10
    "tiny_up":    ("tiny", 2, 32, 10, 16, 16, "randn", 1, 1, 2.0, (1, 1)),
    # T < N (down-sampling interpolation, tests/test_model.py:105-114) and odd frame sizes (ceil/2 thrice)
    "tiny_down":  ("tiny2", 3, 20, 30, 15, 18, "randn", 2, 2, 2.0, (1, 1)),
    # C1: demo default (BASELINE.json configs[0]) on SyntheticAVDataset-shaped inputs
    "c1_dataset": ("default", 2, 63, 50, 32, 32, "dataset", 3, 3, 2.0, (1, 1)),
    "c1_randn":   ("default", 1, 63, 50, 32, 32, "randn", 4, 4, 2.0, (1, 1)),
    # C3: long-form 10 s @ 16 kHz (T=1251, N=500); outputs stored subsampled
    "c3_long":    ("default", 1, 1251, 500, 32, 32, "dataset", 5, 5, 2.0, (4, 5)),
    # C4: scaled model d=512, H=8, 6+6 layers, 3 speakers
    "c4_scaled":  ("scaled", 1, 63, 50, 32, 32, "dataset", 6, 6, 2.0, (1, 1)),
    # C1 at a batch that takes the large-batch schedules of the B=256 benchmark (>= 2048 rows: fused transformer
    # layers / fused FFN, multi-group CNN, multi-tile GEMMs); outputs stored subsampled
    "c1_batch64": ("default", 64, 63, 50, 32, 32, "dataset", 7, 7, 2.0, (4, 3)),
}

STAGES = ("audio_embed", "audio_enc", "visual_pool", "visual_embed", "visual_enc", "visual_interp", "fused")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_av_model", os.path.join(REF_SRC, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(ref, cfg: ModelConfig, P_np, mixed, frames):
    torch.manual_seed(0)
    model = ref.AVSeparationTransformer(**cfg.as_dict())
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in P_np.items()}
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert sum(p.numel() for p in model.parameters()) == num_parameters(cfg)
    model.eval()
    stages = {}

    def grab(name, post=None):
        def hook(_m, _i, out):
            stages[name] = (post(out) if post else out).detach().numpy().copy()
        return hook

    B, N = frames.shape[:2]
    hs = [
        model.audio_encoder.pos_enc.register_forward_hook(grab("audio_embed")),
        model.audio_encoder.register_forward_hook(grab("audio_enc")),
        model.visual_encoder.conv.register_forward_hook(grab("visual_pool", lambda o: o.reshape(B, N, 128))),
        model.visual_encoder.pos_enc.register_forward_hook(grab("visual_embed")),
        model.visual_encoder.transformer.register_forward_hook(grab("visual_enc")),
        model.visual_encoder.register_forward_hook(grab("visual_interp")),
        model.fusion.register_forward_hook(grab("fused")),
    ]
    with torch.no_grad():
        sep, masks = model(torch.from_numpy(mixed), torch.from_numpy(frames))
    for h in hs:
        h.remove()
    return sep.numpy().copy(), masks.numpy().copy(), stages


def main():
    ref = load_reference()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    summary = {}
    # README.md:60 known answer: the d_model=128 demo model has 1,612,738 parameters
    assert num_parameters(ModelConfig(257, 128, 4, 2, 2, 2)) == 1612738
    only = set(sys.argv[1:])
    spath = os.path.join(out_dir, "SUMMARY.json")
    if only and os.path.exists(spath):
        with open(spath) as f:
            summary = json.load(f)
    for name, (cname, B, T, N, Hh, Ww, kind, wseed, iseed, gain, (sf, st)) in CASES.items():
        if only and name not in only:
            continue
        cfg = CONFIGS[cname]
        P = make_state_dict(cfg, seed=wseed, gain=gain)
        mixed, frames = make_inputs(cfg, B, T, N, Hh, Ww, seed=iseed, kind=kind)
        sep, masks, stages = run_reference(ref, cfg, P, mixed, frames)
        # oracle (torch-functional port) vs reference
        Pt = otorch.to_torch(P)
        sep_t, masks_t, st_t = otorch.forward(Pt, cfg, torch.from_numpy(mixed), torch.from_numpy(frames), True)
        e_t = float(np.abs(masks_t.numpy() - masks).max())
        # oracle (numpy restatement) vs reference
        sep_n, masks_n, st_n = onp.forward(P, cfg, mixed, frames, return_stages=True)
        e_n = float(np.abs(masks_n - masks).max())
        e_ns = float(np.abs(sep_n - sep).max() / max(1.0, np.abs(mixed).max()))
        stage_err = {k: float(np.abs(st_n[k] - stages[k]).max()) for k in STAGES}
        assert e_t < 2e-5 and e_n < 2e-5 and e_ns < 2e-5, (name, e_t, e_n, e_ns)
        store = {
            "separated": sep[:, :, ::sf, ::st].astype(np.float32),
            "masks": masks[:, :, ::sf, ::st].astype(np.float32),
            "masks_sum": np.float64(masks.astype(np.float64).sum()),
            "separated_sum": np.float64(sep.astype(np.float64).sum()),
        }
        for k in STAGES:
            v = stages[k]
            store["stage_" + k] = v if v.size <= 70000 else v[:, ::max(1, v.shape[1] // 16)][:, :, ::4].copy()
        meta = dict(config=cname, B=B, T=T, N=N, Hh=Hh, Ww=Ww, kind=kind, weight_seed=wseed,
                    input_seed=iseed, gain=gain, stride_f=sf, stride_t=st,
                    torch=torch.__version__, oracle_numpy_err_masks=e_n, oracle_torch_err_masks=e_t)
        store["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **store)
        summary[name] = dict(masks_min=float(masks.min()), masks_max=float(masks.max()),
                             masks_std=float(masks.std()), mixed_max=float(np.abs(mixed).max()),
                             err_numpy=e_n, err_torch=e_t, err_sep_rel=e_ns, stage_err=stage_err)
        print(name, json.dumps(summary[name]))
    with open(os.path.join(out_dir, "SUMMARY.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
