"""Shared test helpers: golden fixtures, model construction from the deterministic weight generator."""
import json
import os

import numpy as np
import torch

from oracle.weights import CONFIGS, make_inputs, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("tiny_up", "tiny_down", "c1_dataset", "c1_randn", "c3_long", "c4_scaled", "c1_batch64")
STAGES = ("audio_embed", "audio_enc", "visual_pool", "visual_embed", "visual_enc", "fused")

# Stated tolerances (BASELINE.json north_star): 1e-2 abs for the bf16 path, 1e-3 for fp32/TF32.
# `separated` = masks * mixed_spec, so on dataset-scale inputs (|mixed| up to ~120) its bound is the mask
# bound times max(1, |mixed|).
TOL = {"bf16": 1e-2, "tf32": 1e-3}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def case_tensors(meta):
    cfg = CONFIGS[meta["config"]]
    P = make_state_dict(cfg, seed=meta["weight_seed"], gain=meta["gain"])
    mixed, frames = make_inputs(cfg, meta["B"], meta["T"], meta["N"], meta["Hh"], meta["Ww"],
                                seed=meta["input_seed"], kind=meta["kind"])
    return cfg, P, mixed, frames


def subsample_stage(v):
    """Same rule as tests/golden/make_golden.py."""
    return v if v.size <= 70000 else v[:, ::max(1, v.shape[1] // 16)][:, :, ::4].copy()


def build_model(cfg, P, precision="bf16", device="cuda"):
    from avsep_b200 import AVSeparationTransformer
    m = AVSeparationTransformer(**cfg.as_dict(), precision=precision)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def err_report(sep, masks, sep_ref, masks_ref, mixed):
    scale = np.maximum(1.0, np.abs(mixed))[:, None]
    return dict(masks=float(np.abs(masks - masks_ref).max()),
                separated_abs=float(np.abs(sep - sep_ref).max()),
                separated_scaled=float((np.abs(sep - sep_ref) / scale).max()))
