"""Deterministic synthetic weights and inputs for the AV-separation forward path.

TEST INFRASTRUCTURE (part of ``oracle/``): only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py`` may import this module.

Why this exists: the reference's encoder stacks are deep copies of one layer
(torch/nn/modules/transformer.py:363), its BatchNorm running stats are 0/1 and its
attention biases are 0 at construction (hard part 9), so parity on
a fresh random init cannot see layer-index, BN-fold or bias bugs.  ``make_state_dict``
therefore produces a full ``state_dict`` (key set and shapes of SURVEY.md Appendix A,
i.e. /root/reference/src/av_separation/model.py:37-52,81-101,143,155-163,194-199,290-297)
in which every layer, bias and statistic is distinct, from a numpy RNG only, so that
the same tensors can be rebuilt on a box that has neither the reference nor fixtures.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np


@dataclass(frozen=True)
class ModelConfig:
    """Constructor arguments of the reference model (model.py:240-249)."""
    freq_bins: int = 257
    d_model: int = 256
    nhead: int = 4
    num_encoder_layers: int = 2
    num_fusion_layers: int = 2
    num_speakers: int = 2

    def as_dict(self):
        return asdict(self)


# BASELINE.json configs: C1/C2 default model, C4 scaled model, and the
# tiny shapes the reference's own tests use (tests/test_model.py:29-36).
CONFIGS = {
    "default": ModelConfig(257, 256, 4, 2, 2, 2),
    "scaled": ModelConfig(257, 512, 8, 6, 6, 3),
    "tiny": ModelConfig(65, 64, 4, 1, 1, 2),
    "tiny2": ModelConfig(65, 64, 4, 2, 2, 2),
}

PE_MAX_LEN = 5000  # model.py:286


def positional_table(d_model: int, max_len: int = PE_MAX_LEN) -> np.ndarray:
    """Sinusoidal table of model.py:290-297, fp32 arithmetic, shape (1, max_len, d)."""
    pe = np.zeros((max_len, d_model), dtype=np.float32)
    position = np.arange(0, max_len, dtype=np.float32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2, dtype=np.float32)
                      * np.float32(-math.log(10000.0) / d_model)).astype(np.float32)
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe[None]


def state_dict_spec(cfg: ModelConfig):
    """Ordered (key, shape, kind, fan_in) list: the reference's state_dict contract."""
    d, F, S = cfg.d_model, cfg.freq_bins, cfg.num_speakers
    spec = []

    def lin(prefix, out_f, in_f, wname="weight", bname="bias"):
        spec.append((f"{prefix}.{wname}" if wname else prefix, (out_f, in_f), "w", in_f))
        spec.append((f"{prefix}.{bname}", (out_f,), "b", in_f))

    def ln(prefix):
        spec.append((f"{prefix}.weight", (d,), "gamma", 0))
        spec.append((f"{prefix}.bias", (d,), "beta", 0))

    def encoder_stack(prefix):
        for l in range(cfg.num_encoder_layers):
            p = f"{prefix}.transformer.layers.{l}"
            spec.append((f"{p}.self_attn.in_proj_weight", (3 * d, d), "w", d))
            spec.append((f"{p}.self_attn.in_proj_bias", (3 * d,), "b", d))
            lin(f"{p}.self_attn.out_proj", d, d)
            lin(f"{p}.linear1", 4 * d, d)
            lin(f"{p}.linear2", d, 4 * d)
            ln(f"{p}.norm1")
            ln(f"{p}.norm2")

    # AudioEncoder (model.py:37-52)
    spec.append(("audio_encoder.input_proj.0.weight", (d, F, 3), "w", 3 * F))
    spec.append(("audio_encoder.input_proj.0.bias", (d,), "b", 3 * F))
    spec.append(("audio_encoder.input_proj.2.weight", (d, d, 3), "w", 3 * d))
    spec.append(("audio_encoder.input_proj.2.bias", (d,), "b", 3 * d))
    spec.append(("audio_encoder.pos_enc.pe", (1, PE_MAX_LEN, d), "pe", 0))
    encoder_stack("audio_encoder")
    # VisualEncoder (model.py:81-101)
    cin = 1
    for idx, cout in ((0, 32), (3, 64), (6, 128)):
        spec.append((f"visual_encoder.conv.{idx}.weight", (cout, cin, 3, 3), "w", 9 * cin))
        spec.append((f"visual_encoder.conv.{idx}.bias", (cout,), "b", 9 * cin))
        bn = f"visual_encoder.conv.{idx + 1}"
        spec.append((f"{bn}.weight", (cout,), "gamma", 0))
        spec.append((f"{bn}.bias", (cout,), "beta", 0))
        spec.append((f"{bn}.running_mean", (cout,), "mean", 0))
        spec.append((f"{bn}.running_var", (cout,), "var", 0))
        spec.append((f"{bn}.num_batches_tracked", (), "count", 0))
        cin = cout
    lin("visual_encoder.frame_proj", d, 128)
    spec.append(("visual_encoder.pos_enc.pe", (1, PE_MAX_LEN, d), "pe", 0))
    encoder_stack("visual_encoder")
    # CrossModalFusion (model.py:140-163)
    for l in range(cfg.num_fusion_layers):
        p = f"fusion.layers.{l}"
        spec.append((f"{p}.cross_attn.in_proj_weight", (3 * d, d), "w", d))
        spec.append((f"{p}.cross_attn.in_proj_bias", (3 * d,), "b", d))
        lin(f"{p}.cross_attn.out_proj", d, d)
        lin(f"{p}.ff.0", 4 * d, d)
        lin(f"{p}.ff.3", d, 4 * d)
        ln(f"{p}.norm1")
        ln(f"{p}.norm2")
    ln("fusion.norm")
    # SeparationDecoder (model.py:194-199)
    lin("decoder.decoder.0", 2 * d, d)
    lin("decoder.decoder.3", S * F, 2 * d)
    return spec


def make_state_dict(cfg: ModelConfig, seed: int = 0, gain: float = 1.0) -> dict:
    """Full state_dict as numpy arrays (fp32; ``num_batches_tracked`` int64).

    ``gain`` scales the weight matrices (1.0 = torch's default 1/sqrt(fan_in) range);
    a gain above 1 makes activations, attention logits and masks less degenerate.
    """
    rng = np.random.default_rng(seed)
    pe = positional_table(cfg.d_model)
    out = {}
    for key, shape, kind, fan_in in state_dict_spec(cfg):
        if kind == "w":
            a = gain / math.sqrt(fan_in)
            v = rng.uniform(-a, a, size=shape)
        elif kind == "b":
            a = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-a, a, size=shape)
        elif kind == "gamma":
            v = 1.0 + 0.1 * rng.standard_normal(size=shape)
        elif kind == "beta":
            v = 0.1 * rng.standard_normal(size=shape)
        elif kind == "mean":
            v = 0.1 * rng.standard_normal(size=shape)
        elif kind == "var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif kind == "pe":
            out[key] = pe.copy()
            continue
        elif kind == "count":
            out[key] = np.array(0, dtype=np.int64)
            continue
        else:  # pragma: no cover
            raise AssertionError(kind)
        out[key] = np.asarray(v, dtype=np.float32)
    return out


def num_parameters(cfg: ModelConfig) -> int:
    """Trainable parameter count (excludes buffers); README.md:60 quotes 1,612,738 at d=128."""
    n = 0
    for _, shape, kind, _ in state_dict_spec(cfg):
        if kind in ("w", "b", "gamma", "beta"):
            n += int(np.prod(shape)) if shape else 1
    return n


def make_inputs(cfg: ModelConfig, B: int, T: int, N: int, Hh: int = 32, Ww: int = 32,
                seed: int = 0, kind: str = "dataset"):
    """Seeded inputs with the reference's shapes (model.py:268; dataset.py:116-120).

    kind="dataset": SyntheticAVDataset-shaped values: ``mixed_spec`` >= 0 and
      heavy-tailed -- two spectral lines per utterance that peak near 100 over a small
      floor; ``lip_frames`` in [0,1], zero outside the centre half of the frame.
    kind="randn": unit-scale Gaussian inputs, like the reference tests (tests/test_model.py:44-51).
    """
    rng = np.random.default_rng(seed + 7919)
    F = cfg.freq_bins
    if kind == "randn":
        mixed = rng.standard_normal((B, F, T)).astype(np.float32)
        frames = rng.standard_normal((B, N, Hh, Ww)).astype(np.float32)
        return mixed, frames
    if kind != "dataset":
        raise ValueError(kind)
    f = np.arange(F, dtype=np.float32)[None, :, None]
    mixed = np.zeros((B, F, T), dtype=np.float32)
    for _ in range(2):
        centre = rng.uniform(0.04, 0.3, size=(B, 1, 1)) * F
        amp = rng.uniform(30.0, 110.0, size=(B, 1, 1))
        width = rng.uniform(0.8, 1.6, size=(B, 1, 1))
        env = 1.0 + 0.05 * rng.standard_normal((B, 1, T))
        mixed += (amp * np.exp(-0.5 * ((f - centre) / width) ** 2) * env).astype(np.float32)
    mixed += np.abs(rng.standard_normal((B, F, T))).astype(np.float32) * 1e-3
    frames = np.zeros((B, N, Hh, Ww), dtype=np.float32)
    h0, h1, w0, w1 = Hh // 4, 3 * Hh // 4, Ww // 4, 3 * Ww // 4
    bright = rng.uniform(0.0, 1.0, size=(B, N, 1, 1))
    noise = 0.05 * rng.standard_normal((B, N, h1 - h0, w1 - w0))
    frames[:, :, h0:h1, w0:w1] = np.clip(bright + noise, 0.0, 1.0)
    return mixed.astype(np.float32), frames.astype(np.float32)
