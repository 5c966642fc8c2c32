"""CPU oracle (numpy, fp32) for ``AVSeparationTransformer.forward`` in eval mode.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product path
(``av-separation-transformer_b200/``) never does and has no CPU fallback.

This is a restatement, not a copy: the reference (/root/reference/src/av_separation/model.py)
only *wires* stock ``torch.nn`` modules, so the arithmetic restated here is PyTorch's
(un-vendored third-party dependency, ``torch>=2.0.0`` in requirements.txt:2; the container has
torch 2.11.0+cu128).  Each function cites the reference line it follows and, where the
semantics live in torch, the torch source line.

PARITY PINNING.  The reference's own tests hold no golden vectors or known-answer tests
for this path (all 30 tests are shape/range/grad checks on unseeded
randn).  The oracle is therefore pinned against outputs of the reference itself, run in
the build container: ``tests/golden/make_golden.py`` imports the real reference from
/root/reference/src, loads ``oracle.weights.make_state_dict`` into it and stores
(separated, masks, per-stage probes) under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this module against those fixtures.
"""
from __future__ import annotations

import math

import numpy as np

try:  # exact erf for nn.GELU() (approximate='none')
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover - scipy is present in the image
    _erf = np.vectorize(math.erf, otypes=[np.float32])

F32 = np.float32
LN_EPS = F32(1e-5)   # nn.LayerNorm default, model.py:143,162-163; encoder layers torch/nn/modules/transformer.py:738
BN_EPS = F32(1e-5)   # nn.BatchNorm2d default, model.py:83,86,89


# ---------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------

def linear(x, w, b=None):
    """nn.Linear: y = x @ w.T + b (weights stored (out,in))."""
    y = x.astype(F32) @ w.astype(F32).T
    if b is not None:
        y = y + b.astype(F32)
    return y.astype(F32)


def layer_norm(x, gamma, beta, eps=LN_EPS):
    """nn.LayerNorm over the last axis, biased variance (model.py:143,149,162-163,168,172)."""
    x = x.astype(F32)
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    return (xc / np.sqrt(var + eps) * gamma + beta).astype(F32)


def gelu_erf(x):
    """nn.GELU() default = exact erf form (model.py:158,196)."""
    x = x.astype(F32)
    return (F32(0.5) * x * (F32(1.0) + _erf(x * F32(1.0 / math.sqrt(2.0))).astype(F32))).astype(F32)


def relu(x):
    return np.maximum(x, F32(0.0))


def sigmoid(x):
    x = x.astype(F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


def softmax_lastdim(s):
    s = s - s.max(axis=-1, keepdims=True)
    e = np.exp(s)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def conv1d_k3p1(x, w, b):
    """nn.Conv1d(k=3, padding=1): cross-correlation, zero pad (model.py:38,40,56).

    x (B,Cin,T), w (Cout,Cin,3), b (Cout) -> (B,Cout,T)
    """
    B, Cin, T = x.shape
    xp = np.zeros((B, Cin, T + 2), dtype=F32)
    xp[:, :, 1:T + 1] = x
    y = np.zeros((B, w.shape[0], T), dtype=F32)
    for tap in range(3):
        # out[b,o,t] += sum_c w[o,c,tap] * xp[b,c,t+tap]
        y += np.einsum("oc,bct->bot", w[:, :, tap].astype(F32), xp[:, :, tap:tap + T], optimize=True)
    return (y + b.astype(F32)[None, :, None]).astype(F32)


def conv2d_k3s2p1(x, w, b):
    """nn.Conv2d(k=3, stride=2, padding=1) (model.py:82,85,88). x (M,Cin,H,W) -> (M,Cout,ceil(H/2),ceil(W/2))."""
    M, Cin, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    xp = np.zeros((M, Cin, H + 2, W + 2), dtype=F32)
    xp[:, :, 1:H + 1, 1:W + 1] = x
    y = np.zeros((M, w.shape[0], Ho, Wo), dtype=F32)
    for ky in range(3):
        for kx in range(3):
            patch = xp[:, :, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2][:, :, :Ho, :Wo]
            y += np.einsum("oc,mchw->mohw", w[:, :, ky, kx].astype(F32), patch, optimize=True)
    return (y + b.astype(F32)[None, :, None, None]).astype(F32)


def batchnorm2d_eval(x, gamma, beta, mean, var, eps=BN_EPS):
    """BatchNorm2d in eval mode: running statistics (model.py:83,86,89; SURVEY Appendix B rule 5)."""
    scale = gamma.astype(F32) / np.sqrt(var.astype(F32) + eps)
    return ((x - mean.astype(F32)[None, :, None, None]) * scale[None, :, None, None]
            + beta.astype(F32)[None, :, None, None]).astype(F32)


def interp_linear_time(x, T_out):
    """F.interpolate(mode='linear', align_corners=False) along time (model.py:114-116).

    x (B,N,d) -> (B,T_out,d).  Index rule of ATen upsample_linear1d (SURVEY Appendix B rule 6):
    src = max(scale*(j+0.5)-0.5, 0) with scale=(float)N/T in fp32, i0=floor(src),
    i1=min(i0+1,N-1), lam=src-i0.
    """
    B, N, d = x.shape
    scale = F32(N) / F32(T_out)
    j = np.arange(T_out, dtype=F32)
    src = np.maximum(scale * (j + F32(0.5)) - F32(0.5), F32(0.0)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), N - 1)
    i1 = np.minimum(i0 + 1, N - 1)
    lam = (src - i0.astype(F32)).astype(F32)[None, :, None]
    return ((F32(1.0) - lam) * x[:, i0, :] + lam * x[:, i1, :]).astype(F32)


def interp_coefficients(N, T_out):
    """(i0, i1, lam) of the rule above; exported so tests can check the CUDA kernel's table."""
    scale = F32(N) / F32(T_out)
    j = np.arange(T_out, dtype=F32)
    src = np.maximum(scale * (j + F32(0.5)) - F32(0.5), F32(0.0)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), N - 1)
    i1 = np.minimum(i0 + 1, N - 1)
    return i0, i1, (src - i0.astype(F32)).astype(F32)


def multi_head_attention(q_in, kv_in, in_w, in_b, out_w, out_b, nhead):
    """nn.MultiheadAttention / encoder self-attention, eval, no mask.

    Packed in_proj rows are q|k|v (torch/nn/functional.py:5847-5865); q is scaled by
    1/sqrt(hd) before q.k^T (torch/nn/functional.py:6630-6665); heads are contiguous
    hd-wide column slices (SURVEY Appendix B rule 4).
    q_in (B,Lq,d), kv_in (B,Lk,d) -> (B,Lq,d)
    """
    B, Lq, d = q_in.shape
    Lk = kv_in.shape[1]
    hd = d // nhead
    q = linear(q_in, in_w[:d], in_b[:d])
    k = linear(kv_in, in_w[d:2 * d], in_b[d:2 * d])
    v = linear(kv_in, in_w[2 * d:], in_b[2 * d:])
    q = q.reshape(B, Lq, nhead, hd).transpose(0, 2, 1, 3) * F32(1.0 / math.sqrt(hd))
    k = k.reshape(B, Lk, nhead, hd).transpose(0, 2, 1, 3)
    v = v.reshape(B, Lk, nhead, hd).transpose(0, 2, 1, 3)
    s = np.einsum("bhqe,bhke->bhqk", q, k, optimize=True).astype(F32)
    p = softmax_lastdim(s)
    o = np.einsum("bhqk,bhke->bhqe", p, v, optimize=True).astype(F32)
    o = o.transpose(0, 2, 1, 3).reshape(B, Lq, d)
    return linear(o, out_w, out_b)


# ---------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------

def encoder_layer(x, P, p, nhead):
    """nn.TransformerEncoderLayer(norm_first=True, activation=relu, ff=4d), eval
    (model.py:48-52,97-101; math torch/nn/modules/transformer.py:946-950)."""
    h = layer_norm(x, P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"])
    x = x + multi_head_attention(h, h, P[f"{p}.self_attn.in_proj_weight"], P[f"{p}.self_attn.in_proj_bias"],
                                 P[f"{p}.self_attn.out_proj.weight"], P[f"{p}.self_attn.out_proj.bias"], nhead)
    h = layer_norm(x, P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"])
    h = relu(linear(h, P[f"{p}.linear1.weight"], P[f"{p}.linear1.bias"]))
    x = x + linear(h, P[f"{p}.linear2.weight"], P[f"{p}.linear2.bias"])
    return x.astype(F32)


def audio_encoder(P, cfg, mixed, stages=None):
    """AudioEncoder.forward (model.py:54-60)."""
    T = mixed.shape[-1]
    h = relu(conv1d_k3p1(mixed, P["audio_encoder.input_proj.0.weight"], P["audio_encoder.input_proj.0.bias"]))
    h = relu(conv1d_k3p1(h, P["audio_encoder.input_proj.2.weight"], P["audio_encoder.input_proj.2.bias"]))
    x = h.transpose(0, 2, 1) + P["audio_encoder.pos_enc.pe"][:, :T]     # model.py:57-58, 299-301
    if stages is not None:
        stages["audio_embed"] = x.copy()
    for l in range(cfg.num_encoder_layers):                             # model.py:59, no final norm
        x = encoder_layer(x, P, f"audio_encoder.transformer.layers.{l}", cfg.nhead)
    if stages is not None:
        stages["audio_enc"] = x.copy()
    return x


def visual_cnn(P, frames):
    """VisualEncoder.conv on (B*N,1,H,W) -> (B*N,128) (model.py:81-92,106-107)."""
    B, N, H, W = frames.shape
    x = frames.reshape(B * N, 1, H, W).astype(F32)
    for idx in (0, 3, 6):
        x = conv2d_k3s2p1(x, P[f"visual_encoder.conv.{idx}.weight"], P[f"visual_encoder.conv.{idx}.bias"])
        bn = f"visual_encoder.conv.{idx + 1}"
        x = relu(batchnorm2d_eval(x, P[f"{bn}.weight"], P[f"{bn}.bias"],
                                  P[f"{bn}.running_mean"], P[f"{bn}.running_var"]))
    return x.mean(axis=(2, 3), dtype=F32)                               # AdaptiveAvgPool2d((1,1))


def visual_encoder(P, cfg, frames, target_len, stages=None):
    """VisualEncoder.forward (model.py:103-117)."""
    B, N = frames.shape[:2]
    feat = visual_cnn(P, frames)
    if stages is not None:
        stages["visual_pool"] = feat.reshape(B, N, 128).copy()
    x = linear(feat, P["visual_encoder.frame_proj.weight"], P["visual_encoder.frame_proj.bias"])
    x = x.reshape(B, N, cfg.d_model) + P["visual_encoder.pos_enc.pe"][:, :N]
    if stages is not None:
        stages["visual_embed"] = x.copy()
    for l in range(cfg.num_encoder_layers):
        x = encoder_layer(x, P, f"visual_encoder.transformer.layers.{l}", cfg.nhead)
    if stages is not None:
        stages["visual_enc"] = x.copy()
    return interp_linear_time(x, target_len)


def cross_attention_layer(audio, visual, P, p, nhead):
    """CrossAttentionLayer.forward (model.py:166-173): q from LN1(audio), k/v from the
    un-normalised visual embedding; FFN uses exact GELU."""
    h = layer_norm(audio, P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"])
    audio = audio + multi_head_attention(h, visual, P[f"{p}.cross_attn.in_proj_weight"],
                                         P[f"{p}.cross_attn.in_proj_bias"],
                                         P[f"{p}.cross_attn.out_proj.weight"],
                                         P[f"{p}.cross_attn.out_proj.bias"], nhead)
    h = layer_norm(audio, P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"])
    h = gelu_erf(linear(h, P[f"{p}.ff.0.weight"], P[f"{p}.ff.0.bias"]))
    return (audio + linear(h, P[f"{p}.ff.3.weight"], P[f"{p}.ff.3.bias"])).astype(F32)


def fusion(P, cfg, audio, visual):
    """CrossModalFusion.forward (model.py:145-149)."""
    h = audio
    for l in range(cfg.num_fusion_layers):
        h = cross_attention_layer(h, visual, P, f"fusion.layers.{l}", cfg.nhead)
    return layer_norm(h, P["fusion.norm.weight"], P["fusion.norm.bias"])


def decoder_masks(P, cfg, fused):
    """SeparationDecoder.forward (model.py:201-208): column s*F+f -> masks[b,s,f,t]."""
    B, T, _ = fused.shape
    h = gelu_erf(linear(fused, P["decoder.decoder.0.weight"], P["decoder.decoder.0.bias"]))
    logits = linear(h, P["decoder.decoder.3.weight"], P["decoder.decoder.3.bias"])
    logits = logits.reshape(B, T, cfg.num_speakers, cfg.freq_bins).transpose(0, 2, 3, 1)
    return sigmoid(logits)


def forward(P, cfg, mixed_spec, lip_frames, return_stages=False):
    """AVSeparationTransformer.forward (model.py:268-276) -> (separated, masks[, stages])."""
    mixed_spec = np.asarray(mixed_spec, dtype=F32)
    lip_frames = np.asarray(lip_frames, dtype=F32)
    stages = {} if return_stages else None
    T = mixed_spec.shape[-1]
    a = audio_encoder(P, cfg, mixed_spec, stages)
    v = visual_encoder(P, cfg, lip_frames, T, stages)
    f = fusion(P, cfg, a, v)
    masks = decoder_masks(P, cfg, f)
    separated = (masks * mixed_spec[:, None]).astype(F32)               # model.py:220
    if return_stages:
        stages["visual_interp"] = v
        stages["fused"] = f
        return separated, masks, stages
    return separated, masks
