"""Fast CPU port of the reference forward, on torch's CPU operators (MKL / oneDNN).

TEST INFRASTRUCTURE: used only as ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` arm
and by tests at sizes the numpy oracle (``avsep_oracle.py``) would take too long on.  The
product path never imports it.

The reference cannot travel to the GPU box (/root/reference does not exist there), and its
arithmetic lives in torch anyway, so this port calls the *same* torch.nn.functional CPU
operators the reference's modules dispatch to, with all host threads, on the same
state_dict:

* Conv1d/Conv2d/BatchNorm2d(eval)/AdaptiveAvgPool2d/Linear/LayerNorm/GELU/sigmoid/interpolate
  -> the F.* call the corresponding nn.Module makes (model.py:37-42,81-93,114-116,143,194-207);
* encoder layers -> pre-norm self-attention block + ReLU FFN
  (torch/nn/modules/transformer.py:946-950), attention through F.multi_head_attention_forward
  with need_weights=False as ``_sa_block`` does;
* fusion cross-attention -> F.multi_head_attention_forward with need_weights=True, i.e. the
  slow path that materialises and head-averages the (B,T,T) weights exactly as
  ``self.cross_attn(normed, visual, visual)`` does at model.py:169.

It is validated against the real reference in ``tests/golden/make_golden.py`` (run in the build
container) and against the committed fixtures in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _mha(q_in, kv_in, P, p, nhead, need_weights):
    # batch_first=True modules transpose to (L,B,d) around the functional (activation.py:1467-1519)
    q = q_in.transpose(0, 1)
    kv = kv_in.transpose(0, 1)
    out, _ = F.multi_head_attention_forward(
        q, kv, kv, q_in.shape[-1], nhead,
        P[f"{p}.in_proj_weight"], P[f"{p}.in_proj_bias"], None, None, False, 0.0,
        P[f"{p}.out_proj.weight"], P[f"{p}.out_proj.bias"],
        training=False, key_padding_mask=None, need_weights=need_weights, attn_mask=None,
        average_attn_weights=True)
    return out.transpose(0, 1)


def _encoder_layer(x, P, p, nhead):
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], 1e-5)
    x = x + _mha(h, h, P, f"{p}.self_attn", nhead, need_weights=False)
    h = F.layer_norm(x, (d,), P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], 1e-5)
    h = F.relu(F.linear(h, P[f"{p}.linear1.weight"], P[f"{p}.linear1.bias"]))
    return x + F.linear(h, P[f"{p}.linear2.weight"], P[f"{p}.linear2.bias"])


@torch.no_grad()
def forward(P, cfg, mixed_spec, lip_frames, return_stages=False):
    """P: dict of torch CPU tensors with the reference's state_dict keys."""
    d, H = cfg.d_model, cfg.nhead
    B, _, T = mixed_spec.shape
    stages = {}
    # AudioEncoder (model.py:54-60)
    h = F.relu(F.conv1d(mixed_spec, P["audio_encoder.input_proj.0.weight"],
                        P["audio_encoder.input_proj.0.bias"], padding=1))
    h = F.relu(F.conv1d(h, P["audio_encoder.input_proj.2.weight"],
                        P["audio_encoder.input_proj.2.bias"], padding=1))
    a = h.permute(0, 2, 1) + P["audio_encoder.pos_enc.pe"][:, :T]
    stages["audio_embed"] = a
    for l in range(cfg.num_encoder_layers):
        a = _encoder_layer(a, P, f"audio_encoder.transformer.layers.{l}", H)
    stages["audio_enc"] = a
    # VisualEncoder (model.py:103-117)
    _, N, Hh, Ww = lip_frames.shape
    x = lip_frames.reshape(B * N, 1, Hh, Ww)
    for idx in (0, 3, 6):
        x = F.conv2d(x, P[f"visual_encoder.conv.{idx}.weight"], P[f"visual_encoder.conv.{idx}.bias"],
                     stride=2, padding=1)
        bn = f"visual_encoder.conv.{idx + 1}"
        x = F.relu(F.batch_norm(x, P[f"{bn}.running_mean"], P[f"{bn}.running_var"],
                                P[f"{bn}.weight"], P[f"{bn}.bias"], False, 0.1, 1e-5))
    x = F.adaptive_avg_pool2d(x, (1, 1)).reshape(B * N, -1)
    stages["visual_pool"] = x.reshape(B, N, 128)
    v = F.linear(x, P["visual_encoder.frame_proj.weight"], P["visual_encoder.frame_proj.bias"])
    v = v.reshape(B, N, d) + P["visual_encoder.pos_enc.pe"][:, :N]
    stages["visual_embed"] = v
    for l in range(cfg.num_encoder_layers):
        v = _encoder_layer(v, P, f"visual_encoder.transformer.layers.{l}", H)
    stages["visual_enc"] = v
    v = F.interpolate(v.permute(0, 2, 1), size=T, mode="linear", align_corners=False).permute(0, 2, 1)
    stages["visual_interp"] = v
    # CrossModalFusion (model.py:145-149,166-173)
    f = a
    for l in range(cfg.num_fusion_layers):
        p = f"fusion.layers.{l}"
        h = F.layer_norm(f, (d,), P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], 1e-5)
        f = f + _mha(h, v, P, f"{p}.cross_attn", H, need_weights=True)
        h = F.layer_norm(f, (d,), P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], 1e-5)
        h = F.gelu(F.linear(h, P[f"{p}.ff.0.weight"], P[f"{p}.ff.0.bias"]))
        f = f + F.linear(h, P[f"{p}.ff.3.weight"], P[f"{p}.ff.3.bias"])
    f = F.layer_norm(f, (d,), P["fusion.norm.weight"], P["fusion.norm.bias"], 1e-5)
    stages["fused"] = f
    # SeparationDecoder (model.py:201-220)
    h = F.gelu(F.linear(f, P["decoder.decoder.0.weight"], P["decoder.decoder.0.bias"]))
    logits = F.linear(h, P["decoder.decoder.3.weight"], P["decoder.decoder.3.bias"])
    masks = torch.sigmoid(logits.view(B, T, cfg.num_speakers, cfg.freq_bins).permute(0, 2, 3, 1))
    separated = masks * mixed_spec.unsqueeze(1)
    if return_stages:
        return separated, masks, stages
    return separated, masks


def to_torch(P_np):
    """numpy state_dict (oracle.weights.make_state_dict) -> torch CPU tensors."""
    return {k: torch.from_numpy(v.copy()) if v.ndim else torch.tensor(int(v)) for k, v in P_np.items()}
