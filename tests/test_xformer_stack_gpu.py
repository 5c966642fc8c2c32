"""The fused transformer-stack kernel (csrc/xformer_stack_sm100.cu) through the C ABI against the oracle's torch port
of the same layers (oracle/avsep_oracle_torch.py: _encoder_layer / the fusion loop, pinned on the reference by the
golden fixtures), on the same seeded weights and inputs: encoder stacks (self-attention) for audio- and lip-shaped
sequences, the fusion stack (cross-attention on interpolated visual rows), ragged last tiles and multi-tile CTAs."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import avsep_oracle_torch as otorch
from oracle.weights import CONFIGS, make_state_dict
from tests.helpers import build_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    cfg = CONFIGS["default"]
    P = make_state_dict(cfg, seed=5, gain=2.0)
    model = build_model(cfg, P, "bf16")
    model.prepack()
    return cfg, otorch.to_torch(P), model.engine


def _run(eng, which, x, kv, B, L, want_x, want_op, final_ln):
    dev = torch.device("cuda", eng.device)
    xd = torch.from_numpy(x).to(dev).contiguous()
    kvd = kv.to(dev).contiguous() if kv is not None else None
    out_x = torch.full_like(xd, float("nan")) if want_x else None
    out_op = torch.full((B * L, x.shape[-1]), float("nan"), device=dev, dtype=torch.bfloat16) if want_op else None
    rc = eng.lib.avsep_test_xformer_stack(eng.h, which, xd.data_ptr(), kvd.data_ptr() if kvd is not None else None, B, L,
                                          out_x.data_ptr() if want_x else None, out_op.data_ptr() if want_op else None,
                                          1 if final_ln else 0, None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, eng.lib.avsep_last_error(eng.h).decode()
    torch.cuda.synchronize()
    return (out_x.cpu().numpy() if want_x else None), (out_op.float().cpu().numpy() if want_op else None)


def _enc_ref(P, cfg, pre, x):
    a = torch.from_numpy(x)
    with torch.no_grad():
        for l in range(cfg.num_encoder_layers):
            a = otorch._encoder_layer(a, P, f"{pre}.transformer.layers.{l}", cfg.nhead)
    return a


@pytest.mark.parametrize("which,pre,B,L", [(0, "audio_encoder", 2, 63), (1, "visual_encoder", 2, 50),
                                           (0, "audio_encoder", 5, 63), (1, "visual_encoder", 7, 50),
                                           (0, "audio_encoder", 3, 128), (0, "audio_encoder", 9, 20),
                                           (1, "visual_encoder", 1, 1), (0, "audio_encoder", 300, 63)])
def test_encoder_stack_matches_oracle(setup, which, pre, B, L):
    cfg, P, eng = setup
    rng = np.random.default_rng(100 * which + B + L)
    x = rng.standard_normal((B, L, cfg.d_model)).astype(np.float32)
    ref = _enc_ref(P, cfg, pre, x)
    got_x, got_op = _run(eng, which, x.reshape(B * L, -1), None, B, L, True, True, which == 0)
    assert np.isfinite(got_x).all() and np.isfinite(got_op).all()
    refn = ref.numpy().reshape(B * L, -1)
    scale = max(1.0, float(np.abs(refn).max()))
    err_x = float(np.abs(got_x - refn).max()) / scale
    if which == 0:       # followed by fusion layer 0's norm1 in the model
        op_ref = F.layer_norm(ref, (cfg.d_model,), P["fusion.layers.0.norm1.weight"], P["fusion.layers.0.norm1.bias"], 1e-5)
    else:
        op_ref = ref
    op_ref = op_ref.numpy().reshape(B * L, -1)
    err_op = float(np.abs(got_op - op_ref).max()) / max(1.0, float(np.abs(op_ref).max()))
    print(f"stack {pre} B={B} L={L}: err_x={err_x:.3e} err_op={err_op:.3e} (|x| max {scale:.1f})")
    assert err_x < 2e-2 and err_op < 3e-2, (err_x, err_op)


@pytest.mark.parametrize("B,T,N", [(2, 63, 50), (5, 63, 50), (3, 32, 10), (260, 63, 50)])
def test_fusion_stack_matches_oracle(setup, B, T, N):
    cfg, P, eng = setup
    d, Lf = cfg.d_model, cfg.num_fusion_layers
    rng = np.random.default_rng(B + T)
    a = rng.standard_normal((B, T, d)).astype(np.float32)
    v = rng.standard_normal((B, N, d)).astype(np.float32)
    with torch.no_grad():
        vi = F.interpolate(torch.from_numpy(v).permute(0, 2, 1), size=T, mode="linear", align_corners=False).permute(0, 2, 1)
        f = torch.from_numpy(a)
        kv_cols = []
        for l in range(Lf):
            p = f"fusion.layers.{l}"
            h = F.layer_norm(f, (d,), P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], 1e-5)
            f = f + otorch._mha(h, vi, P, f"{p}.cross_attn", cfg.nhead, need_weights=True)
            h = F.layer_norm(f, (d,), P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], 1e-5)
            h = F.gelu(F.linear(h, P[f"{p}.ff.0.weight"], P[f"{p}.ff.0.bias"]))
            f = f + F.linear(h, P[f"{p}.ff.3.weight"], P[f"{p}.ff.3.bias"])
            W, bias = P[f"{p}.cross_attn.in_proj_weight"], P[f"{p}.cross_attn.in_proj_bias"]
            kv_cols.append(F.linear(vi, W[d:], bias[d:]))          # K | V rows of this layer (what gemm.cross_kv produces)
        fused = F.layer_norm(f, (d,), P["fusion.norm.weight"], P["fusion.norm.bias"], 1e-5)
        kv = torch.cat(kv_cols, dim=-1).reshape(B * T, Lf * 2 * d).to(torch.bfloat16)
    got_x, got_op = _run(eng, 2, a.reshape(B * T, d), kv, B, T, True, True, True)
    assert np.isfinite(got_x).all() and np.isfinite(got_op).all()
    fn, on = f.numpy().reshape(B * T, d), fused.numpy().reshape(B * T, d)
    err_x = float(np.abs(got_x - fn).max()) / max(1.0, float(np.abs(fn).max()))
    err_op = float(np.abs(got_op - on).max()) / max(1.0, float(np.abs(on).max()))
    print(f"fusion stack B={B} T={T} N={N}: err_x={err_x:.3e} err_op={err_op:.3e}")
    assert err_x < 2e-2 and err_op < 3e-2, (err_x, err_op)


def test_outputs_are_selectable_and_rows_outside_the_batch_untouched(setup):
    cfg, P, eng = setup
    B, L = 3, 63
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B * L, cfg.d_model)).astype(np.float32)
    x_only, _ = _run(eng, 0, x, None, B, L, True, False, False)
    _, op_only = _run(eng, 0, x, None, B, L, False, True, True)
    both_x, both_op = _run(eng, 0, x, None, B, L, True, True, True)
    assert np.array_equal(x_only, both_x) and np.array_equal(op_only, both_op)


@pytest.mark.parametrize("B,T,N", [(1, 63, 50), (5, 63, 50), (3, 32, 10), (9, 20, 7), (300, 63, 50)])
def test_fusion_stack_with_fused_decoder_matches_oracle(setup, B, T, N):
    """CrossModalFusion + SeparationDecoder in the one kernel the forward launches (avsep_test_fusion_decoder): masks and
    separated against the oracle's fusion loop + decoder (model.py:166-173,201-220) on the same K|V rows; odd batches
    leave half a tile empty, B=300 gives every CTA more than one tile, short clips pack 4 utterances per tile."""
    cfg, P, eng = setup
    d, Lf, S, Fq = cfg.d_model, cfg.num_fusion_layers, cfg.num_speakers, cfg.freq_bins
    rng = np.random.default_rng(7 * B + T)
    a = rng.standard_normal((B, T, d)).astype(np.float32)
    v = rng.standard_normal((B, N, d)).astype(np.float32)
    mixed = rng.random((B, Fq, T)).astype(np.float32) * 3.0
    with torch.no_grad():
        vi = F.interpolate(torch.from_numpy(v).permute(0, 2, 1), size=T, mode="linear", align_corners=False).permute(0, 2, 1)
        f = torch.from_numpy(a)
        kv_cols = []
        for l in range(Lf):
            p = f"fusion.layers.{l}"
            h = F.layer_norm(f, (d,), P[f"{p}.norm1.weight"], P[f"{p}.norm1.bias"], 1e-5)
            f = f + otorch._mha(h, vi, P, f"{p}.cross_attn", cfg.nhead, need_weights=True)
            h = F.layer_norm(f, (d,), P[f"{p}.norm2.weight"], P[f"{p}.norm2.bias"], 1e-5)
            h = F.gelu(F.linear(h, P[f"{p}.ff.0.weight"], P[f"{p}.ff.0.bias"]))
            f = f + F.linear(h, P[f"{p}.ff.3.weight"], P[f"{p}.ff.3.bias"])
            W, bias = P[f"{p}.cross_attn.in_proj_weight"], P[f"{p}.cross_attn.in_proj_bias"]
            kv_cols.append(F.linear(vi, W[d:], bias[d:]))
        f = F.layer_norm(f, (d,), P["fusion.norm.weight"], P["fusion.norm.bias"], 1e-5)
        h = F.gelu(F.linear(f, P["decoder.decoder.0.weight"], P["decoder.decoder.0.bias"]))
        logits = F.linear(h, P["decoder.decoder.3.weight"], P["decoder.decoder.3.bias"])
        masks_ref = torch.sigmoid(logits.view(B, T, S, Fq).permute(0, 2, 3, 1)).numpy()
        kv = torch.cat(kv_cols, dim=-1).reshape(B * T, Lf * 2 * d).to(torch.bfloat16)
    dev = torch.device("cuda", eng.device)
    xd = torch.from_numpy(a.reshape(B * T, d)).to(dev).contiguous()
    kvd = kv.to(dev).contiguous()
    md = torch.from_numpy(mixed).to(dev).contiguous()
    # one guard utterance behind the outputs: the kernel must not write past the batch
    sep = torch.full((B + 1, S, Fq, T), float("nan"), device=dev)
    msk = torch.full((B + 1, S, Fq, T), float("nan"), device=dev)
    rc = eng.lib.avsep_test_fusion_decoder(eng.h, xd.data_ptr(), kvd.data_ptr(), B, T, md.data_ptr(), sep.data_ptr(),
                                           msk.data_ptr(), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, eng.lib.avsep_last_error(eng.h).decode()
    torch.cuda.synchronize()
    assert torch.isnan(sep[B]).all() and torch.isnan(msk[B]).all()
    got_m, got_s = msk[:B].cpu().numpy(), sep[:B].cpu().numpy()
    assert np.isfinite(got_m).all() and np.isfinite(got_s).all()
    err = float(np.abs(got_m - masks_ref).max())
    print(f"fusion+decoder B={B} T={T} N={N}: max|d masks| = {err:.3e}")
    assert err < 1e-2, err                                            # BASELINE.json north_star: masks within 1e-2 (bf16)
    assert np.array_equal(got_s, got_m * mixed[:, None])             # separated is exactly masks * mixed_spec (fp32)
