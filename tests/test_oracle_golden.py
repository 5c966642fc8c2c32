"""CPU: the oracle restatements against golden vectors produced by the real reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import numpy as np
import pytest
import torch

from oracle import avsep_oracle as onp
from oracle import avsep_oracle_torch as otorch
from oracle.weights import CONFIGS, ModelConfig, make_state_dict, num_parameters, state_dict_spec
from tests.helpers import GOLDEN_CASES, STAGES, case_tensors, load_golden, subsample_stage

SMALL = ("tiny_up", "tiny_down", "c1_dataset", "c1_randn")


def test_known_answer_parameter_count():
    # README.md:60 of the reference: the d_model=128 demo model has 1,612,738 parameters
    assert num_parameters(ModelConfig(257, 128, 4, 2, 2, 2)) == 1612738
    assert num_parameters(CONFIGS["default"]) == 5654978
    assert num_parameters(CONFIGS["scaled"]) == 59400899


def test_state_dict_is_layerwise_distinct():
    P = make_state_dict(CONFIGS["default"], seed=0)
    a = P["audio_encoder.transformer.layers.0.linear1.weight"]
    b = P["audio_encoder.transformer.layers.1.linear1.weight"]
    assert np.abs(a - b).max() > 1e-3
    assert np.abs(P["visual_encoder.conv.1.running_mean"]).max() > 0
    assert len(P) == len(state_dict_spec(CONFIGS["default"]))


@pytest.mark.parametrize("name", SMALL)
def test_numpy_oracle_matches_reference_golden(name):
    z, meta = load_golden(name)
    cfg, P, mixed, frames = case_tensors(meta)
    sep, masks, stages = onp.forward(P, cfg, mixed, frames, return_stages=True)
    sf, st = meta["stride_f"], meta["stride_t"]
    assert np.abs(masks[:, :, ::sf, ::st] - z["masks"]).max() < 2e-5
    assert (np.abs(sep[:, :, ::sf, ::st] - z["separated"]) / max(1.0, np.abs(mixed).max())).max() < 2e-5
    for k in STAGES:
        ref = z["stage_" + k]
        assert np.abs(subsample_stage(stages[k]) - ref).max() < 2e-4, k
    assert masks.min() >= 0.0 and masks.max() <= 1.0          # reference tests/test_model.py:163-170


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_torch_port_matches_reference_golden(name):
    z, meta = load_golden(name)
    cfg, P, mixed, frames = case_tensors(meta)
    Pt = otorch.to_torch(P)
    sep, masks = otorch.forward(Pt, cfg, torch.from_numpy(mixed), torch.from_numpy(frames))
    sep, masks = sep.numpy(), masks.numpy()
    sf, st = meta["stride_f"], meta["stride_t"]
    assert np.abs(masks[:, :, ::sf, ::st] - z["masks"]).max() < 2e-5
    assert (np.abs(sep[:, :, ::sf, ::st] - z["separated"]) / max(1.0, np.abs(mixed).max())).max() < 2e-5
    assert abs(masks.astype(np.float64).sum() - float(z["masks_sum"])) < 1e-3 * masks.size ** 0.5 + 1e-2


def test_interp_rule_matches_torch():
    # F.interpolate(linear, align_corners=False) for up- and down-sampling (tests/test_model.py:105-114)
    import torch.nn.functional as F
    rng = np.random.default_rng(0)
    for N, T in ((10, 32), (10, 20), (30, 20), (50, 63), (500, 1251), (7, 7), (1, 5)):
        x = rng.standard_normal((2, N, 8)).astype(np.float32)
        ref = F.interpolate(torch.from_numpy(x).permute(0, 2, 1), size=T, mode="linear",
                            align_corners=False).permute(0, 2, 1).numpy()
        # ATen contracts scale*(j+0.5)-0.5 into an FMA; the un-fused numpy form differs by <= 1/4 ulp(src) (SURVEY App. B rule 6)
        assert np.abs(onp.interp_linear_time(x, T) - ref).max() < 2e-5 + 4e-7 * N, (N, T)
