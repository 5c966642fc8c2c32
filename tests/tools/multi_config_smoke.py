"""Small end-to-end invocations of every forward-path kernel in several configurations, both precisions, plus synthesis
and SNR evaluation (test infrastructure: it uses the oracle's weight / input generator)."""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from helpers import build_model
from oracle.weights import CONFIGS, make_state_dict, make_inputs
from avsep_b200.dataset import SyntheticAVDataset

for cfg_name, prec, B, T, N, hw in (("default", "bf16", 40, 63, 50, 32), ("default", "bf16", 3, 300, 120, 32), ("tiny2", "bf16", 3, 20, 30, 15),
                                    ("default", "tf32", 2, 63, 50, 32), ("scaled", "bf16", 2, 63, 50, 32)):
    cfg = CONFIGS[cfg_name]
    P = make_state_dict(cfg, seed=1, gain=2.0)
    mixed, frames = make_inputs(cfg, B, T, N, hw, hw, seed=2, kind="randn")
    m = build_model(cfg, P, prec)
    for _ in range(3):
        sep, masks = m(torch.from_numpy(mixed).cuda(), torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(sep).all()
    if cfg_name == "default" and T == 63 and prec == "bf16":
        sh, mh = m(torch.from_numpy(mixed).pin_memory(), torch.from_numpy(frames).pin_memory())
        assert torch.isfinite(sh).all()
    print("ok", cfg_name, prec, B, T, flush=True)
ds = SyntheticAVDataset(num_samples=16)
b = ds.batch(range(5))
i_snr, o_snr, perm, si = ds.engine.eval_snr(b["clean_specs"].flip(1).contiguous(), b["clean_specs"], b["mixed_spec"])
torch.cuda.synchronize()
print("ok synth/snr", float(o_snr.mean()))
