"""The reference's algorithm as library-eager PyTorch on the same B200 (recommended baseline): the oracle's torch
port (the operators the reference's nn.Modules dispatch to) run on cuda in fp32 (TF32 off / on) and under bf16 autocast,
next to this repo's forward.  Test infrastructure: documentation numbers only, not part of bench.py."""
import json, os, sys
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from oracle import avsep_oracle_torch as otorch
from oracle.weights import CONFIGS, make_state_dict
from avsep_b200 import AVSeparationTransformer
from avsep_b200.synth import synthetic_batch

cfg = CONFIGS["default"]
B = 256
P = {k: v.cuda() for k, v in otorch.to_torch(make_state_dict(cfg, seed=1, gain=1.0)).items()}
mixed, frames = synthetic_batch(B, device="cuda")

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
with torch.no_grad():
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    out["torch eager fp32 (TF32 off)"] = timed(lambda: otorch.forward(P, cfg, mixed, frames))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    out["torch eager fp32 (TF32 on)"] = timed(lambda: otorch.forward(P, cfg, mixed, frames))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out["torch eager bf16 autocast"] = timed(lambda: otorch.forward(P, cfg, mixed, frames))
m = AVSeparationTransformer().cuda().eval(); m.prepack("cuda")
out["this repo (bf16, graph replay)"] = timed(lambda: m(mixed, frames), 50)
print(json.dumps({k: {"ms_per_forward": round(v, 3), "utt_s_per_s": round(B / v * 1e3)} for k, v in out.items()}, indent=1))
