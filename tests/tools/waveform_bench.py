"""Times avsep_stft / avsep_istft at the bench workload (B=256, 1 s @ 8 kHz, S=2) with CUDA events; prints JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig   # noqa: E402

B, L, S = 256, 8000, 2
eng = Engine(EngineConfig(257, 256, 4, 2, 2, S), 0)
x = 0.3 * torch.randn(B, L, device="cuda")
masks = torch.rand(B, S, 257, 63, device="cuda")
spec, mag = eng.stft(x)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = torch.empty(B, S, L, device="cuda")
st = torch.cuda.current_stream().cuda_stream
t_s = timed(lambda: eng.lib.avsep_stft(eng.h, x.data_ptr(), B, L, 512, 128, spec.data_ptr(), mag.data_ptr(), st))
t_i = timed(lambda: eng.lib.avsep_istft(eng.h, spec.data_ptr(), masks.data_ptr(), B, S, 63, 512, 128, L, out.data_ptr(), st))
assert torch.equal(out, eng.istft(spec, masks, L))
io_s = B * L * 4 + B * 257 * 63 * 12
io_i = B * 257 * 63 * 8 + B * S * 257 * 63 * 4 + B * S * L * 4
print(json.dumps({"B": B, "L": L, "S": S, "stft_ms": round(t_s, 4), "stft_gb_s": round(io_s / t_s / 1e6, 1),
                  "istft_ms": round(t_i, 4), "istft_gb_s": round(io_i / t_i / 1e6, 1),
                  "note": "C-ABI calls on preallocated buffers; algorithmic bytes = inputs + outputs once"}))
