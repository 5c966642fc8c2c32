"""HBM rate of avsep_separate (SeparationDecoder.separate, model.py:210-220) at the benchmark shapes: algorithmic bytes
(masks in + mixture in + separated out = 4*F*T*(2S+1) per utterance) / device time, CUDA events, rotating buffer sets
larger than L2, against MEASURED_PEAKS.json's copy bandwidth.  Also the shape of the 8-GPU root rebuild (7 x 256)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig   # noqa: E402

F, S, T = 257, 2, 63
eng = Engine(EngineConfig(F, 256, 4, 2, 2, S), 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
st = torch.cuda.current_stream().cuda_stream
for B in (256, 1792):
    sets = [(torch.rand(B, S, F, T, device="cuda"), torch.randn(B, F, T, device="cuda"), torch.empty(B, S, F, T, device="cuda"))
            for _ in range(3)]
    assert torch.equal(eng.separate(sets[0][0], sets[0][1]), sets[0][0] * sets[0][1].unsqueeze(1))

    def call(i):
        m, x, o = sets[i % 3]
        rc = eng.lib.avsep_separate(eng.h, m.data_ptr(), x.data_ptr(), B, T, o.data_ptr(), st)
        assert rc == 0

    for i in range(6):
        call(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    a.record()
    for i in range(n):
        call(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    nbytes = 4 * F * T * (2 * S + 1) * B
    print(json.dumps({"kernel": "separate", "B": B, "ms": round(ms, 4), "algorithmic_bytes": nbytes,
                      "gb_s": round(nbytes / ms / 1e6, 1), "peak_gb_s": peak, "frac": round(nbytes / ms / 1e6 / peak, 3)}),
          flush=True)
    del sets
