"""Time the two attention kernels (mma.sync flash kernel vs tcgen05 kernel) on the shapes the forward uses."""
import math, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from avsep_b200.engine import Engine, EngineConfig
import ctypes as C

eng = Engine(EngineConfig(257, 256, 4, 2, 2, 2, "bf16"), 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
H, hd = 4, 64
d = H * hd
for (B, L, lerp) in [(256, 63, 0), (256, 50, 0), (256, 63, 50), (32, 1251, 0), (32, 500, 0), (32, 1251, 500), (8, 63, 0),
                     (64, 313, 0)]:
    Lk = L
    q = torch.randn(B, L, d, device="cuda").bfloat16()
    if lerp:
        k = torch.randn(B, lerp, d, device="cuda"); v = torch.randn(B, lerp, d, device="cuda")
    else:
        k = torch.randn(B, Lk, d, device="cuda").bfloat16(); v = torch.randn(B, Lk, d, device="cuda").bfloat16()
    out = torch.zeros(B, L, d, device="cuda", dtype=torch.bfloat16)
    res = {}
    for mode in (0, 2):
        eng.set_option("attn_tc", mode)
        outs = torch.zeros_like(out)
        def run():
            rc = eng.lib.avsep_test_attention(eng.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), outs.data_ptr(), B, H, hd, L, Lk, lerp, s)
            assert rc == 0, eng.lib.avsep_last_error(eng.h).decode()
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[mode] = (ms, outs.clone())
    fl = 4.0 * B * H * L * Lk * hd
    diff = (res[0][1].float() - res[2][1].float()).abs().max().item()
    print(f"B={B} L={L} lerp={lerp}: mma.sync {res[0][0]*1e3:.1f} us ({fl/res[0][0]/1e9:.0f} TF/s)  tcgen05 {res[2][0]*1e3:.1f} us "
          f"({fl/res[2][0]/1e9:.0f} TF/s)  maxdiff {diff:.4f}", flush=True)
