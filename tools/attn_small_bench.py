"""Single-tile attention kernel (Lq, Lk <= 64; 72 registers, 7 CTAs/SM) vs the generic mma.sync kernel."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig(257, 256, 4, 2, 2, 2, "bf16"), 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
H, hd = 4, 64
d = H * hd
for (B, L, lerp) in [(256, 63, 0), (256, 50, 0), (256, 63, 50), (64, 63, 0), (8, 63, 0)]:
    q = torch.randn(B, L, d, device="cuda").bfloat16()
    if lerp:
        k = torch.randn(B, lerp, d, device="cuda"); v = torch.randn(B, lerp, d, device="cuda")
    else:
        k = torch.randn(B, L, d, device="cuda").bfloat16(); v = torch.randn(B, L, d, device="cuda").bfloat16()
    res = {}
    for small in (0, 1):
        eng.set_option("attn_small", small)
        out = torch.zeros(B, L, d, device="cuda", dtype=torch.bfloat16)
        def run():
            assert eng.lib.avsep_test_attention(eng.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, hd, L, L, lerp, s) == 0
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): run()
        e1.record(); torch.cuda.synchronize()
        res[small] = (e0.elapsed_time(e1) / 50 * 1e3, out.clone())
    print(f"B={B} L={L} lerp={lerp}: generic {res[0][0]:.1f} us  single-tile {res[1][0]:.1f} us  identical={torch.equal(res[0][1], res[1][1])}", flush=True)
