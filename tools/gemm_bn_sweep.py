"""Time the tcgen05 GEMM (fp32-output test hook) for forced tile widths on the forward's K=256 shapes."""
import ctypes as C, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K) in [(16128, 256, 256), (12800, 256, 256), (16128, 512, 256), (16128, 768, 256), (12800, 768, 256), (12800, 1024, 256), (4032, 256, 256), (4032, 768, 256)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    line = f"M={M:6d} N={N:5d} K={K}:"
    for bn in (0, 64, 128, 192, 256):
        if bn and N % bn and bn != 192: pass
        def run():
            rc = eng.lib.avsep_test_gemm(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 0, bn, s)
            assert rc == 0, eng.lib.avsep_last_error(eng.h)
        try:
            for _ in range(3): run()
        except AssertionError as e:
            line += f"  bn={bn}: n/a"; continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): run()
        e1.record(); torch.cuda.synchronize()
        line += f"  bn={bn if bn else 'auto'}: {e0.elapsed_time(e1) / 50 * 1e3:5.1f} us"
    print(line, flush=True)
