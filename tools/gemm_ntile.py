"""MMA-rate check: same GEMM with n_tile = 64 / 128 / 256 (force_bn) on an MMA-bound shape (large K)."""
import ctypes as C, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K) in ((18944, 1024, 4096), (18944, 256, 4096)):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device="cuda")
    for bn in (64, 128, 256):
        ts = []
        for it in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(4):
                assert eng.lib.avsep_test_gemm(eng.h, A.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, 0, bn, s) == 0
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 4)
        t = min(ts)
        print(f"M={M} N={N} K={K} n_tile={bn}: {t*1e3:8.1f} us  {2*M*N*K/t/1e9:8.1f} TFLOP/s")
