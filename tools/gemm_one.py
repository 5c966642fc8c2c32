"""Run one GEMM configuration a few times (for ncu captures): python tools/gemm_one.py M N K ln act"""
import ctypes as C, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
M, N, K, ln, act = (int(a) for a in sys.argv[1:6])
eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda"); x = torch.randn(M, N, device="cuda")
g = torch.randn(N, device="cuda"); b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(4):
    rc = eng.lib.avsep_test_gemm_trace(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), x.data_ptr(), g.data_ptr(),
                                       b.data_ptr(), out.data_ptr(), M, N, K, ln, act, None, s)
    assert rc == 0
torch.cuda.synchronize()
print("ok")
