"""One tcgen05 attention launch at the long-form shape (for ncu captures)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
B, L, H, hd = (int(x) for x in (sys.argv[1:5] + ["32", "1251", "4", "64"][len(sys.argv) - 1:]))
mode = int(os.environ.get("ATTN_TC", "2"))
eng = Engine(EngineConfig(257, 256, 4, 2, 2, 2, "bf16"), 0)
eng.set_option("attn_tc", mode)
d = H * hd
q = torch.randn(B, L, d, device="cuda").bfloat16(); k = torch.randn(B, L, d, device="cuda").bfloat16()
v = torch.randn(B, L, d, device="cuda").bfloat16(); out = torch.zeros_like(q)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert eng.lib.avsep_test_attention(eng.h, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, hd, L, L, 0, s) == 0
torch.cuda.synchronize()
print("ok")
