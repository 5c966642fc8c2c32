"""Phase timeline of the fused FFN kernel (debug hook avsep_test_ffn_fused_trace)."""
import ctypes as C, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig
eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
for M, act in ((16128, 1), (16128, 2)):
    d, hid = 256, 1024
    a = torch.randn(M, d, device="cuda").bfloat16()
    w1 = (torch.randn(hid, d, device="cuda") / math.sqrt(d)).bfloat16(); b1 = torch.randn(hid, device="cuda")
    w2 = (torch.randn(d, hid, device="cuda") / math.sqrt(hid)).bfloat16(); b2 = torch.randn(d, device="cuda")
    x = torch.randn(M, d, device="cuda"); g = torch.randn(d, device="cuda"); b = torch.randn(d, device="cuda")
    out = torch.zeros(M, d, device="cuda", dtype=torch.bfloat16)
    trace = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        trace.zero_(); torch.cuda.synchronize()
        assert eng.lib.avsep_test_ffn_fused_trace(eng.h, a.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), act,
                                                  x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), M, trace.data_ptr(), s) == 0
        torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(148, 64).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    t = t[t[:, 8] > 0]          # CTAs whose MMA thread stamped (with cta_group::2 pairs: the leaders)
    r = (t - t0) / 1e3
    print(f"== fused FFN M={M} act={act} ctas={len(t)} (us since first CTA entry, mean over CTAs)")
    print(f"   A landed {r[:,1].mean():.2f}   acc2 complete {r[:,2].mean():.2f}   epilogue-2 done {r[:,3].mean():.2f} (max {r[:,3].max():.2f})")
    for j in range(8):
        c = r[:, 8 + 4 * j: 12 + 4 * j].mean(axis=0)
        print(f"   chunk {j}: GEMM1 issued {c[0]:6.2f}  acc1 ready {c[2]:6.2f}  H published {c[3]:6.2f}  GEMM2 issued {c[1]:6.2f}")
