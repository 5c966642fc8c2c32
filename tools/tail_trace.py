"""Phase timeline of the decoder-head GEMM (EPI_TAIL, M = B*T = 16128, N = 2*257, K = 512) through the trace hook:
per tile of each CTA - first operands landed, all MMAs issued, accumulator ready, epilogue done (us since the first
CTA entered; globaltimer ticks every 512 ns, numbers are means over CTAs)."""
import ctypes as C, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
B, T, F, S, K = 256, 63, 257, 2, 512
M, N = B * T, S * F
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
mixed = torch.rand(B, F, T, device="cuda")
masks = torch.empty(B, S, F, T, device="cuda")
sep = torch.empty_like(masks)
trace = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    trace.zero_()
    torch.cuda.synchronize()
    rc = eng.lib.avsep_test_gemm_trace(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), masks.data_ptr(), mixed.data_ptr(),
                                       None, sep.data_ptr(), M, N, K, 7, T, trace.data_ptr(), s)
    assert rc == 0, eng.lib.avsep_last_error(eng.h)
    torch.cuda.synchronize()
ref = torch.sigmoid((A.float() @ W.float().t() + bias)[:T]).t().reshape(S, F, T)
print("max |masks - ref| on utterance 0:", (masks[0] - ref).abs().max().item())
t = trace.cpu().numpy().reshape(148, 64).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print(f"ctas={len(t)}  setup done {(t[:,1]-t0).mean()/1e3:.2f}  CTA done mean {(t[:,7]-t0).mean()/1e3:.2f} max {(t[:,7]-t0).max()/1e3:.2f}")
for lt in range(4):
    cols = t[:, 8 + 4 * lt: 12 + 4 * lt]
    ok = cols[:, 3] > 0
    if not ok.any():
        break
    r = (cols[ok] - t0) / 1e3
    print(f"tile {lt} ({ok.sum()} CTAs): operands {r[:,0].mean():.2f}  MMAs issued {r[:,1].mean():.2f}  acc ready {r[:,2].mean():.2f}  epilogue done {r[:,3].mean():.2f} (max {r[:,3].max():.2f})")
