"""Per-phase timeline of the tcgen05 GEMM (debug hook avsep_test_gemm_trace): mean/max ns per phase over CTAs."""
import ctypes as C, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200.engine import Engine, EngineConfig

eng = Engine(EngineConfig(65, 64, 4, 1, 1, 2, "bf16"), 0)
names = ["setup", "first_load", "mma_issue", "acc_ready", "acc_drained", "epi_done", "cta_done"]
for (M, N, K, ln, act, tag) in [(16128, 256, 256, 1, 0, "out_proj+LN"), (16128, 256, 1024, 1, 0, "ffn2+LN"),
                                (16128, 1024, 256, 0, 1, "ffn1 relu"), (16128, 768, 256, 0, 0, "qkv"),
                                (12800, 256, 256, 1, 0, "out_proj+LN visual")]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    g = torch.randn(N, device="cuda"); b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    trace = torch.zeros(148 * 8, device="cuda", dtype=torch.int64)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for it in range(3):
        trace.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = eng.lib.avsep_test_gemm_trace(eng.h, A.data_ptr(), W.data_ptr(), bias.data_ptr(), x.data_ptr(), g.data_ptr(),
                                           b.data_ptr(), out.data_ptr(), M, N, K, ln, act, trace.data_ptr(), s)
        assert rc == 0, eng.lib.avsep_last_error(eng.h)
        e1.record(); torch.cuda.synchronize()
    t = trace.cpu().numpy().reshape(148, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = t - t0
    print(f"== {tag}: M={M} N={N} K={K} ctas={len(t)} kernel {e0.elapsed_time(e1)*1e3:.1f} us (event)")
    print("   entry spread (us): max", (rel[:, 0].max()) / 1e3)
    prev = rel[:, 0]
    for k, nm in enumerate(names, start=1):
        col = rel[:, k]
        ok = t[:, k] > 0
        if not ok.any():
            continue
        print(f"   {nm:12s} at mean {col[ok].mean()/1e3:7.2f} us  max {col[ok].max()/1e3:7.2f} us")
