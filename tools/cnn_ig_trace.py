"""Phase timeline of the shifted-view implicit-GEMM CNN kernel (visual_cnn_ig_sm100.cu): clock64 stamps of the second
frame group of every CTA (builder thread 0 and the MMA thread), median over CTAs.  Debug aid."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer
torch.manual_seed(0)
m = AVSeparationTransformer().cuda(); m.prepack("cuda")
eng = m.engine
M = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
frames = torch.rand(M, 32, 32, device="cuda")
pooled = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    trace.zero_(); torch.cuda.synchronize()
    assert eng.lib.avsep_test_visual_cnn_trace(eng.h, frames.data_ptr(), M, pooled.data_ptr(), trace.data_ptr(), s) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    eng.lib.avsep_test_visual_cnn(eng.h, frames.data_ptr(), M, 32, 32, pooled.data_ptr(), s)
e1.record(); torch.cuda.synchronize()
print(f"M={M}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (no trace)")
t = trace.cpu().numpy().reshape(148, 64).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, :1]
bn = {}
for f in range(3):
    bn.update({f * 8: f"A frame{f} slot start", f * 8 + 3: f"A frame{f} staging stores issued", f * 8 + 4: f"A frame{f} barrier passed",
               f * 8 + 1: f"A frame{f} act1 buffer free", f * 8 + 5: f"A frame{f} conv1 stores issued", f * 8 + 6: f"A frame{f} proxy fence done",
               f * 8 + 2: f"A frame{f} arrived"})
en = {}
for f in range(3):
    en.update({48 + f * 4: f"B frame{f} waiting acc2", 49 + f * 4: f"B frame{f} conv2 epilogue done", 50 + f * 4: f"B frame{f} (+conv3 epilogue) done"})
mn = {}
for f in range(3):
    mn.update({32 + f * 2: f"frame{f} act1 seen", 33 + f * 2: f"frame{f} conv2 issued"})
mn.update({40: "conv3 start", 41: "act2 seen", 42: "conv3 issued"})
for title, names in (("group A thread 0", bn), ("group B thread 256", en), ("MMA thread", mn)):
    print(title)
    prev = None
    for k in sorted(names, key=lambda k: float(np.median(t[:, k] - t0[:, 0])) if not (t[:, k] == 0).all() else 1e18):
        col = t[:, k]
        if (col == 0).all():
            continue
        v = float(np.median(col - t0[:, 0]))
        print(f"  {names[k]:30s} {v:9.0f} clk" + ("" if prev is None else f"  (+{v - prev:7.0f})"))
        prev = v
