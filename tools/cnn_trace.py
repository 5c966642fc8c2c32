"""Phase timeline of the tcgen05 visual CNN kernel (debug hook): second frame group of every CTA."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer
torch.manual_seed(0)
m = AVSeparationTransformer().cuda(); m.prepack("cuda")
eng = m.engine
M = 12800
frames = torch.rand(M, 32, 32, device="cuda")
pooled = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(148 * 64, device="cuda", dtype=torch.int64)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    trace.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    assert eng.lib.avsep_test_visual_cnn_trace(eng.h, frames.data_ptr(), M, pooled.data_ptr(), trace.data_ptr(), s) == 0
    e1.record(); torch.cuda.synchronize()
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us (event incl. launch)")
t = trace.cpu().numpy().reshape(148, 64).astype(np.int64)
t = t[t[:, 0] > 0]
rel = (t - t[:, :1]) / 1e3
names = {0: "group start"}
for p in range(4):
    names[1 + p * 5] = f"pair{p} input staged"; names[2 + p * 5] = f"pair{p} conv1 done"
    names[3 + p * 5] = f"pair{p} deferred conv2 epilogue done"; names[4 + p * 5] = f"pair{p} conv2 slabs issued"
names.update({21: "last conv2 epilogue done", 22: "conv3 slabs issued", 23: "conv3 acc ready", 24: "group done"})
prev = 0.0
for k in sorted(names):
    v = rel[:, k].mean()
    print(f"  {names[k]:38s} at {v:7.2f} us  (+{v - prev:5.2f})")
    prev = v
