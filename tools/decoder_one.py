"""A few decoder-only invocations at B=256 (dec0 GEMM + decoder-tail GEMM), for ncu captures of the tail kernel."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer
m = AVSeparationTransformer().cuda().eval(); m.prepack("cuda")
eng = m.engine
B, T = 256, 63
fused = torch.randn(B, T, 256, device="cuda"); mixed = torch.rand(B, 257, T, device="cuda") * 50
for _ in range(4):
    sep, masks = eng.decoder(fused, mixed)
torch.cuda.synchronize()
print("ok", float(masks.mean()))
