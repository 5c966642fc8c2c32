"""A few audio-encoder-only invocations at B=256 (prep, conv1d_0, conv1d_2, two encoder layers), for ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer
m = AVSeparationTransformer().cuda().eval(); m.prepack("cuda")
eng = m.engine
mixed = torch.rand(256, 257, 63, device="cuda") * 50
for _ in range(4):
    out = eng.audio_encoder(mixed)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
