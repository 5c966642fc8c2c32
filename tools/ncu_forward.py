"""Workload for the ncu captures under profiles/: N forwards of BASELINE configs[1] (B=256, 1 s clips) through the public
module call with CUDA graphs off, so that every kernel is an ordinary launch ncu can select by name.

    python tools/ncu_forward.py [--batch 256] [--iters 4]

One forward = 10 launches: prep_audio, gemm (conv1d_0, conv1d_2), visual_cnn_tc, gemm (frame_proj), xformer_stack x2
(audio, visual encoder), lerp_rows, gemm (cross_kv), xformer_stack (fusion + decoder)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer  # noqa: E402
from avsep_b200.synth import synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()
torch.manual_seed(0)
m = AVSeparationTransformer().cuda().eval()
m.prepack("cuda")
m.engine.set_option("use_graph", 0)
mixed, frames = synthetic_batch(args.batch, device="cuda")
for _ in range(args.iters):
    sep, masks = m(mixed, frames)
torch.cuda.synchronize()
print("ok", float(masks.mean()))
