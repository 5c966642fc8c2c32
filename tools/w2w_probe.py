"""Splits bench.py's waveform_to_waveform step (avsep_stft -> avsep_forward -> avsep_istft) into its calls."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
import bench  # noqa: E402
from avsep_b200.synth import synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
B, F, S, T, N, HW, L = 256, 257, 2, 63, 50, 32, 8000
model = bench.build_state().to(dev)
model.prepack(dev)
eng = model.engine
st = torch.cuda.current_stream().cuda_stream
sets = [synthetic_batch(B, F, T, N, HW, HW, seed=100003 * i, device=dev) for i in range(3)]
g = torch.Generator(device=dev).manual_seed(7)
waves = [0.3 * torch.randn(B, L, device=dev, generator=g) for _ in range(3)]
spec = torch.empty(B, F, T, device=dev, dtype=torch.complex64)
mag = torch.empty(B, F, T, device=dev)
sep = torch.empty(B, S, F, T, device=dev)
masks = torch.empty_like(sep)
out = torch.empty(B, S, L, device=dev)


def stft(i):
    return eng.lib.avsep_stft(eng.h, waves[i % 3].data_ptr(), B, L, 512, 128, spec.data_ptr(), mag.data_ptr(), st)


def fwd(i, m=None):
    m = mag if m is None else m
    return eng.lib.avsep_forward(eng.h, m.data_ptr(), sets[i % 3][1].data_ptr(), B, T, N, HW, HW, sep.data_ptr(),
                                 masks.data_ptr(), None, 0, st)


def istft(i):
    return eng.lib.avsep_istft(eng.h, spec.data_ptr(), masks.data_ptr(), B, S, T, 512, 128, L, out.data_ptr(), st)


def timed(label, fn, n=50):
    for i in range(6):
        assert fn(i) in (0, None), eng.lib.avsep_last_error(eng.h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b) / n:.4f} ms", flush=True)


timed("stft", stft)
timed("forward(mag)", fwd)
timed("forward(dataset-shaped)", lambda i: fwd(i, sets[i % 3][0]))
timed("istft", istft)
timed("stft+forward", lambda i: stft(i) or fwd(i))
timed("forward+istft", lambda i: fwd(i) or istft(i))
timed("stft+forward+istft", lambda i: stft(i) or fwd(i) or istft(i))
print("mag max", float(mag.max()), "finite", bool(torch.isfinite(mag).all()), "masks finite", bool(torch.isfinite(masks).all()))
