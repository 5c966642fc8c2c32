"""End-to-end (host buffers in, host buffers out) throughput of avsep_forward_host vs chunk size / lanes, next to the
raw pinned-copy times of the same bytes."""
import json, os, sys, time
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200")); sys.path.insert(0, ROOT)
from avsep_b200 import AVSeparationTransformer
from avsep_b200.synth import synthetic_batch

B = 256
m = AVSeparationTransformer().eval(); m.prepack("cuda")
eng = m.engine
mixed, frames = synthetic_batch(B, device="cpu")
mixed, frames = mixed.pin_memory(), frames.pin_memory()
sep = torch.empty(B, 2, 257, 63).pin_memory(); masks = torch.empty_like(sep).pin_memory()
dm, df = mixed.cuda(), frames.cuda(); ds, dk = sep.cuda(), masks.cuda()
def timed(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d():
    dm.copy_(mixed, non_blocking=True); df.copy_(frames, non_blocking=True)
def d2h():
    sep.copy_(ds, non_blocking=True); masks.copy_(dk, non_blocking=True)
def both():
    with torch.cuda.stream(s1): h2d()
    with torch.cuda.stream(s2): d2h()
print(json.dumps({"h2d_ms": timed(h2d), "d2h_ms": timed(d2h), "both_ms": timed(both),
                  "h2d_MB": (mixed.numel() + frames.numel()) * 4 / 1e6, "d2h_MB": 2 * sep.numel() * 4 / 1e6}), flush=True)
NS = int(os.environ.get("E2E_SLOTS", "2"))
outs = [(sep, masks)] + [(torch.empty_like(sep).pin_memory(), torch.empty_like(masks).pin_memory()) for _ in range(NS - 1)]
def stream_run(n):
    for i in range(n):
        sl = i % NS
        if i >= NS: eng.host_wait(sl)
        eng.forward_host_async(mixed, frames, outs[sl][0], outs[sl][1], sl)
    for sl in range(NS): eng.host_wait(sl)
for chunk, lanes in [tuple(int(v) for v in x.split("x")) for x in os.environ.get("E2E_GRID", "64x1,64x2,96x2,128x1,128x2,256x1,32x2,48x2").split(",")]:
    eng.set_option("host_chunk", chunk); eng.set_option("host_lanes", lanes)
    for _ in range(3): eng.forward_host(mixed, frames, sep, masks)
    ms = timed(lambda: eng.forward_host(mixed, frames, sep, masks), 10)
    stream_run(6); torch.cuda.synchronize(); t0 = time.perf_counter(); stream_run(20); torch.cuda.synchronize()
    ms2 = (time.perf_counter() - t0) / 20 * 1e3
    print(json.dumps({"chunk": chunk, "lanes": lanes, "sync_ms": round(ms, 3), "sync_utt_s_per_s": round(B / ms * 1e3),
                      "stream_ms": round(ms2, 3), "stream_utt_s_per_s": round(B / ms2 * 1e3)}), flush=True)
