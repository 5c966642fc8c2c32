"""Throughput with one vs two forwards in flight (two streams, two workspaces, two output sets) at B=256."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer
from avsep_b200.synth import synthetic_batch
B, T, N, HW = 256, 63, 50, 32
m = AVSeparationTransformer().cuda().eval(); m.prepack("cuda")
eng = m.engine
sets = [synthetic_batch(B, seed=s, device="cuda") for s in range(4)]
wsb = eng.lib.avsep_workspace_bytes(eng.h, B, T, N, HW, HW)
ws = [torch.empty(wsb + 1024, dtype=torch.uint8, device="cuda") for _ in range(2)]
wsp = [(w.data_ptr() + 1023) // 1024 * 1024 for w in ws]
outs = [(torch.empty(B, 2, 257, T, device="cuda"), torch.empty(B, 2, 257, T, device="cuda")) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def fwd(i, lane):
    mixed, frames = sets[i % 4]
    rc = eng.lib.avsep_forward(eng.h, mixed.data_ptr(), frames.data_ptr(), B, T, N, HW, HW, outs[lane][0].data_ptr(),
                               outs[lane][1].data_ptr(), C.c_void_p(wsp[lane]), wsb, C.c_void_p(streams[lane].cuda_stream))
    assert rc == 0, eng.lib.avsep_last_error(eng.h)
def run(n, lanes):
    for i in range(n): fwd(i, i % lanes)
for lanes in (1, 2):
    for s in streams: s.synchronize()
    run(16, lanes); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0]); streams[1].wait_event(e0)
    n = 200
    run(n, lanes)
    ev = torch.cuda.Event(); ev.record(streams[1]); streams[0].wait_event(ev)
    e1.record(streams[0]); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{lanes} forward(s) in flight: {ms:.4f} ms per forward, {B / ms * 1e3:.0f} utt-s/s", flush=True)
