"""Where does a scatter -> forward -> gather step (avsep_b200/sharded.py, as bench.py times it) spend its time?
One process per GPU; times the step with parts of it switched off and prints every rank's device time.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_probe.py [steps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
import bench  # noqa: E402
from avsep_b200.sharded import PeerMemoryCuda, ShardedForward  # noqa: E402
from avsep_b200.synth import synthetic_batch  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("PROBE_BIND", "1") == "1":
    bench.bind_to_gpu_numa_node(local)
dist.init_process_group("nccl", device_id=dev)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B, F, S = 256, bench.MODEL["freq_bins"], bench.MODEL["num_speakers"]
T, N, HW = bench.T_FRAMES, bench.N_FRAMES, bench.FRAME_HW
model = bench.build_state().to(dev)
model.prepack(dev)
eng = model.engine
stream = torch.cuda.current_stream()


def _raw(mixed, frames, sep, masks):
    import ctypes as C
    rc = eng.lib.avsep_forward(eng.h, mixed.data_ptr(), frames.data_ptr(), mixed.shape[0], T, N, HW, HW, sep.data_ptr(),
                               masks.data_ptr(), None, 0, C.c_void_p(stream.cuda_stream))
    if rc != 0:
        raise RuntimeError(eng.lib.avsep_last_error(eng.h).decode())


class Probe(ShardedForward):
    skip = ()          # which of (0 = scatter, 1 = gather) to leave out

    def _copy(self, lead, extra, which, jobs):
        if which in self.skip:
            return
        super()._copy(lead, extra, which, jobs)


shapes = dict(mixed=(F, T), frames=(N, HW, HW), out=(S, F, T))


WARM = int(os.environ.get("PROBE_WARMUP", "6"))


def run(label, fwd_fn, skip=(), lanes=1, gather="both"):
    sg = Probe(PeerMemoryCuda(eng), fwd_fn, B, shapes, rank, world, n_input_sets=3, copy_lanes=lanes, gather=gather)
    sg.skip = skip
    if rank == 0:
        for gm, gf in sg.root_in:
            m_r, f_r = synthetic_batch(B, F, T, N, HW, HW, seed=1, device=dev)
            for r in range(world):
                gm[r * B:(r + 1) * B].copy_(m_r)
                gf[r * B:(r + 1) * B].copy_(f_r)
    for i in range(WARM):
        sg.step(i)
    sg.finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS)]
    e0.record(stream)
    for i in range(STEPS):
        sg.step(i)
        marks[i].record(stream)
    if rank != 0:
        stream.wait_stream(sg.s_in)
        stream.wait_stream(sg.s_out)
    elif hasattr(sg, "s_rb"):
        stream.wait_stream(sg.s_rb)
    e1.record(stream)
    sg.finish()
    per = sorted(a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks))      # forward-to-forward intervals
    t = torch.tensor([e0.elapsed_time(e1) / STEPS, per[len(per) // 2], per[-1]], device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        ms = [round(float(x[0]), 3) for x in out]
        med = [round(float(x[1]), 3) for x in out]
        worst = [round(float(x[2]), 3) for x in out]
        print(f"world={world} {label}: per-rank ms/step {ms} (median step {med}, slowest step {worst}); "
              f"{world * B / max(ms) :.0f} k utt-s/s", flush=True)
    import ctypes as C
    for ptr in sg.ptr.values():       # peers unmap before the root frees
        if rank != 0:
            eng.lib.avsep_shared_close(eng.h, C.c_void_p(ptr))
    torch.cuda.synchronize()
    dist.barrier()
    for ptr in sg.ptr.values():
        if rank == 0:
            eng.lib.avsep_shared_free(eng.h, C.c_void_p(ptr))
    del sg
    torch.cuda.synchronize()
    dist.barrier()


noop = lambda *a: None  # noqa: E731
ONLY = os.environ.get("PROBE_ONLY")            # e.g. "full,masks": run only the labels containing one of these words
_run = run


def run(label, *a, **k):  # noqa: F811
    if ONLY is None or any(w in label for w in ONLY.split(",")):
        _run(label, *a, **k)


for _ in range(int(os.environ.get("PROBE_REPEAT", "1"))):
    run("full", _raw)
    run("copies only (no forward)", noop)
    run("scatter + forward", _raw, skip=(1,))
    run("forward + gather", _raw, skip=(0,))
    run("forward only", _raw, skip=(0, 1))
    run("full, 2 copy lanes", _raw, lanes=2)
    run("full (again)", _raw)
    run("full, masks only on the wire", _raw, gather="masks")
    run("copies only, masks only on the wire", noop, gather="masks")
dist.destroy_process_group()
