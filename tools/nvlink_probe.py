"""Root-centred NVLink copy rates with the library's peer-memory ABI (avsep_shared_alloc / avsep_copy_async), one
process per GPU: every rank r > 0 pulls `in_mb` from rank 0's buffer and / or pushes `out_mb` into it, all at once,
with 1..4 copy streams per rank.  Names the ceiling of the scatter -> forward -> gather headline.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nvlink_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from avsep_b200 import AVSeparationTransformer  # noqa: E402
from avsep_b200.sharded import PeerMemoryCuda  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
model = AVSeparationTransformer().cuda()
model.prepack()
mem = PeerMemoryCuda(model.engine)
IN_F, OUT_F = 69008384 // 4, 66318336 // 4           # floats per rank per step (B=256 inputs / outputs)
box = [None]
if rank == 0:
    pin, hin = mem.alloc(IN_F * world)
    pout, hout = mem.alloc(OUT_F * world)
    box = [(hin, hout)]
dist.broadcast_object_list(box, src=0)
if rank != 0:
    pin, pout = mem.open(box[0][0]), mem.open(box[0][1])
loc_in = torch.empty(IN_F, device="cuda")
loc_out = torch.ones(OUT_F, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]


def run(pull, push, nstreams, iters=20):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if rank != 0:
        for s in streams[:nstreams]:
            s.wait_stream(torch.cuda.current_stream())
        for _ in range(iters):
            for k in range(nstreams):
                s = streams[k]
                if pull:
                    n = IN_F // nstreams
                    mem.copy(loc_in.data_ptr() + 4 * k * n, pin + 4 * (rank * IN_F + k * n), n, s)
                if push:
                    n = OUT_F // nstreams
                    mem.copy(pout + 4 * (rank * OUT_F + k * n), loc_out.data_ptr() + 4 * k * n, n, s)
        for s in streams[:nstreams]:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    ms = float(t) / iters
    if rank == 0:
        gb_out = (world - 1) * IN_F * 4 / 1e9 if pull else 0.0
        gb_in = (world - 1) * OUT_F * 4 / 1e9 if push else 0.0
        print(f"world={world} pull={int(pull)} push={int(push)} streams/rank={nstreams}: {ms:.3f} ms per round; "
              f"root egress {gb_out / ms * 1e3:.0f} GB/s, root ingress {gb_in / ms * 1e3:.0f} GB/s", flush=True)


for ns in (1, 2, 4):
    run(True, False, ns)
    run(False, True, ns)
    run(True, True, ns)
dist.destroy_process_group()
