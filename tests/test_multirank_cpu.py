"""CPU, world_size 2, gloo: the N>1 logic of bench.py -- per-rank shards of the utterance batch (no data-path
collective), barrier, max-over-ranks of the step time, whole-job throughput from the slowest rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank owns a disjoint, contiguous shard of the global batch (weak scaling: per-rank batch fixed)
        per_rank = 4
        lo, hi = bench.shard_bounds(rank, world, per_rank)
        assert hi - lo == per_rank and lo == rank * per_rank
        ids = torch.arange(lo, hi)
        gathered = [torch.zeros(per_rank, dtype=torch.long) for _ in range(world)]
        dist.all_gather(gathered, ids)
        assert torch.cat(gathered).tolist() == list(range(world * per_rank))      # shards tile the global batch
        # slowest rank defines the step time
        ms = bench.max_over_ranks_cpu(10.0 + 5.0 * rank)
        assert ms == 10.0 + 5.0 * (world - 1)
        value = bench.throughput(world, per_rank, ms)
        assert abs(value - world * per_rank / (ms * 1e-3)) < 1e-9
        dist.barrier()
        out[rank] = value
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == 2 and out[0] == out[1]


def test_reference_arm_only_runs_on_rank0(monkeypatch, capsys):
    class A:
        steps, warmup, gpus, batch = 1, 1, 2, 256
    bench.run_reference(A, rank=1, world=2)
    assert capsys.readouterr().out == ""
