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


# ---- the scatter -> forward -> gather bookkeeping of avsep_b200.sharded, world_size 2, gloo -------------
class _HostBackend:
    """CPU stand-in for PeerMemoryCuda: POSIX shared memory instead of cudaIpc handles, memmove instead of copy engines,
    no-op streams / events (everything is synchronous)."""

    class _Obj:
        def wait_event(self, ev): pass
        def record(self, stream=None): pass

    def __init__(self):
        self.shms = []

    def _addr(self, shm):
        import ctypes
        return ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))

    def alloc(self, nfloats):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=nfloats * 4)
        self.shms.append(shm)
        return self._addr(shm), shm.name.encode()

    def open(self, handle):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(name=handle.decode())
        self.shms.append(shm)
        return self._addr(shm)

    def view(self, ptr, shape):
        import ctypes
        n = 1
        for s in shape:
            n *= s
        buf = (ctypes.c_float * n).from_address(ptr)
        return torch.frombuffer(buf, dtype=torch.float32).view(*shape)

    def empty(self, shape):
        return torch.empty(*shape, dtype=torch.float32)

    def copy(self, dst, src, nfloats, stream):
        import ctypes
        ctypes.memmove(dst, src, nfloats * 4)

    # tickets of the masks-only gather: 32-bit words in shared memory, polled by the waiting process
    def zero(self, ptr, nfloats):
        import ctypes
        ctypes.memset(ptr, 0, nfloats * 4)

    def signal(self, flag_ptrs, value, stream):
        import ctypes
        for p in flag_ptrs:
            ctypes.c_uint32.from_address(p).value = value

    def wait(self, flag_ptrs, value, stream, timeout_s=30.0):
        import ctypes
        import time
        t0 = time.time()
        for p in flag_ptrs:
            while ctypes.c_uint32.from_address(p).value < value:
                time.sleep(0.0005)
                assert time.time() - t0 < timeout_s, "ticket did not arrive"
        self.waits = getattr(self, "waits", 0) + 1

    def separate(self, masks_ptr, mixed_ptr, n_utt, T, sep_ptr, stream):
        S, F = self.out_shape[0], self.out_shape[1]
        masks = self.view(masks_ptr, (n_utt, S, F, T))
        mixed = self.view(mixed_ptr, (n_utt, F, T))
        self.view(sep_ptr, (n_utt, S, F, T)).copy_(masks * mixed.unsqueeze(1))

    def stream(self): return self._Obj()
    def current_stream(self): return self._Obj()
    def event(self): return self._Obj()
    def synchronize(self): pass

    def close(self, unlink):
        for shm in self.shms:
            try:
                shm.close()
                if unlink:
                    shm.unlink()
            except Exception:
                pass


def _fake_forward(mixed, frames, sep, masks):
    # any per-utterance function: an output row depends on its own utterance only (like the model in eval mode)
    S = sep.shape[1]
    w = frames.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
    for s in range(S):
        masks[:, s] = torch.sigmoid(mixed * (s + 1) + w[:, 0])
        sep[:, s] = masks[:, s] * mixed


def _sharded_worker(rank, world, port, out, lanes=1, gather="both", n_steps=5):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avsep_b200.sharded import ShardedForward, shard_slices
    be = _HostBackend()
    try:
        B, shapes = 3, dict(mixed=(5, 7), frames=(4, 2, 2), out=(2, 5, 7))
        assert shard_slices(world, B) == [(r * 3, r * 3 + 3) for r in range(world)]
        be.out_shape = shapes["out"]
        sf = ShardedForward(be, _fake_forward, B, shapes, rank, world, n_input_sets=3, copy_lanes=lanes, gather=gather)
        if rank == 0:
            g = torch.Generator().manual_seed(0)
            for m, f in sf.root_in:
                m.copy_(torch.randn(m.shape, generator=g))
                f.copy_(torch.rand(f.shape, generator=g))
        dist.barrier()
        for i in range(n_steps):
            sf.step(i)
        sf.finish()
        if rank == 0:
            ok = True
            for i in (n_steps - 2, n_steps - 1):         # the two steps still held by the double-buffered outputs
                m, f = sf.root_in[i % 3]
                sep_ref, masks_ref = torch.empty(world * B, 2, 5, 7), torch.empty(world * B, 2, 5, 7)
                _fake_forward(m, f, sep_ref, masks_ref)
                sep, masks = sf.root_out[i & 1]
                # (vectorised sigmoid may differ in the last bit between the sharded and the whole-batch call)
                err = max(float((sep - sep_ref).abs().max()), float((masks - masks_ref).abs().max()))
                ok = ok and err < 1e-6
                out["err"] = err
            out["ok"] = ok
            out["bytes"] = (sf.bytes_in_per_step, sf.bytes_out_per_step)
            out["root_waits"] = getattr(be, "waits", 0)
        else:
            out["peer_waits"] = getattr(be, "waits", 0)
        dist.barrier()
    finally:
        be.close(unlink=(rank == 0))
        dist.destroy_process_group()


def test_sharded_forward_scatter_gather_world_two_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
        assert out["ok"] is True, dict(out)
        assert out["bytes"] == (4 * 3 * (35 + 16), 4 * 3 * 2 * 70)


def test_sharded_forward_with_three_copy_lanes_world_two_gloo():
    """Every transfer cut into three pieces on three streams (the 8-GPU configuration): the pieces tile each shard."""
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_sharded_worker, args=(world, port, out, 3), nprocs=world, join=True)
        assert out["ok"] is True, dict(out)


def test_sharded_forward_masks_only_gather_world_two_gloo():
    """gather='masks': only the masks shard crosses to the root, which rebuilds `separated` after the rank's ticket
    arrived and acknowledges; the root's buffers end up identical to the two-tensor gather."""
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_sharded_worker, args=(world, port, out, 1, "masks"), nprocs=world, join=True)
        assert out["ok"] is True, dict(out)
        assert out["bytes"] == (4 * 3 * (35 + 16), 4 * 3 * 70)          # half of the bytes into the root
        assert out["root_waits"] == 5          # one ticket wait per step on the root ...
        assert out["peer_waits"] == 3          # ... and one acknowledge wait per reuse of an output slot (steps 2, 3, 4)


def test_sharded_forward_masks_only_gather_world_three_gloo():
    """Two remote ranks: the root's rebuild waits for both tickets, rebuilds both shards in one pass and acknowledges
    both; eight steps walk every (input set, output slot) pair and reuse each output slot three times."""
    world, port = 3, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_sharded_worker, args=(world, port, out, 2, "masks", 8), nprocs=world, join=True)
        assert out["ok"] is True, dict(out)
        assert out["root_waits"] == 8 and out["peer_waits"] == 6
