"""ShardedForward on real peer memory (needs >= 2 GPUs; skipped on a one-GPU box): scatter -> forward -> gather with
both tensors on the wire, and with masks only (tickets + avsep_separate on the root), must leave the same bytes in the
root's buffers as the root's own forward of the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle.weights import CONFIGS, make_inputs, make_state_dict
        from tests.helpers import build_model
        from avsep_b200.sharded import PeerMemoryCuda, ShardedForward
        cfg = CONFIGS["default"]
        B, T, N, HW = 6, 63, 50, 32
        model = build_model(cfg, make_state_dict(cfg, seed=81, gain=2.0), "bf16", device=dev)
        model.prepack(dev)
        eng = model.engine
        shapes = dict(mixed=(cfg.freq_bins, T), frames=(N, HW, HW), out=(cfg.num_speakers, cfg.freq_bins, T))

        def fwd(mixed, frames, sep, masks):
            eng.forward(mixed, frames, out=(sep, masks))

        results = {}
        for gather in ("both", "masks"):
            sf = ShardedForward(PeerMemoryCuda(eng), fwd, B, shapes, rank, world, n_input_sets=3, gather=gather)
            if rank == 0:
                for s_i, (gm, gf) in enumerate(sf.root_in):
                    m, f = make_inputs(cfg, world * B, T, N, HW, HW, seed=90 + s_i, kind="dataset")
                    gm.copy_(torch.from_numpy(m))
                    gf.copy_(torch.from_numpy(f))
                for o_s, o_m in sf.root_out:
                    o_s.fill_(float("nan"))
                    o_m.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            n_steps = 7
            for i in range(n_steps):
                sf.step(i)
            sf.finish()
            if rank == 0:
                ok = True
                for i in (n_steps - 2, n_steps - 1):
                    gm, gf = sf.root_in[i % 3]
                    want_sep, want_masks = eng.forward(gm, gf)
                    got_sep, got_masks = sf.root_out[i & 1]
                    ok = ok and bool(torch.equal(got_sep, want_sep)) and bool(torch.equal(got_masks, want_masks))
                results[gather] = (ok, sf.bytes_out_per_step)
            dist.barrier()
        if rank == 0:
            out.update(results)
    finally:
        dist.destroy_process_group()


def test_scatter_forward_gather_two_gpus_both_wire_formats():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert out["both"][0] is True and out["masks"][0] is True, dict(out)
        assert out["masks"][1] * 2 == out["both"][1]
