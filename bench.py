#!/usr/bin/env python
"""Benchmark of the hot path: AVSeparationTransformer.forward on SyntheticAVDataset-shaped input.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward pass of the default model (BASELINE.json configs[1]: freq_bins=257, d_model=256, nhead=4,
2 encoder + 2 fusion layers, 2 speakers, bf16 operands) over one batch of B=256 one-second utterances per GPU
(1 s @ 8 kHz: T=63 STFT frames, N=50 lip frames of 32x32).  The batch shards across ranks with no data-path
collective (weak scaling: per-GPU batch fixed); NCCL is used only for the barrier and the max-over-ranks time.
Prints ONE JSON line (rank 0).  Metric: separated utterance-seconds per second = B_total * 1 s / forward time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "av-separation-transformer_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "separated utterance-sec/sec"
UNIT = "utt-s/s"
MODEL = dict(freq_bins=257, d_model=256, nhead=4, num_encoder_layers=2, num_fusion_layers=2, num_speakers=2)
CLIP_SECONDS = 1.0
T_FRAMES, N_FRAMES, FRAME_HW = 63, 50, 32          # dataset.py:63-65,114 at 8 kHz / 1 s / hop 128 / 25 fps x 2 speakers
CPU_SAMPLE_B = 32                                  # CPU throughput is flat beyond B~32


def workload_name(batch):
    return (f"configs[1]: default model (F=257,d=256,H=4,2+2 layers,S=2), B={batch}/GPU, 1 s @ 8 kHz "
            f"(T={T_FRAMES}, N={N_FRAMES}, {FRAME_HW}x{FRAME_HW}), SyntheticAVDataset-shaped")


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]),
                    tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------
# algorithmic work per kernel class, per forward of B utterances (SURVEY.md section 8d formulas)
# ----------------------------------------------------------------------------------------------
def kernel_work(B):
    d, F, S = MODEL["d_model"], MODEL["freq_bins"], MODEL["num_speakers"]
    Le, Lf, T, N = MODEL["num_encoder_layers"], MODEL["num_fusion_layers"], T_FRAMES, N_FRAMES
    Ma, Mv = B * T, B * N
    rows_enc = Le * (Ma + Mv)
    cnn_per_frame = 2 * 256 * 32 * 9 + 2 * 64 * 64 * 288 + 2 * 16 * 128 * 576
    flops = {
        "gemm.conv1d_0": 2 * Ma * d * 3 * F, "gemm.conv1d_2": 2 * Ma * d * 3 * d,
        "gemm.qkv": rows_enc * 2 * 3 * d * d, "gemm.out_proj": (rows_enc + Lf * Ma) * 2 * d * d,
        "gemm.ffn1": (rows_enc + Lf * Ma) * 2 * 4 * d * d, "gemm.ffn2": (rows_enc + Lf * Ma) * 2 * 4 * d * d,
        "gemm.cross_q": Lf * Ma * 2 * d * d, "gemm.cross_kv": Mv * 2 * (Lf * 2 * d) * d,
        "attn.self": Le * B * 4 * d * (T * T + N * N), "attn.cross": Lf * B * 4 * d * T * T,
        "visual_cnn": Mv * cnn_per_frame, "gemm.frame_proj": Mv * 2 * 128 * d,
        "gemm.dec0": Ma * 2 * 2 * d * d, "gemm.dec3_tail": Ma * 2 * S * F * 2 * d,
    }
    flops["ffn.fused"] = flops["gemm.ffn1"] + flops["gemm.ffn2"]     # linear1 + act + linear2 + residual + LN in one kernel
    n_ln = 2 + 2 * Le * 2 + 2 * Lf            # launches per forward
    ln_rows = (1 + 2 * Le) * Ma + (1 + 2 * Le) * Mv + 2 * Lf * Ma
    bytes_ = {
        # x (fp32) + y (fp32) in, x (fp32) + normalised operand (bf16) out = 14 B per element
        "add_layernorm": ln_rows * d * 14,
        # masks + separated fp32 out, mixed fp32 in, bf16 A operand in
        "gemm.dec3_tail": Ma * S * F * 8 + B * F * T * 4 + Ma * 2 * d * 2,
        "prep_audio": B * F * T * 4 + B * (T + 2) * ((F + 7) // 8 * 8) * 2,
    }
    return flops, bytes_, n_ln


def shard_bounds(rank, world, per_rank_batch):
    """[lo, hi) utterance indices of this rank's shard of the global batch (weak scaling: fixed per-rank batch)."""
    return rank * per_rank_batch, (rank + 1) * per_rank_batch


def throughput(world, per_rank_batch, ms_per_step):
    """Whole-job utterance-seconds per second from the slowest rank's step time."""
    return world * per_rank_batch * CLIP_SECONDS / (ms_per_step * 1e-3)


def max_over_ranks_cpu(x):
    """MAX all-reduce of a host scalar on the default process group (gloo in tests)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total_flops(B):
    return sum(v for k, v in kernel_work(B)[0].items() if k != "ffn.fused")   # ffn.fused = ffn1 + ffn2, counted once


# ----------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle port) -- the ONLY place bench.py touches oracle/
# ----------------------------------------------------------------------------------------------
def cpu_reference_throughput(state, mixed, frames, min_seconds=10.0, max_reps=8):
    import torch
    from oracle import avsep_oracle_torch as otorch
    from oracle.weights import ModelConfig
    cfg = ModelConfig(**MODEL)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = {k: v.detach().cpu() for k, v in state.items()}
    otorch.forward(P, cfg, mixed, frames)            # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_reps and (time.perf_counter() - t_all < min_seconds or len(times) < 3):
        t0 = time.perf_counter()
        otorch.forward(P, cfg, mixed, frames)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return mixed.shape[0] * CLIP_SECONDS / best, cores, best, len(times)


def build_state(seed=0):
    """Random-init weights of the architecture (torch default init, as the reference constructs them), with
    non-trivial BatchNorm statistics.  Uses only torch.nn parameter containers -- no kernels involved."""
    import torch
    from avsep_b200 import AVSeparationTransformer
    torch.manual_seed(seed)
    model = AVSeparationTransformer(**MODEL, precision="bf16")
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    return model


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU forward (oracle port on torch CPU ops, all host threads)."""
    if rank != 0:
        return
    import torch
    from avsep_b200.synth import synthetic_batch
    model = build_state()
    state = model.state_dict()
    mixed, frames = synthetic_batch(CPU_SAMPLE_B, MODEL["freq_bins"], T_FRAMES, N_FRAMES, FRAME_HW, FRAME_HW, seed=100)
    from oracle import avsep_oracle_torch as otorch
    from oracle.weights import ModelConfig
    cfg = ModelConfig(**MODEL)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = {k: v.detach().cpu() for k, v in state.items()}
    for _ in range(max(1, min(args.warmup, 3))):
        otorch.forward(P, cfg, mixed, frames)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        otorch.forward(P, cfg, mixed, frames)
    dt = time.perf_counter() - t0
    value = args.steps * CPU_SAMPLE_B * CLIP_SECONDS / dt
    sample = f"B={CPU_SAMPLE_B} utterances of the same workload per step (CPU throughput is flat beyond B~32)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads (and therefore its first-touch pinned host buffers) to the CPU cores local to its GPU,
    so that the e2e host<->device copies of 8 ranks do not all cross one socket.  Returns the core count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as e:                                   # affinity is an optimisation, never a requirement
        sys.stderr.write(f"bench.py: NUMA binding skipped ({e})\n")
    return None


def run_ours(args, rank, local_rank, world):
    # Everything except the final JSON line goes to stderr (NCCL / torchrun print to stdout on their own).
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from avsep_b200.synth import synthetic_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    F, S = MODEL["freq_bins"], MODEL["num_speakers"]
    model = build_state().to(dev)
    model.prepack(dev)
    eng = model.engine
    # rotating input sets: inputs (69 MB) + outputs (133 MB) + activations (~350 MB) per step already exceed the
    # 126 MB L2; three distinct input sets make sure no step re-reads the previous step's inputs from L2.
    n_sets = 3
    lo, hi = shard_bounds(rank, world, B)     # this rank's utterances of the global batch; seeds differ per shard
    sets = [synthetic_batch(hi - lo, F, T_FRAMES, N_FRAMES, FRAME_HW, FRAME_HW, seed=100003 * i + lo, device=dev)
            for i in range(n_sets)]
    sep = torch.empty((B, S, F, T_FRAMES), device=dev)
    masks = torch.empty_like(sep)
    import ctypes as C
    stream = torch.cuda.current_stream()

    def step(i):
        mixed, frames = sets[i % n_sets]
        rc = eng.lib.avsep_forward(eng.h, mixed.data_ptr(), frames.data_ptr(), B, T_FRAMES, N_FRAMES, FRAME_HW,
                                   FRAME_HW, sep.data_ptr(), masks.data_ptr(), None, 0,
                                   C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(eng.lib.avsep_last_error(eng.h).decode())

    for i in range(args.warmup):
        step(i)
    launches_per_step = eng.launch_count()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if rank == 0 else 0)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = throughput(world, B, ms_per_step)

    # ---- e2e: the public host-buffer call; pinned inputs H2D and results D2H every step -----------------
    h_sets = [(m.cpu().pin_memory(), f.cpu().pin_memory()) for m, f in sets[:2]]
    h_out = [(torch.empty((B, S, F, T_FRAMES)).pin_memory(), torch.empty((B, S, F, T_FRAMES)).pin_memory())
             for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 20))

    def e2e_run(n):
        """n batches through the streaming host entry point: batch i goes to I/O slot i % 2, and its results are
        waited for (= are in the pinned host buffers) before that slot is submitted again."""
        for i in range(n):
            slot = i % 2
            if i >= 2:
                eng.host_wait(slot)
            eng.forward_host_async(h_sets[slot][0], h_sets[slot][1], h_out[slot][0], h_out[slot][1], slot)
        eng.host_wait(0)
        eng.host_wait(1)

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # the synchronous call (one batch at a time, results on the host when it returns), for reference
    for _ in range(2):
        eng.forward_host(h_sets[0][0], h_sets[0][1], h_out[0][0], h_out[0][1])
    t0 = time.perf_counter()
    for i in range(4):
        eng.forward_host(h_sets[i % 2][0], h_sets[i % 2][1], h_out[0][0], h_out[0][1])
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h_sep = h_out[0][0]
    h_masks = h_out[0][1]
    e2e_stream_value = world * B * CLIP_SECONDS * e2e_steps / e2e_s
    e2e_sync_value = world * B * CLIP_SECONDS * 4 / e2e_sync_s
    # both are the repo's public host-buffer API; report the faster form (on one GPU the streaming form is PCIe-bound
    # and ~30 % ahead; with 8 ranks sharing one host the copies of all ranks contend and the forms are close)
    e2e_value = max(e2e_stream_value, e2e_sync_value)
    h2d = (h_sets[0][0].numel() + h_sets[0][1].numel()) * 4
    d2h = (h_sep.numel() + h_masks.numel()) * 4

    # ---- N > 1: the same steps with the outputs collected on rank 0 (the headline keeps every rank's
    # outputs on its own GPU; a consumer that wants them in one place pays an NCCL gather over NVLink).  Two wire
    # formats: masks + separated in fp32 (what the reference returns), and masks only in bf16 (rank 0 holds the
    # mixture and can rebuild `separated = masks * mixed`); each measured back to back and with the gather of step i
    # overlapped with the kernels of step i+1 (two output buffer sets).
    gathered = None
    if world > 1:
        gathered = {}
        sep2, masks2 = torch.empty_like(sep), torch.empty_like(masks)
        out_sets = [(sep, masks), (sep2, masks2)]

        def step_into(i, o):
            mixed, frames = sets[i % n_sets]
            rc = eng.lib.avsep_forward(eng.h, mixed.data_ptr(), frames.data_ptr(), B, T_FRAMES, N_FRAMES, FRAME_HW,
                                       FRAME_HW, out_sets[o][0].data_ptr(), out_sets[o][1].data_ptr(), None, 0,
                                       C.c_void_p(stream.cuda_stream))
            if rc != 0:
                raise RuntimeError(eng.lib.avsep_last_error(eng.h).decode())

        wires = (("masks+separated fp32", lambda o: (out_sets[o][1], out_sets[o][0])),
                 ("masks bf16", lambda o: (out_sets[o][1].to(torch.bfloat16),)))
        for tag, pick in wires:
            for overlapped in (False, True):
                # destination lists on rank 0, one per output buffer set
                dst = [[[torch.empty_like(t) for _ in range(world)] if rank == 0 else None for t in pick(o)] for o in (0, 1)]
                pending = [None, None]

                def gstep(i):
                    o = i & 1 if overlapped else 0
                    if pending[o] is not None:                 # the gather that last read this buffer set
                        for w in pending[o]:
                            w.wait()
                    step_into(i, o)
                    works = [dist.gather(t, gl, dst=0, async_op=True) for t, gl in zip(pick(o), dst[o])]
                    if overlapped:
                        pending[o] = works                     # the next step (other buffer set) runs under it
                    else:
                        for w in works:
                            w.wait()

                def drain():
                    for o in (0, 1):
                        if pending[o] is not None:
                            for w in pending[o]:
                                w.wait()
                            pending[o] = None

                for i in range(4):
                    gstep(i)
                drain()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n_g = max(6, min(args.steps, 50))
                g0.record(stream)
                for i in range(n_g):
                    gstep(i)
                drain()
                g1.record(stream)
                barrier()
                g_ms = max_over_ranks(g0.elapsed_time(g1)) / n_g
                gathered[tag + (", overlapped with the next step" if overlapped else "")] = {
                    "ms_per_step": round(g_ms, 4), "value": throughput(world, B, g_ms),
                    "bytes_into_rank0_per_step": sum(t.numel() * t.element_size() for t in pick(0)) * (world - 1)}
                del dst
        del sep2, masks2

    # ---- per-kernel pass (same steps again, every launch bracketed by CUDA events on the launching stream) ----
    eng.set_profile(True)
    eng.profile_report(reset=True)
    prof_steps = max(2, min(args.steps, 10))
    for i in range(prof_steps):
        step(i)
    prof = eng.profile_report(reset=True)
    eng.set_profile(False)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    flops, bytes_, _ = kernel_work(B)
    tot_ms = sum(v[1] for v in prof.values()) or 1.0
    breakdown = {}
    for label, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per_step_ms = ms / prof_steps
        entry = {"launches_per_step": n // prof_steps, "ms_per_step": round(per_step_ms, 4),
                 "share": round(ms / tot_ms, 4)}
        if label in flops and label != "gemm.dec3_tail":
            entry["tflops"] = round(flops[label] / (per_step_ms * 1e-3) / 1e12, 2)
            entry["frac_of_tensor_peak"] = round(entry["tflops"] / peaks["tensor_sustained"], 4)
        if label in bytes_:
            entry["gbs"] = round(bytes_[label] / (per_step_ms * 1e-3) / 1e9, 1)
            entry["frac_of_hbm_peak"] = round(entry["gbs"] / peaks["hbm"], 4)
        breakdown[label] = entry
    top = next(iter(breakdown))
    te = breakdown[top]
    traffic = None      # dram bytes per launch of the dominant kernel from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and B == 256:
        with open(tpath) as f:
            traffic = json.load(f).get(top, {}).get("traffic")
    n_top = max(1, te["launches_per_step"])
    if "tflops" in te:
        roofline = {"kernel": top, "bound": "tensor", "achieved": te["tflops"], "peak": peaks["tensor_sustained"],
                    "unit": "TFLOP/s", "frac": round(te["tflops"] / peaks["tensor_sustained"], 4), "traffic": traffic,
                    "peak_source": peaks["source"] + " (sustained bf16: kernel timed inside a long step)",
                    "launch_ms": round(te["ms_per_step"] / n_top, 4),
                    "algorithmic_flops_per_launch": flops[top] / n_top}
    else:
        roofline = {"kernel": top, "bound": "hbm", "achieved": te["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(te["gbs"] / peaks["hbm"], 4), "traffic": traffic, "peak_source": peaks["source"],
                    "launch_ms": round(te["ms_per_step"] / n_top, 4),
                    "algorithmic_bytes_per_launch": bytes_[top] / n_top}

    # Secondary limiter of the GEMM-class kernels (tools/tma_rate.cu): a B200 SM ingests TMA boxes from L2 at
    # 41.4 B/clk = 81 GB/s whatever the number of boxes in flight or of SMs streaming (12 TB/s chip-wide).  The fused
    # FFN re-streams W1 + W2 (1 MB) for every 128-row tile, plus the A tile (64 KB) and the fp32 residual (128 KB).
    if top == "ffn.fused":
        d = MODEL["d_model"]
        tiles = sum((m + 127) // 128 for m in ([B * T_FRAMES] * (MODEL["num_encoder_layers"] + MODEL["num_fusion_layers"])
                                                + [B * N_FRAMES] * MODEL["num_encoder_layers"]))
        per_tile = 2 * 4 * d * d * 2 + 128 * d * 2 + 128 * d * 4
        sms_busy = min(148, max(1, tiles // n_top))
        gbs_per_sm = tiles * per_tile / (te["ms_per_step"] * 1e-3) / 1e9 / sms_busy
        roofline["operand_stream"] = {"what": "TMA ingest L2 -> shared memory per SM, whole kernel incl. prologue and epilogue",
                                      "achieved_gb_s_per_sm": round(gbs_per_sm, 1), "peak_gb_s_per_sm": 81.0,
                                      "frac": round(gbs_per_sm / 81.0, 3), "peak_source": "tools/tma_rate.cu on this pool's B200"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(B), "global_batch": world * B, "parallelism": f"batch-sharded x{world}", "rank_cpu_affinity_cores": numa_cores,
                   "l2": f"{n_sets} rotating input sets; per-step footprint (inputs+outputs+activations) > 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "api": ("avsep_forward_host_async on two I/O slots + avsep_host_wait" if e2e_stream_value >= e2e_sync_value
                        else "avsep_forward_host (one call per batch)") +
                       " (pinned host buffers; every step's H2D and D2H inside the timed region)",
                "streaming_value": e2e_stream_value, "sync_call_value": e2e_sync_value},
        "gpu_launches": launches_per_step * args.steps,
        "gflop_per_step": total_flops(B) / 1e9,
        "model_tflops": round(total_flops(B) / (ms_per_step * 1e-3) / 1e12, 2),
        "roofline": roofline,
        "kernels": breakdown,
    }
    if gathered is not None:
        out["outputs_gathered_to_rank0"] = gathered
    if world == 1:
        # The same workload waveform to waveform: STFT of the 1 s mixture, forward on its magnitude,
        # masks applied to the complex mixture, inverse STFT -- three C-ABI calls per step, buffers preallocated.
        n_fft, hop, L = 2 * (F - 1), 128, int(8000 * CLIP_SECONDS)
        g = torch.Generator(device=dev).manual_seed(7)
        waves_in = [0.3 * torch.randn(B, L, device=dev, generator=g) for _ in range(n_sets)]
        spec = torch.empty(B, F, T_FRAMES, device=dev, dtype=torch.complex64)
        mag = torch.empty(B, F, T_FRAMES, device=dev)
        waves_out = torch.empty(B, S, L, device=dev)
        st = C.c_void_p(stream.cuda_stream)

        def wstep(i):
            frames = sets[i % n_sets][1]
            rc = eng.lib.avsep_stft(eng.h, waves_in[i % n_sets].data_ptr(), B, L, n_fft, hop, spec.data_ptr(),
                                    mag.data_ptr(), st)
            rc = rc or eng.lib.avsep_forward(eng.h, mag.data_ptr(), frames.data_ptr(), B, T_FRAMES, N_FRAMES, FRAME_HW,
                                             FRAME_HW, sep.data_ptr(), masks.data_ptr(), None, 0, st)
            rc = rc or eng.lib.avsep_istft(eng.h, spec.data_ptr(), masks.data_ptr(), B, S, T_FRAMES, n_fft, hop, L,
                                           waves_out.data_ptr(), st)
            if rc != 0:
                raise RuntimeError(eng.lib.avsep_last_error(eng.h).decode())

        w_steps = max(4, min(args.steps, 50))
        for i in range(3):
            wstep(i)
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(stream)
        for i in range(w_steps):
            wstep(i)
        w1.record(stream)
        torch.cuda.synchronize()
        w_ms = w0.elapsed_time(w1) / w_steps
        out["waveform_to_waveform"] = {"ms_per_step": round(w_ms, 4), "value": throughput(1, B, w_ms), "unit": UNIT,
                                       "steps": w_steps, "samples_per_utterance": L,
                                       "calls": "avsep_stft -> avsep_forward -> avsep_istft, inputs resident"}
    if world == 1 and not args.no_cpu_baseline:
        m_cpu, f_cpu = sets[0][0][:CPU_SAMPLE_B].cpu(), sets[0][1][:CPU_SAMPLE_B].cpu()
        v, cores, best, reps = cpu_reference_throughput(model.state_dict(), m_cpu, f_cpu)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"first {CPU_SAMPLE_B} utterances of the same batch, best of {reps} forwards "
                                         f"({best * 1e3:.0f} ms each), torch CPU ops, {cores} threads"}
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--batch", str(args.batch)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
