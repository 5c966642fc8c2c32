"""Input synthesis and SNR evaluation on the device vs the reference's algorithm on the host (oracle port).

Prints items/s for: host draws only (numpy RNG, unavoidable), device arithmetic (avsep_synth_batch, draws resident),
the whole GPU-backed dataset.batch(), and the oracle's CPU item loop."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200")); sys.path.insert(0, ROOT)
from avsep_b200.dataset import SyntheticAVDataset
from oracle import synth_oracle as so

B = 256
ds = SyntheticAVDataset(num_samples=100000)
idx = list(range(B))
ds.batch(idx); torch.cuda.synchronize()
t0 = time.perf_counter(); d = [ds.draws(i) for i in idx]; t_draw = time.perf_counter() - t0
dev = torch.device("cuda", 0)
up = lambda k, dt: torch.from_numpy(np.stack([x[k] for x in d], 0)).to(dt).to(dev)
amps, freqs, phases, noise = up(0, torch.float64), up(1, torch.float64), up(2, torch.float64), up(3, torch.float32)
# warm up holding the previous outputs, as the timed loop does: the caching allocator needs its second set of output
# buffers (three cudaMallocs, milliseconds each) before the steady state is reached
for _ in range(4): out = ds.engine.synth_batch(ds._geom, amps, freqs, phases, noise, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = ds.engine.synth_batch(ds._geom, amps, freqs, phases, noise, True)
e1.record(); torch.cuda.synchronize()
ms_dev = e0.elapsed_time(e1) / 20
t0 = time.perf_counter(); ds.batch(idx); torch.cuda.synchronize(); t_batch = time.perf_counter() - t0
cfg = so.SynthConfig()
t0 = time.perf_counter()
for i in range(16): so.synth_item(cfg, i)
t_cpu = (time.perf_counter() - t0) / 16
# SNR evaluation
sep = out[2].flip(1).contiguous()
for _ in range(3): ds.engine.eval_snr(sep, out[2], out[0])
torch.cuda.synchronize(); e0.record()
for _ in range(20): ds.engine.eval_snr(sep, out[2], out[0])
e1.record(); torch.cuda.synchronize()
ms_snr = e0.elapsed_time(e1) / 20
sn, tn, mn = sep[:8].cpu().numpy(), out[2][:8].cpu().numpy(), out[0][:8].cpu().numpy()
t0 = time.perf_counter()
for b in range(8): so.permutation_snr(sn[b], tn[b]); so.input_snrs(mn[b], tn[b])
t_snr_cpu = (time.perf_counter() - t0) / 8
print(json.dumps({"batch": B, "host_draws_ms_per_item": t_draw / B * 1e3, "device_synth_ms_per_batch": ms_dev,
                  "device_synth_items_per_s": B / (ms_dev * 1e-3), "dataset_batch_items_per_s": B / t_batch,
                  "cpu_oracle_ms_per_item": t_cpu * 1e3, "cpu_oracle_items_per_s": 1 / t_cpu,
                  "device_eval_snr_ms_per_batch": ms_snr, "cpu_eval_snr_ms_per_item": t_snr_cpu * 1e3}))
