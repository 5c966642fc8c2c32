"""Per-kernel device times of SyntheticAVDataset.batch(256) (torch profiler): STFT, lip frames, waveforms."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from torch.profiler import profile, ProfilerActivity
from avsep_b200.dataset import SyntheticAVDataset
ds = SyntheticAVDataset(num_samples=100000)
for _ in range(3): ds.batch(range(256))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): ds.batch(range(256))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
