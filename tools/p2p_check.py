"""Peer-access matrix, NVLink topology and raw device-to-device copy rates of the visible GPUs (single process)."""
import subprocess, torch
n = torch.cuda.device_count()
print("devices", n)
print("canAccessPeer:", [[int(torch.cuda.can_device_access_peer(i, j)) if i != j else 1 for j in range(n)] for i in range(n)])
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout[:3000])
except Exception as e:
    print("topo failed", e)
x = [torch.empty(64 << 20, device=f"cuda:{i}", dtype=torch.uint8) for i in range(n)]
for j in range(1, n):
    for _ in range(2):
        x[j].copy_(x[0]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.set_device(j)
    e0.record()
    for _ in range(10):
        x[j].copy_(x[0], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"0 -> {j}: {10 * 64 / 1024 / (e0.elapsed_time(e1) / 1e3):.1f} GB/s")
