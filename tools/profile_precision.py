import os, sys, torch
sys.path.insert(0, "/root/repo/av-separation-transformer_b200")
from avsep_b200 import AVSeparationTransformer
from avsep_b200.synth import synthetic_batch
for prec in ("tf32",):
    m = AVSeparationTransformer(precision=prec).cuda().eval(); m.prepack("cuda")
    mixed, frames = synthetic_batch(256, device="cuda")
    for _ in range(3): m(mixed, frames)
    eng = m.engine
    eng.set_profile(True); eng.profile_report(reset=True)
    for _ in range(5): m(mixed, frames)
    prof = eng.profile_report(reset=True); eng.set_profile(False)
    tot = sum(v[1] for v in prof.values())
    print(prec, "total per forward ms", tot / 5)
    for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:18s} {n//5:3d} launches {ms/5:8.4f} ms {ms/tot*100:5.1f}%")
