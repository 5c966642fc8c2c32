"""Print the interesting parts of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        e2e = d.get("e2e", {})
        print(f"{f}: {d.get('ms_per_step'):.4f} ms/step  value={d.get('value'):.0f}  e2e={e2e.get('value', 0):.0f}  n={d.get('n_gpus')}  launches={d.get('gpu_launches')}")
        for k, v in (d.get("kernels") or {}).items():
            fr = v.get("frac_of_tensor_peak", v.get("frac_of_hbm_peak", ""))
            print(f"    {k:24s} {v['launches_per_step']:3d} x  {v['ms_per_step']*1000:7.1f} us  share {v['share']:.3f}  frac {fr}")
        for k in ("python_api", "per_gpu_batch_sweep", "multi_gpu"):
            if k in d:
                print("   ", k, d[k])
