"""Print the golden-case parity numbers (max |error| of masks / separated) for a precision path."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))
from helpers import load_golden, case_tensors, build_model, err_report
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for name in ("tiny_up", "tiny_down", "c1_dataset", "c1_randn", "c3_long", "c4_scaled"):
    z, meta = load_golden(name)
    cfg, P, mixed, frames = case_tensors(meta)
    model = build_model(cfg, P, prec)
    with torch.no_grad():
        sep, masks = model(torch.from_numpy(mixed).cuda(), torch.from_numpy(frames).cuda())
    sep, masks = sep.cpu().numpy(), masks.cpu().numpy()
    sf, st = meta["stride_f"], meta["stride_t"]
    sep, masks, mixed = sep[:, :, ::sf, ::st], masks[:, :, ::sf, ::st], mixed[:, ::sf, ::st]
    print(prec, name, json.dumps({k: float(f"{v:.3e}") for k, v in err_report(sep, masks, z["separated"], z["masks"], mixed).items()}), flush=True)
