#!/usr/bin/env python
"""Phase timeline of the fused transformer-stack kernel (csrc/xformer_stack_sm100.cu) from its clock64 stamps.

    python tools/stack_trace.py [--which 0|1|2] [--batch B] [--len L]

Prints, for the median CTA, the clock deltas between consecutive stamps of the first row thread and of the MMA thread
(first tile of the CTA), per layer.  Debug / measurement aid; not on the product path."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av-separation-transformer_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", type=int, default=0)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--len", type=int, default=63)
    ap.add_argument("--decoder", action="store_true", help="which=2 with the SeparationDecoder fused behind the stack")
    args = ap.parse_args()
    from avsep_b200 import AVSeparationTransformer
    torch.manual_seed(0)
    model = AVSeparationTransformer().cuda()
    model.prepack()
    eng = model.engine
    B, L, d = args.batch, args.len, 256
    x = torch.randn(B * L, d, device="cuda")
    kv = torch.randn(B * L, 1024, device="cuda").to(torch.bfloat16) if args.which == 2 else None
    out = torch.empty(B * L, d, device="cuda", dtype=torch.bfloat16)
    U = 1
    while U < 16 and 2 * U * L <= 128:
        U *= 2
    grid = min(148, (B + U - 1) // U)
    trace = torch.zeros(grid * 256, device="cuda", dtype=torch.int64)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if args.decoder:
        assert args.which == 2
        mixed = torch.rand(B, 257, L, device="cuda")
        sep = torch.empty(B, 2, 257, L, device="cuda")
        masks = torch.empty_like(sep)

    def launch(tr):
        if args.decoder:
            return eng.lib.avsep_test_fusion_decoder(eng.h, x.data_ptr(), kv.data_ptr(), B, L, mixed.data_ptr(), sep.data_ptr(),
                                                     masks.data_ptr(), tr, st)
        return eng.lib.avsep_test_xformer_stack(eng.h, args.which, x.data_ptr(), kv.data_ptr() if kv is not None else None, B, L,
                                                None, out.data_ptr(), 1 if args.which != 1 else 0, tr, st)

    for it in range(3):
        rc = launch(trace.data_ptr())
        assert rc == 0, eng.lib.avsep_last_error(eng.h).decode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(10):
        launch(None)
    e1.record()
    torch.cuda.synchronize()
    print(f"which={args.which} B={B} L={L} grid={grid}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch (no trace)")
    t = trace.cpu().numpy().reshape(grid, 256).astype(np.int64)
    row, mma = t[:, :128], t[:, 128:]
    t0 = row[:, 0:1]

    def med(a):
        return np.median(a, axis=0)

    def show(name, stamps, labels):
        prev = None
        for idx, lab in labels:
            col = stamps[:, idx]
            if (col == 0).all():
                continue
            v = med(col - t0[:, 0])
            dl = "" if prev is None else f"  (+{v - prev:7.0f})"
            print(f"  {name} {lab:28s} {v:9.0f} clk{dl}")
            prev = v

    rl = [(0, "start"), (1, "x staged")]
    ml = [(0, "mma thread start")]
    for l in range(2):
        tb = 2 + l * 60
        rl += [(tb, f"L{l} LN1 start"), (tb + 1, f"L{l} LN1 done")]
        for h in range(4):
            th = tb + 2 + h * 6
            rl += [(th, f"L{l} h{h} Q|K acc complete"), (th + 1, f"L{l} h{h} Q op / K tile"), (th + 2, f"L{l} h{h} S complete"),
                   (th + 3, f"L{l} h{h} P written"), (th + 4, f"L{l} h{h} O complete"), (th + 5, f"L{l} h{h} O operand")]
        rl += [(tb + 26, f"L{l} attention complete"), (tb + 27, f"L{l} LN2 done")]
        for j in range(8):
            rl += [(tb + 28 + 2 * j, f"L{l} acc1_{j} complete"), (tb + 29 + 2 * j, f"L{l} H_{j} written")]
        rl += [(tb + 44, f"L{l} FFN complete")]
        mb = 1 + l * 60
        ml += [(mb, f"L{l} LN1 seen")]
        for h in range(4):
            th = mb + 1 + h * 6
            ml += [(th, f"L{l} h{h} Q op / K tile seen"), (th + 1, f"L{l} h{h} S (+V) issued"), (th + 2, f"L{l} h{h} P and V seen"),
                   (th + 3, f"L{l} h{h} o_ready seen")]
        ml += [(mb + 25, f"L{l} attention issued"), (mb + 26, f"L{l} LN2 seen")]
        for j in range(8):
            ml += [(mb + 27 + 3 * j, f"L{l} G1_{j} issued"), (mb + 28 + 3 * j, f"L{l} H_{j} seen"), (mb + 29 + 3 * j, f"L{l} G2_{j} issued")]
    rl += [(107, "dec final LN done"), (108, "dec vectors loaded")] + [(109 + j, f"dec H_{j} written") for j in range(4)]
    for c in range(5):
        rl += [(113 + 2 * c, f"dec out {c} mixture loads issued"), (114 + 2 * c, f"dec out {c} acc complete")]
    rl += [(127, "tile done")]
    ml += [(112, "dec a seen"), (113, "dec hidden GEMMs issued"), (114, "dec H complete")] + [(115 + c, f"dec out {c} issued") for c in range(5)]
    show("row", row, rl)
    print()
    show("mma", mma, ml)


if __name__ == "__main__":
    main()
