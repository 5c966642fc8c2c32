#!/usr/bin/env python
"""Instruction census of the built library: which kernels use the Blackwell tensor-core (UTCHMMA = tcgen05.mma),
tensor-memory (LDTM / STTM = tcgen05.ld / st), TMA (UTMALDG / UTMASTG tensor copies, UBLKCP bulk copies) and mbarrier
(SYNCS) instructions, which still use the legacy HMMA path, and how many local-memory (spill) accesses each has.

    python tools/sass_census.py [lib] > profiles/<round>_sass_census.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "av-separation-transformer_b200/lib/libavsep.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDL", "STL"]
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[fn]["total"] += 1
        for k in keys:
            if op == k or (k == "HMMA" and op.startswith("HMMA")):
                counts[fn][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k:>7s}" for k in keys))
rows = []
for (fn, c), name in zip(counts.items(), names):
    name = name.replace("avsep::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
    rows.append((name, c))
for name, c in sorted(rows):
    print(f"{name[:58]:58s} {c['total']:7d} " + " ".join(f"{c[k]:7d}" for k in keys))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"{'TOTAL':58s} {tot['total']:7d} " + " ".join(f"{tot[k]:7d}" for k in keys))
