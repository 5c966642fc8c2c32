"""Test infrastructure: stage the UNMODIFIED reference next to the oracle so that it can travel to the GPU box.

    python oracle/make_ref.py            # /root/reference -> oracle/_ref/   (git-ignored, shipped by gpurun)

The reference (danieleschmidt/AV-Separation-Transformer) is four pure-Python files on torch; there is nothing to
compile.  This recipe copies them byte for byte from where they lie under /root/reference into ``oracle/_ref/``
(never into the tracked tree), together with the reference's own test file, and writes a manifest of SHA-256 sums.
Users of ``oracle/_ref`` (all test / measurement infrastructure, never the product path):
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg: the reference's own modules on the host cores
    (``cpu_baseline.kind == "reference"``); falls back to the oracle port (kind "port") when ``oracle/_ref`` is absent;
  * ``tests/test_reference_conformance_gpu.py``: the reference's own tests/test_model.py run against the drop-in;
  * ``tests/test_oracle_golden.py``: the live reference against the committed golden vectors when present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AVSEP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = {
    "src/av_separation/__init__.py": "av_separation/__init__.py",
    "src/av_separation/model.py": "av_separation/model.py",
    "src/av_separation/dataset.py": "av_separation/dataset.py",
    "src/av_separation/losses.py": "av_separation/losses.py",
    "tests/test_model.py": "tests/test_model.py",
}


def stage(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is (now) populated, False when the reference tree is not available here."""
    if not os.path.isdir(REF_ROOT):
        return os.path.exists(os.path.join(OUT, "MANIFEST.json"))
    manifest = {}
    for src, dst in FILES.items():
        sp, dp = os.path.join(REF_ROOT, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(dp), exist_ok=True)
        shutil.copyfile(sp, dp)
        os.chmod(dp, 0o644)
        with open(dp, "rb") as f:
            manifest[dst] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: staged {len(manifest)} files from {REF_ROOT}")
    return True


def ref_available() -> bool:
    return os.path.exists(os.path.join(OUT, "av_separation", "model.py"))


def load_reference_model_module():
    """Import oracle/_ref/av_separation/model.py under a private name (the product also ships an ``av_separation``
    package -- the drop-in -- so the reference must not be imported by that name)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_avsep_reference_model", os.path.join(OUT, "av_separation", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
