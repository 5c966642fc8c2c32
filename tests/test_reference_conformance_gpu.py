"""The reference's OWN test file (tests/test_model.py, staged unmodified into oracle/_ref/tests by
oracle/make_ref.py) run against the drop-in ``av_separation`` package on the GPU box.

The reference tests construct each module and call it with unseeded CPU ``randn`` tensors
(/root/reference/tests/test_model.py:39-51,77-224); the drop-in routes CPU tensors through the device and returns CPU
tensors, so the shape / range / dependence assertions run unchanged.  Deselected: the tests that need autograd or a
training objective (gradient_flow, backward_pass, TestLosses, one_training_step) -- training is out of scope.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "av-separation-transformer_b200")
REF_TEST = os.path.join(ROOT, "oracle", "_ref", "tests", "test_model.py")
DESELECT = "not gradient_flow and not backward and not TestLosses and not training_step"


def test_reference_test_suite_runs_against_the_dropin(tmp_path):
    if not os.path.exists(REF_TEST):
        pytest.skip("oracle/_ref not staged (run python oracle/make_ref.py where /root/reference exists)")
    # run from an empty directory: the file does sys.path.insert(0, "src"), which must not find anything, and the
    # drop-in package directory comes first on PYTHONPATH so that `av_separation` is the B200 implementation
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([PKG, env.get("PYTHONPATH", "")])
    probe = subprocess.run([sys.executable, "-c", "import av_separation, avsep_b200; print(av_separation.__file__)"],
                           cwd=tmp_path, env=env, capture_output=True, text=True)
    assert probe.returncode == 0 and PKG in probe.stdout, probe.stdout + probe.stderr
    res = subprocess.run([sys.executable, "-m", "pytest", REF_TEST, "-q", "-x", "-p", "no:cacheprovider",
                          "-k", DESELECT, "--rootdir", str(tmp_path)],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-3000:]
    print(tail)
    assert res.returncode == 0, tail
    # 22 of the reference's 30 tests remain after the deselection above; none may be skipped
    assert " passed" in res.stdout and "skipped" not in res.stdout and "failed" not in res.stdout, tail
    n_passed = int(res.stdout.rsplit(" passed", 1)[0].split()[-1])
    assert n_passed >= 22, tail
