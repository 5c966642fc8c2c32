"""CPU: the C-ABI library loads and exports every symbol include/avsep.h declares; the host mirror keeps the
reference's module/state_dict contract.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle.weights import CONFIGS, make_state_dict, state_dict_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "avsep.h")).read()
    return sorted(set(re.findall(r"AVSEP_API[^;(]*?\b(avsep_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from avsep_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms       # binding, header and library agree


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from avsep_b200 import AVSeparationTransformer
    m = AVSeparationTransformer(freq_bins=65, d_model=64, nhead=4, num_encoder_layers=1, num_fusion_layers=1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 65, 8), torch.zeros(1, 4, 16, 16))     # host path still needs the device: no CPU fallback


@pytest.mark.parametrize("cname", ["tiny", "default"])
def test_state_dict_contract(cname):
    from avsep_b200 import AVSeparationTransformer
    from avsep_b200.engine import EngineConfig, expected_shapes
    cfg = CONFIGS[cname]
    m = AVSeparationTransformer(**cfg.as_dict())
    sd = m.state_dict()
    spec = {k: tuple(s) for k, s, _, _ in state_dict_spec(cfg)}
    assert set(sd) == set(spec)
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k], k
    want = expected_shapes(EngineConfig(**cfg.as_dict()))
    assert set(want) == {k for k in spec if not k.endswith("num_batches_tracked")}
    P = make_state_dict(cfg, seed=0)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}, strict=True)
    assert not m.training
    with pytest.raises(NotImplementedError):
        m.train()


def test_import_names_of_the_reference_package():
    import av_separation
    from av_separation.model import AVSeparationTransformer, AudioEncoder  # noqa: F401
    assert av_separation.AVSeparationTransformer is AVSeparationTransformer
