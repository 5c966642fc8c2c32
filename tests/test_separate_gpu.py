"""avsep_separate (SeparationDecoder.separate, reference model.py:210-220) and the ticket flags of the masks-only
gather, through the C ABI.  The product is one fp32 multiply per element, so the bar is bit-exact: against the
reference's expression ``masks * mixed_spec.unsqueeze(1)`` evaluated on the CPU, and against the `separated` the fused
decoder epilogue of the full forward writes."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle.weights import CONFIGS, make_inputs, make_state_dict
from tests.helpers import build_model

pytestmark = pytest.mark.gpu


def _engine(F, S):
    from avsep_b200.engine import Engine, EngineConfig
    return Engine(EngineConfig(F, 64, 1, 1, 1, S), 0)


@pytest.mark.parametrize("B,S,F,T", [(3, 2, 257, 63), (1, 3, 65, 32), (2, 1, 5, 7), (1, 2, 3, 1), (5, 3, 2, 1), (64, 2, 257, 63)])
def test_separate_is_the_reference_product_bit_for_bit(B, S, F, T):
    g = torch.Generator().manual_seed(B * 1000 + T)
    masks = torch.rand(B, S, F, T, generator=g)
    mixed = torch.randn(B, F, T, generator=g) * 30.0
    want = masks * mixed.unsqueeze(1)                      # model.py:220 on the CPU
    eng = _engine(F, S)
    got = eng.separate(masks.cuda(), mixed.cuda())
    assert got.shape == (B, S, F, T) and torch.equal(got.cpu(), want)
    # caller-owned output, odd total (not a multiple of the 4-element vector width when B*S*F*T is odd)
    out = torch.full((B, S, F, T), float("nan"), device="cuda")
    assert eng.separate(masks.cuda(), mixed.cuda(), out=out) is out and torch.equal(out.cpu(), want)
    eng.close()


@pytest.mark.parametrize("off_masks,off_out", [(1, 1), (3, 3), (1, 2), (0, 2)])
def test_separate_on_buffers_off_the_16_byte_grid(off_masks, off_out):
    """Views into a larger allocation (e.g. the two halves of the forward's output block): equal misalignment keeps
    the vector walk (head / tail elements one by one), different misalignment takes the element-wise kernel."""
    B, S, F, T = 4, 2, 257, 63
    n = B * S * F * T
    g = torch.Generator().manual_seed(5)
    masks_h, mixed_h = torch.rand(B, S, F, T, generator=g), torch.randn(B, F, T, generator=g)
    eng = _engine(F, S)
    masks = torch.zeros(n + 4, device="cuda")[off_masks:off_masks + n].view(B, S, F, T)
    masks.copy_(masks_h)
    block = torch.full((n + 8,), 7.0, device="cuda")
    out = block[off_out:off_out + n].view(B, S, F, T)
    eng.separate(masks, mixed_h.cuda(), out=out)
    assert torch.equal(out.cpu(), masks_h * mixed_h.unsqueeze(1))
    assert bool((block[:off_out] == 7.0).all()) and bool((block[off_out + n:] == 7.0).all())     # nothing outside
    eng.close()


def test_separate_matches_the_fused_decoder_epilogue_and_the_module_call():
    cfg = CONFIGS["default"]
    P = make_state_dict(cfg, seed=71, gain=2.0)
    mixed, frames = make_inputs(cfg, 5, 63, 50, 32, 32, seed=71, kind="dataset")
    model = build_model(cfg, P, "bf16")
    m_d, f_d = torch.from_numpy(mixed).cuda(), torch.from_numpy(frames).cuda()
    sep, masks = model(m_d, f_d)
    assert torch.equal(model.engine.separate(masks, m_d), sep)
    # the sub-module drop-in: CUDA tensors in place, CPU tensors through the device (how the reference's tests call it)
    assert torch.equal(model.decoder.separate(masks, m_d), sep)
    cpu = model.decoder.separate(masks.cpu(), torch.from_numpy(mixed))
    assert not cpu.is_cuda and torch.equal(cpu, sep.cpu())


def test_separate_refuses_wrong_shapes():
    eng = _engine(65, 2)
    with pytest.raises(ValueError):
        eng.separate(torch.rand(2, 2, 64, 8, device="cuda"), torch.rand(2, 64, 8, device="cuda"))      # F
    with pytest.raises(ValueError):
        eng.separate(torch.rand(2, 3, 65, 8, device="cuda"), torch.rand(2, 65, 8, device="cuda"))      # S
    with pytest.raises(ValueError):
        eng.separate(torch.rand(3, 2, 65, 8, device="cuda"), torch.rand(2, 65, 8, device="cuda"))      # B
    eng.close()


def test_ticket_flags_order_two_streams():
    """avsep_flag_wait holds its stream until avsep_flag_signal (issued later, on another stream) raised every flag."""
    eng = _engine(65, 2)
    flags = torch.zeros(3 * 16, dtype=torch.int32, device="cuda")
    ptrs = [flags.data_ptr() + 64 * i for i in range(3)]
    arr = (C.c_void_p * 3)(*ptrs)
    waiter, signaller = torch.cuda.Stream(), torch.cuda.Stream()
    data = torch.zeros(1 << 20, device="cuda")
    seen = torch.zeros(1, device="cuda")
    # every kernel the test launches while the waiter spins is loaded beforehand: with lazy module loading a first
    # launch may wait for running kernels (the library preloads its own for the same reason)
    data.fill_(1.0)
    seen.copy_(data[-1:])
    data.zero_()
    torch.cuda.synchronize()
    rc = eng.lib.avsep_flag_wait(eng.h, arr, 3, 7, C.c_double(20.0), C.c_void_p(waiter.cuda_stream))
    assert rc == 0, eng.lib.avsep_last_error(eng.h)
    with torch.cuda.stream(waiter):
        seen.copy_(data[-1:])                  # runs only after the wait: must observe what the signaller wrote first
    assert not waiter.query()                  # still waiting: no ticket yet
    with torch.cuda.stream(signaller):
        data.fill_(3.0)
    for i in range(3):                         # the three flags are raised one by one, the last one completes the wait
        one = (C.c_void_p * 1)(ptrs[i])
        rc = eng.lib.avsep_flag_signal(eng.h, one, 1, 7 + i, C.c_void_p(signaller.cuda_stream))
        assert rc == 0, eng.lib.avsep_last_error(eng.h)
    torch.cuda.synchronize()
    assert float(seen) == 3.0
    assert flags.view(3, 16)[:, 0].tolist() == [7, 8, 9]
    # tickets only grow: a wait for an older ticket returns at once
    assert eng.lib.avsep_flag_wait(eng.h, arr, 3, 5, C.c_double(1.0), C.c_void_p(waiter.cuda_stream)) == 0
    torch.cuda.synchronize()
    assert eng.lib.avsep_flag_wait(eng.h, arr, 0, 1, C.c_double(1.0), None) != 0       # n out of range
    assert eng.lib.avsep_flag_wait(eng.h, arr, 3, 1, C.c_double(0.0), None) != 0       # no timeout
    eng.close()
