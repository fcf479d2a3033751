"""Pins the two torch-CPU behaviours the POT/APOT kernels reproduce (oracle/torch_semantics.py)
against torch on this host, and the library's host tables against both.  CPU only."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import torch_semantics as TS
from b200q import _lib


@pytest.mark.parametrize("G", [1, 3, 7, 8, 9, 16, 31, 32, 64, 100, 128, 200, 256, 768, 1024, 4096, 11008])
def test_rowsum_order_model_is_torchs(G):
    g = torch.Generator().manual_seed(G)
    for rows in (1, 5, 1000):
        x = (torch.randn(rows, G, generator=g) * 0.02) ** 2
        want = x.sum(dim=1, keepdim=True).numpy()[:, 0]
        got = TS.rowsum_model(x.numpy())
        assert np.array_equal(want, got), f"G={G} rows={rows}"


def _scan(center: float, half_width: int = 1 << 12):
    c = np.array([center], np.float32).view(np.int32)[0]
    bits = np.arange(c - half_width, c + half_width, dtype=np.int64)
    bits = bits[(bits > 0) & (bits < 0x7F800000)].astype(np.int32)
    r = torch.from_numpy(bits.view(np.float32).copy())
    return bits, torch.log2(r)


def test_log2_round_steps_match_torch_and_library():
    lib = _lib.load()
    for e in range(-20, 128):
        bits, lg = _scan(np.sqrt(2.0) * 2.0 ** e)
        E = torch.round(lg).numpy()
        assert (np.diff(E) >= 0).all()
        first = int(bits[np.argmax(E >= e + 1)])
        assert E[np.argmax(E >= e + 1) - 1] == e
        assert first == TS.log2_round_threshold(e), f"model, e={e}"
        assert first == lib.b200q_log2_round_threshold_bits(e), f"library, e={e}"


def test_log2_floor_steps_match_torch_and_library():
    lib = _lib.load()
    for e in range(-60, 128):
        bits, lg = _scan(2.0 ** e)
        F = torch.floor(lg).numpy()
        assert (np.diff(F) >= 0).all()
        first = int(bits[np.argmax(F >= e)])
        assert first == TS.log2_floor_threshold(e), f"model, e={e}"
        assert first == lib.b200q_log2_floor_threshold_bits(e), f"library, e={e}"


def test_pot_exponent_rule_matches_torch_on_random_ratios():
    """E = clamp(rne(log2(r)), 0, Emax) evaluated from the step table equals torch's, for ratios
    of the magnitude the POT search produces."""
    lib = _lib.load()
    thr = np.array([lib.b200q_log2_round_threshold_bits(e) for e in range(0, 127)], dtype=np.int64)
    g = torch.Generator().manual_seed(7)
    r = torch.exp2(torch.rand(2_000_000, generator=g) * 12 - 2)
    # add exact step neighbours
    near = torch.from_numpy(np.concatenate([thr[:12] + d for d in (-2, -1, 0, 1)]).astype(np.int32)
                            .view(np.float32).copy())
    r = torch.cat([r, near])
    for emax in (3, 7, 127):
        want = torch.clamp(torch.round(torch.log2(torch.clamp(r, min=1e-10))), 0, emax).numpy()
        bits = r.numpy().view(np.int32).astype(np.int64)
        e = np.clip((bits >> 23) - 127, 0, emax)
        t = np.where(e < emax, thr[np.minimum(e, 126)], 0xFFFFFFFF)
        got = e + (bits >= t)
        assert np.array_equal(want.astype(np.int64), got)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("G", [1, 5, 12, 15, 16, 24, 40, 96, 100, 128, 200, 256, 768, 4096])
def test_rowsum_order_model_16bit_is_torchs(dtype, G):
    g = torch.Generator().manual_seed(G + 1)
    for rows in (1, 6, 500):
        x = ((torch.randn(rows, G, generator=g) * 0.02) ** 2).to(dtype)
        want = x.sum(dim=1)
        got = torch.from_numpy(TS.rowsum_model16(x.float().numpy())).to(dtype)
        assert torch.equal(want, got), f"{dtype} G={G} rows={rows}"


@pytest.mark.parametrize("dtype,code", [(torch.float16, 1), (torch.bfloat16, 2)])
def test_log2_steps_16bit_match_torch_and_library(dtype, code):
    """Every positive finite fp16 / bf16 value: rne(log2(r)) and floor(log2(r)) evaluated in that
    dtype by torch step exactly where the library's per-dtype tables say."""
    lib = _lib.load()
    if dtype == torch.float16:
        vals = torch.arange(1, 0x7C00, dtype=torch.int32).to(torch.int16).view(torch.float16)
    else:
        vals = torch.arange(1, 0x7F80, dtype=torch.int32).to(torch.int16).view(torch.bfloat16)
    lg = torch.log2(vals)
    R = torch.round(lg).float().numpy()
    F = torch.floor(lg).float().numpy()
    bits = vals.float().numpy().view(np.int32).astype(np.int64)
    assert (np.diff(R) >= 0).all() and (np.diff(F) >= 0).all()
    for e in range(-127, 128):
        hit = np.nonzero(R >= e + 1)[0]
        want = int(bits[hit[0]]) if len(hit) else 0x7F800000
        assert lib.b200q_log2_round_threshold_bits_dt(e, code) == want, f"round e={e}"
    for e in range(-149, 128):
        hit = np.nonzero(F >= e)[0]
        want = int(bits[hit[0]]) if len(hit) else 0x7F800000
        assert lib.b200q_log2_floor_threshold_bits_dt(e, code) == want, f"floor e={e}"
