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
