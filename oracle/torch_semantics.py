"""Explicit models of two torch-CPU behaviours the CUDA kernels reproduce literally.

TEST INFRASTRUCTURE ONLY (see oracle/quant_oracle.py for the import rule).

1. `rowsum_model`: the order in which ATen adds a contiguous fp32 row (aten/src/ATen/native/cpu/
   SumKernel.cpp: vectorized_inner_sum -> row_sum -> multi_row_sum).  ((w - wq)**2).sum(dim=1) in
   pot_apot_quantizer.py:94,307 goes through it, and the strict `<` over candidate errors makes the
   POT/APOT scale choice depend on every bit of that sum.  The kernel is compiled for AVX2 only in
   torch (8-float vectors) even on AVX-512 hosts, so the order is host-independent.
2. `log2_round_threshold` / `log2_floor_threshold`: where rne(log2f(r)) and floor(log2f(m)) step,
   modelled as double log2 rounded to float (what SLEEF's u10 log2f returns at those points).

tests/test_torch_semantics.py checks both against torch itself on the machine the tests run on.
"""
from __future__ import annotations

import numpy as np


def _ceil_log2(x: int) -> int:
    return 0 if x <= 1 else (x - 1).bit_length()


def rowsum_model(x: np.ndarray, V: int = 8) -> np.ndarray:
    """fp32 row sums of x [rows, G] in ATen's order."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, G = x.shape
    if G < V:
        # shorter than one vector: ATen's scalar_inner_sum, 4 interleaved scalar accumulators
        a = np.zeros((4, rows), np.float32)
        for i in range(G // 4):
            for k in range(4):
                a[k] += x[:, i * 4 + k]
        for t in range(G // 4 * 4, G):
            a[0] += x[:, t]
        for k in range(1, 4):
            a[0] += a[k]
        return a[0]
    vec_size = G // V
    ilp = 4
    size_ilp = vec_size // ilp
    num_levels = 4
    level_power = max(4, _ceil_log2(size_ilp) // num_levels)
    level_step = 1 << level_power
    level_mask = level_step - 1
    acc = np.zeros((num_levels, ilp, rows, V), np.float32)

    def vec(i):
        return x[:, i * V:(i + 1) * V]

    i = 0
    while i + level_step <= size_ilp:
        for _ in range(level_step):
            for k in range(ilp):
                acc[0, k] += vec(i * ilp + k)
            i += 1
        for j in range(1, num_levels):
            acc[j] += acc[j - 1]
            acc[j - 1] = 0
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size_ilp:
        for k in range(ilp):
            acc[0, k] += vec(i * ilp + k)
        i += 1
    for j in range(1, num_levels):
        acc[0] += acc[j]
    part = acc[0]
    for vi in range(size_ilp * ilp, vec_size):
        part[0] += vec(vi)
    for k in range(1, ilp):
        part[0] += part[k]
    total = np.zeros(rows, np.float32)
    for t in range(vec_size * V, G):
        total += x[:, t]
    for lane in range(V):
        total += part[0][:, lane]
    return total


def rowsum_model16(x: np.ndarray) -> np.ndarray:
    """fp32 row sums (before the final rounding to the tensor dtype) of x [rows, G] holding fp16 /
    bf16 VALUES, in the order ATen uses for 16-bit input: a load brings 16 elements and returns the
    8-lane fp32 vector low8 + high8; those vectors then go through the same 4-accumulator cascade
    as fp32 input; rows shorter than 16 use 4 interleaved scalar accumulators."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, G = x.shape
    if G < 16:
        return rowsum_model(x, V=1 << 30)          # forces the scalar_inner_sum branch
    nv = G // 16
    vecs = np.concatenate([(x[:, j * 16:j * 16 + 8] + x[:, j * 16 + 8:j * 16 + 16]).astype(np.float32)
                           for j in range(nv)], axis=1)
    head = rowsum_model_parts(vecs)                 # [rows, 8] lane totals, not yet added up
    total = np.zeros(rows, np.float32)
    for t in range(nv * 16, G):
        total += x[:, t]
    for lane in range(8):
        total += head[:, lane]
    return total


def rowsum_model_parts(x: np.ndarray, V: int = 8) -> np.ndarray:
    """The 8 lane totals ATen holds before its final horizontal add (x is [rows, n*V])."""
    rows, G = x.shape
    vec_size = G // V
    ilp, num_levels = 4, 4
    size_ilp = vec_size // ilp
    level_power = max(4, _ceil_log2(size_ilp) // num_levels)
    level_step = 1 << level_power
    level_mask = level_step - 1
    acc = np.zeros((num_levels, ilp, rows, V), np.float32)
    i = 0
    while i + level_step <= size_ilp:
        for _ in range(level_step):
            for k in range(ilp):
                acc[0, k] += x[:, (i * ilp + k) * V:(i * ilp + k + 1) * V]
            i += 1
        for j in range(1, num_levels):
            acc[j] += acc[j - 1]
            acc[j - 1] = 0
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size_ilp:
        for k in range(ilp):
            acc[0, k] += x[:, (i * ilp + k) * V:(i * ilp + k + 1) * V]
        i += 1
    for j in range(1, num_levels):
        acc[0] += acc[j]
    part = acc[0]
    for vi in range(size_ilp * ilp, vec_size):
        part[0] += x[:, vi * V:(vi + 1) * V]
    for k in range(1, ilp):
        part[0] += part[k]
    return part[0]


def _log2f_model(bits: np.ndarray) -> np.ndarray:
    r = bits.astype(np.uint32).view(np.float32)
    with np.errstate(divide="ignore"):
        return np.log2(r.astype(np.float64)).astype(np.float32)


def _first_bits(pred) -> int:
    lo, hi = 1, 0x7F800000
    while lo < hi:
        mid = (lo + hi) // 2
        if pred(mid):
            hi = mid
        else:
            lo = mid + 1
    return lo


def log2_round_threshold(e: int) -> int:
    """bit pattern of the smallest positive float r with rne(log2f(r)) >= e + 1"""
    return _first_bits(lambda b: np.rint(_log2f_model(np.array([b]))[0]) >= e + 1)


def log2_floor_threshold(e: int) -> int:
    """bit pattern of the smallest positive float m with floor(log2f(m)) >= e"""
    return _first_bits(lambda b: np.floor(_log2f_model(np.array([b]))[0]) >= e)
