"""Packed integer export of quantized Linears.

The reference only fake-quantizes: codes, scales and zero points are temporaries inside
`pseudo_quantize_tensor` (quantization_utils.py:395-407) and `_gptq_quantize_layer`
(gptq_quantizer.py:183-186), and `benchmark_runner.py:732-743` saves metrics only.  The kernels
here already emit those integers; this module packs them into the on-disk form (one little-endian
bit stream per row, `b200q_pack_codes`) next to the fp32 scales / zero points, and turns a packed
record back into the fake-quantized weight so that

        dequantize(export_*(W, ...)) == the drop-in quantizer's output, bit for bit.

Records are plain dicts of tensors and ints, ready for `torch.save` / safetensors.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib
from . import ops as _ops


def pack_codes(codes: torch.Tensor, n_bit: int) -> torch.Tensor:
    """uint8 CUDA codes [..., K] (values < 2^n_bit) -> int32 [N, ceil(K * n_bit / 32)]."""
    assert codes.is_cuda and codes.dtype == torch.uint8
    codes = codes.contiguous()
    K = codes.shape[-1]
    N = codes.numel() // K
    lib = _lib.load()
    words = lib.b200q_packed_words_per_row(K, n_bit)
    if words <= 0:
        raise AssertionError(f"pack_codes: unsupported shape/bit width (K={K}, n_bit={n_bit})")
    packed = torch.empty((N, words), dtype=torch.int32, device=codes.device)
    with _ops._on(codes.device):
        rc = lib.b200q_pack_codes(codes.data_ptr(), N, K, n_bit, packed.data_ptr(), _ops._stream())
    _lib.check(rc, "pack_codes")
    return packed


def unpack_codes(packed: torch.Tensor, K: int, n_bit: int) -> torch.Tensor:
    """int32 [N, words] -> uint8 [N, K]."""
    assert packed.is_cuda and packed.dtype == torch.int32 and packed.dim() == 2
    packed = packed.contiguous()
    N = packed.shape[0]
    lib = _lib.load()
    assert packed.shape[1] == lib.b200q_packed_words_per_row(K, n_bit), "packed width does not match K"
    codes = torch.empty((N, K), dtype=torch.uint8, device=packed.device)
    with _ops._on(packed.device):
        rc = lib.b200q_unpack_codes(packed.data_ptr(), N, K, n_bit, codes.data_ptr(), _ops._stream())
    _lib.check(rc, "unpack_codes")
    return codes


def export_uniform(W: torch.Tensor, n_bit: int, group: int) -> Dict:
    """`pseudo_quantize_tensor` (quantization_utils.py:362-413) with the integers kept: asymmetric
    codes in [0, 2^b - 1] packed at b bits, fp32 scale and zero point per group."""
    W = _ops.to_device(W)
    out, codes, scales, zeros = _ops.group_fakequant(W, n_bit, group, return_codes=True)
    K = W.shape[-1]
    return {"scheme": "uniform_asym", "bits": n_bit, "group": group if group > 0 else K,
            "shape": tuple(W.shape), "dtype": str(W.dtype).replace("torch.", ""),
            "qweight": pack_codes(codes, n_bit), "scales": scales, "zeros": zeros}


def export_gptq_parity(W: torch.Tensor, n_bit: int) -> Dict:
    """The reference's GPTQ column stage (gptq_quantizer.py:167-206) with the integers kept: signed
    codes in [-2^b, 2^b - 1] stored with offset 2^b at b + 1 bits, one fp32 scale per column."""
    W = _ops.to_device(W)
    out, codes, scales = _ops.gptq_parity_quant(W, n_bit, return_codes=True)
    offset = 1 << n_bit
    ucodes = (codes.to(torch.int16) + offset).to(torch.uint8)
    return {"scheme": "gptq_column_sym", "bits": n_bit + 1, "offset": offset, "shape": tuple(W.shape),
            "dtype": str(W.dtype).replace("torch.", ""), "qweight": pack_codes(ucodes, n_bit + 1),
            "scales": scales}


def _bit_plane(mask: torch.Tensor) -> torch.Tensor:
    return pack_codes(mask.to(torch.uint8).contiguous(), 1)


def export_pot(W: torch.Tensor, n_bit: int, group: int) -> Dict:
    """`pot_quantize_tensor` (pot_apot_quantizer.py:25-115) with the integers kept.  Every value is
    s * sign(w) * 2^E (:104-107): the code of an element is E (n_bit - 1 bits, E in
    [0, 2^(n_bit-1) - 1]) with the sign in the top bit, packed at n_bit bits; one fp32 scale per
    group (the winning point of the 200-candidate search, `grid_index` says which).  sign(0) = 0
    makes exact zeros a 2^n_bit + 1-th symbol: they are kept as a 1-bit plane, present only when the
    tensor contains one."""
    W = _ops.to_device(W)
    K = W.shape[-1]
    G = group if group > 0 else K
    assert K % G == 0
    groups = W.reshape(-1, G)
    out, exps, scale, idx = _ops.pot_quant(groups, n_bit, torch.arange(0.01, 2.01, 0.01), return_codes=True)
    neg = torch.signbit(out) & (out != 0)
    codes = exps | (neg.to(torch.uint8) << (n_bit - 1))
    zero = out == 0
    rec = {"scheme": "pot", "bits": n_bit, "group": G, "shape": tuple(W.shape),
           "dtype": str(W.dtype).replace("torch.", ""),
           "qweight": pack_codes(codes.reshape(-1, K), n_bit), "scales": scale, "grid_index": idx}
    if bool(zero.any()):
        rec["zero_mask"] = _bit_plane(zero.reshape(-1, K))
    return rec


def export_apot(W: torch.Tensor, n_bit: int, group: int, k: int = 2, total_elements: int = None) -> Dict:
    """`apot_quantize_tensor` (pot_apot_quantizer.py:192-351) with the integers kept.  Every value
    is s * level[i] (:294-298, :331-340): the code is the index i into the signed level table
    (31 entries at w4 k2, at most 32), packed at ceil(log2(#levels)) bits; one fp32 scale per group;
    the table itself travels in the record.  total_elements = element count of the whole tensor when
    W is a row shard (it selects the 20- or 40-point grid, :258-262)."""
    import sys
    from pathlib import Path
    pkg = str(Path(__file__).resolve().parent.parent)
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from pot_apot_quantizer import _apot_signed_levels
    W = _ops.to_device(W)
    K = W.shape[-1]
    G = group if group > 0 else K
    assert K % G == 0
    groups = W.reshape(-1, G)
    levels = _apot_signed_levels(n_bit, k)
    total = W.numel() if total_elements is None else int(total_elements)
    grid = torch.arange(0.01, 2.01, 0.1 if total > 500000 else 0.05)
    out, lidx, scale, idx = _ops.apot_quant(groups, levels, grid, return_codes=True)
    bits = max(1, (levels.numel() - 1).bit_length())
    return {"scheme": "apot", "bits": bits, "group": G, "shape": tuple(W.shape),
            "dtype": str(W.dtype).replace("torch.", ""), "levels": levels.to(W.device),
            "qweight": pack_codes(lidx.reshape(-1, K), bits), "scales": scale, "grid_index": idx}


def dequantize(record: Dict) -> torch.Tensor:
    """The fake-quantized weight a record stands for, in the record's dtype."""
    shape = tuple(record["shape"])
    K = shape[-1]
    dtype = getattr(torch, record["dtype"])
    codes = unpack_codes(record["qweight"], K, record["bits"]).float()
    if record["scheme"] == "uniform_asym":
        G = record["group"]
        q = codes.reshape(-1, G)
        s = record["scales"].reshape(-1, 1).to(dtype)
        z = record["zeros"].reshape(-1, 1).to(dtype)
        # (q - z) * s evaluated per op in the weight's dtype, like quantization_utils.py:407
        w = ((q.to(dtype) - z) * s)
    elif record["scheme"] == "gptq_column_sym":
        q = (codes - float(record["offset"])).to(dtype)
        w = q * record["scales"].reshape(1, K).to(dtype)          # gptq_quantizer.py:186
    elif record["scheme"] == "pot":
        G, b = record["group"], record["bits"]
        c = codes.to(torch.int32).reshape(-1, G)
        E = (c & ((1 << (b - 1)) - 1)).to(dtype)
        sign = torch.where((c >> (b - 1)) != 0, -1.0, 1.0).to(dtype)
        # scale * sign * 2^E (pot_apot_quantizer.py:105-107): powers of two, exact in any order
        w = record["scales"].reshape(-1, 1).to(dtype) * sign * torch.pow(torch.tensor(2.0, dtype=dtype, device=c.device), E)
        if "zero_mask" in record:
            z = unpack_codes(record["zero_mask"], K, 1).reshape(-1, G).bool()
            w = torch.where(z, torch.zeros((), dtype=dtype, device=w.device), w)
    elif record["scheme"] == "apot":
        G = record["group"]
        idx = codes.long().reshape(-1, G)
        # the reference gathers the levels into a tensor of the WEIGHT's dtype (zeros_like(w), :326,
        # :335) and multiplies by the scale there (:340): level rounded to dtype, product rounded
        w = record["scales"].reshape(-1, 1).to(dtype) * record["levels"].to(dtype)[idx]
    else:
        raise ValueError(f"unknown scheme {record['scheme']!r}")
    return w.reshape(shape)


# --------------------------------------------------------------------------------------------------
# model level: one record per nn.Linear, one file per model
# --------------------------------------------------------------------------------------------------
FORMAT_VERSION = 1


def export_model(model, n_bit: int, group: int, scheme: str = "uniform_asym") -> Dict[str, Dict]:
    """{module name: record} for every nn.Linear of `model` (weights may live on the host; they are
    streamed through the GPU).  scheme: "uniform_asym" (pseudo_quantize_tensor),
    "gptq_column_sym" (the reference's GPTQ column stage), "pot" or "apot" (exponent / level-index
    planes of pot_apot_quantizer.py)."""
    import torch.nn as nn
    fn = {"uniform_asym": lambda w: export_uniform(w, n_bit, group),
          "gptq_column_sym": lambda w: export_gptq_parity(w, n_bit),
          "pot": lambda w: export_pot(w, n_bit, group),
          "apot": lambda w: export_apot(w, n_bit, group)}.get(scheme)
    if fn is None:
        raise ValueError(f"unknown scheme {scheme!r}")
    return {name: fn(m.weight.data) for name, m in model.named_modules() if isinstance(m, nn.Linear)}


def save_records(path, records: Dict[str, Dict]) -> None:
    """Write {name: record} to `path` (torch.save of plain host tensors and ints)."""
    host = {name: {k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in rec.items()}
            for name, rec in records.items()}
    torch.save({"format": "b200q-packed", "version": FORMAT_VERSION, "records": host}, path)


def load_records(path, device=None) -> Dict[str, Dict]:
    """Read a file written by save_records; tensors go to `device` when given."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(blob, dict) or blob.get("format") != "b200q-packed":
        raise ValueError(f"{path}: not a b200q packed-weights file")
    if blob.get("version") != FORMAT_VERSION:
        raise ValueError(f"{path}: format version {blob.get('version')} (this build reads {FORMAT_VERSION})")
    records = blob["records"]
    if device is not None:
        records = {name: {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in rec.items()}
                   for name, rec in records.items()}
    return records
