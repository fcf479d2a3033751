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
    else:
        raise ValueError(f"unknown scheme {record['scheme']!r}")
    return w.reshape(shape)


# --------------------------------------------------------------------------------------------------
# model level: one record per nn.Linear, one file per model
# --------------------------------------------------------------------------------------------------
FORMAT_VERSION = 1


def export_model(model, n_bit: int, group: int, scheme: str = "uniform_asym") -> Dict[str, Dict]:
    """{module name: record} for every nn.Linear of `model` (weights may live on the host; they are
    streamed through the GPU).  scheme: "uniform_asym" (pseudo_quantize_tensor) or
    "gptq_column_sym" (the reference's GPTQ column stage)."""
    import torch.nn as nn
    fn = {"uniform_asym": lambda w: export_uniform(w, n_bit, group),
          "gptq_column_sym": lambda w: export_gptq_parity(w, n_bit)}.get(scheme)
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
