"""Linear layers evaluated straight from packed 4-bit records.

The reference measures a quantizer by running its perplexity loop over the FAKE-quantized model
(quantization_utils.py:269-322: every nn.Linear still multiplies by a full 16-bit weight).  Here
the same loop can run over the packed export instead: `pack_model` swaps every nn.Linear for a
`QuantLinear` that keeps only the int4 codes + per-group scale / zero point (b200q.export record,
scheme "uniform_asym") and whose forward is `b200q_w4a16_gemm` -- the dequantisation happens inside
the tcgen05 GEMM's operand pipeline, 0.5 byte per weight from HBM.  The reference's
`evaluate_perplexity(model, ...)` then runs unchanged on the packed model.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, export as _export
from .ops import DTYPE_CODE, _on, _stream


def w4a16_linear(x: torch.Tensor, record: Dict, bias: Optional[torch.Tensor] = None,
                 out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """x [..., K] (CUDA, fp16 / bf16) times the packed weight of `record` ([N, K], 4-bit
    "uniform_asym"): returns [..., N] in x's dtype (or fp32 with out_dtype=torch.float32)."""
    assert record["scheme"] == "uniform_asym" and record["bits"] == 4, \
        "w4a16_linear takes 4-bit uniform_asym records (export_uniform / AWQ / SmoothQuant)"
    assert x.is_cuda and x.dtype in (torch.float16, torch.bfloat16), "activations must be CUDA fp16 / bf16"
    N, K = record["shape"]
    assert x.shape[-1] == K
    x2 = x.reshape(-1, K).contiguous()
    M = x2.shape[0]
    out_f32 = out_dtype == torch.float32
    y = torch.empty((M, N), dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    if M > 0:
        q, s, z = record["qweight"], record["scales"], record["zeros"]
        assert q.is_cuda and q.dtype == torch.int32 and q.is_contiguous()
        with _on(x.device):
            rc = _lib.load().b200q_w4a16_gemm(x2.data_ptr(), M, K, DTYPE_CODE[x.dtype], q.data_ptr(),
                                              s.data_ptr(), z.data_ptr(), N, int(record["group"]),
                                              DTYPE_CODE[getattr(torch, record["dtype"])], y.data_ptr(),
                                              int(out_f32), _stream())
        _lib.check(rc, "w4a16_gemm")
    if bias is not None:
        y = y + bias.to(y.dtype)
    return y.reshape(*x.shape[:-1], N)


class QuantLinear(nn.Module):
    """Drop-in for an nn.Linear whose weight exists only as a packed 4-bit record."""

    def __init__(self, record: Dict, bias: Optional[torch.Tensor] = None):
        super().__init__()
        self.out_features, self.in_features = record["shape"]
        self.group, self.bits, self.w_dtype = int(record["group"]), int(record["bits"]), record["dtype"]
        self.register_buffer("qweight", record["qweight"])
        self.register_buffer("scales", record["scales"])
        self.register_buffer("zeros", record["zeros"])
        self.bias = None if bias is None else nn.Parameter(bias.detach(), requires_grad=False)

    def record(self) -> Dict:
        return {"scheme": "uniform_asym", "bits": self.bits, "group": self.group,
                "shape": (self.out_features, self.in_features), "dtype": self.w_dtype,
                "qweight": self.qweight, "scales": self.scales, "zeros": self.zeros}

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return w4a16_linear(x, self.record(), self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, w4 g{self.group}"


@torch.no_grad()
def pack_model(model: nn.Module, q_group_size: int = 128, skip=()) -> nn.Module:
    """Replace (in place) every nn.Linear of `model` by a QuantLinear holding the 4-bit
    pseudo_quantize_tensor codes (quantization_utils.py:395-402) of its CURRENT weight: call it
    INSTEAD of the fake-quantization step -- on the original model for plain w4 g128, or after
    `smooth_weights` for SmoothQuant (the forward-pre-hook that multiplies the inputs by s,
    smooth_quant_quantizer.py:178-199, moves to the new module).  dequantize(record) then equals
    the weight pseudo_quantize_tensor would have written, bit for bit.  (AWQ's salient-channel
    factor is not part of this record format: with scale_factor != 1 the codes describe W * f.)
    Returns the model."""
    for name, m in list(model.named_modules()):
        for child_name, child in list(m.named_children()):
            full = f"{name}.{child_name}" if name else child_name
            if isinstance(child, nn.Linear) and full not in skip:
                rec = _export.export_uniform(child.weight.data, 4, q_group_size)
                ql = QuantLinear(rec, None if child.bias is None else child.bias.data.to(rec["qweight"].device))
                if hasattr(child, "smoothing_scale"):
                    ql.smoothing_scale = child.smoothing_scale
                for hook in child._forward_pre_hooks.values():
                    ql.register_forward_pre_hook(hook)
                setattr(m, child_name, ql)
    return model
