"""gptq_quantizer — drop-in for the reference module of the same name.

Entry points and defaults are the reference's (gptq_quantizer.py:22,79,112,210).  The arithmetic
runs in libb200quant:

  * Hessian  H = sum_i x_i^T x_i / (||x_i|| + 1e-5)^2, / len(feats) + damp I   -> b200q.hessian
  * damped SPD inverse                                                        -> b200q.spd_inverse
  * column stage                                                              -> b200q kernels

MODE selects what the column stage does with H^-1:
  "parity" (default)  exactly what the reference computes: every column rounded with its own
                      scale over all rows, no error compensation, H^-1 unused by the output
                      (gptq_quantizer.py:189-194).  Outputs are bit-identical to the reference.
  "compensated"       the GPTQ-paper loop the reference sketches and skips (opt-in, parity
                      unpinned by the reference).
BUILD_HESSIAN controls whether parity mode still builds H and H^-1 like the reference does (they
cannot influence its output); the default keeps the work for like-for-like timing.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.nn as nn

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from b200q import ops as _ops  # noqa: E402
from b200q import dist as _dist  # noqa: E402
from b200q import pipeline as _pipeline  # noqa: E402

MODE = "parity"
BUILD_HESSIAN = True


# ==================================================================================================
# model walker
# ==================================================================================================
@torch.no_grad()
def gptq_quantize_model_weight(
    model: nn.Module,
    w_bit: int,
    q_group_size: int,
    input_feat: Dict[str, List[torch.Tensor]],
    perp_damp: float = 0.01,
    blocksize: int = 128,
    nsamples: int = 128,
    actorder: bool = False,
    verbose: bool = True,
) -> None:
    """Quantize every nn.Linear in place: GPTQ for layers with calibration features, the symmetric
    group quantizer for the rest (reference: gptq_quantizer.py:58-75)."""
    if verbose:
        print("Applying GPTQ quantization...")
    items = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)]
    calibrated = [(n, m) for n, m in items if n in input_feat]
    position = {n: i for i, (n, _) in enumerate(calibrated)}
    ready = {}

    def compute(name, _module, W):
        if name not in input_feat:
            return _ops.group_fakequant(W, w_bit, q_group_size, symmetric=True)
        if name not in ready:
            # Under row sharding the inverse of one layer's Hessian is a single-GPU job, so a GROUP
            # of world-size layers is prepared at once: every rank adds its calibration samples to
            # each layer's Hessian (all-reduced), then rank j factors layer j of the group while the
            # others factor theirs, and the factors are broadcast.  Unsharded: groups of one.
            i = position[name]
            group = calibrated[i:i + _dist.world_size()]
            prepared = [_prepare(input_feat[n], m.weight.shape[1], W.device, perp_damp, nsamples,
                                 actorder, owner=j) for j, (n, m) in enumerate(group)]
            for j, ((n, _m), p) in enumerate(zip(group, prepared)):
                if p[2] is not None:
                    _dist.broadcast(p[2], j)
                ready[n] = p
        H, perm, factor = ready.pop(name)
        return _column_stage(W, w_bit, q_group_size, blocksize, H, perm, factor)

    _pipeline.run_layers(items, compute)


# ==================================================================================================
# per-layer stages
# ==================================================================================================
@torch.no_grad()
def _simple_quantize_layer(layer: nn.Linear, n_bit: int, q_group_size: int) -> None:
    """Symmetric |max| group quantization, codes in [-2^b, 2^b-1] (reference: :79-108)."""
    w = layer.weight.data
    src = w.device
    out = _ops.group_fakequant(_ops.to_device(w), n_bit, q_group_size, symmetric=True)
    layer.weight.data = out if out.device == src else out.to(src)


@torch.no_grad()
def gptq_hessian(input_feat: List[torch.Tensor], in_features: int, device, perp_damp: float = 0.01,
                 nsamples: int = 128) -> torch.Tensor:
    """Damped, normalised Hessian exactly as gptq_quantizer.py:133-150 defines it, fp32 [K,K] on
    `device`.  1-D features are rank-1 samples; non-tensor features give I (+ damping)."""
    from b200q import tensor_ops as _tops
    return _tops.gptq_hessian(input_feat, in_features, device, perp_damp, nsamples)


@torch.no_grad()
def gptq_inverse(H: torch.Tensor) -> torch.Tensor:
    """inv(H + 1e-6 I) for the SPD damped Hessian (reference: :160-165)."""
    from b200q import tensor_ops as _tops
    return _tops.spd_inverse(H, ridge=1e-6)


@torch.no_grad()
def _gptq_quantize_layer(
    layer: nn.Linear,
    n_bit: int,
    q_group_size: int,
    input_feat: List[torch.Tensor],
    perp_damp: float = 0.01,
    blocksize: int = 128,
    nsamples: int = 128,
    actorder: bool = False,
    verbose: bool = True,
) -> None:
    """GPTQ on one Linear (reference: gptq_quantizer.py:112-206)."""
    w = layer.weight.data
    src = w.device
    out = _gptq_device(_ops.to_device(w), n_bit, q_group_size, input_feat, perp_damp, blocksize,
                       nsamples, actorder)
    layer.weight.data = out if out.device == src else out.to(src)


def _prepare(input_feat, K: int, device, perp_damp: float, nsamples: int, actorder: bool,
             owner: int = 0):
    """(H, perm, factor) of one layer: the damped Hessian, the act-order permutation (compensated
    mode only) and what the column stage needs from the inverse -- H^-1 in parity mode (built
    like the reference builds it; its output does not depend on it), U = chol(H^-1) in compensated
    mode.  Under row sharding `owner` computes the factor; the caller broadcasts it."""
    from b200q import tensor_ops as _tops
    if MODE not in ("parity", "compensated"):
        raise ValueError(f"gptq_quantizer.MODE must be 'parity' or 'compensated', got {MODE!r}")
    if MODE == "parity" and not BUILD_HESSIAN:
        return None, None, None
    H = gptq_hessian(input_feat, K, device, perp_damp, nsamples)
    if MODE == "compensated":
        perm = torch.argsort(torch.diag(H), descending=True) if actorder else None
        return H, perm, _tops.compensation_factor(H, perm, owner=owner, broadcast=False)
    return H, None, _tops.spd_inverse(H, ridge=1e-6, owner=owner, broadcast=False)


def _gptq_device(W: torch.Tensor, n_bit: int, q_group_size: int, input_feat, perp_damp: float,
                 blocksize: int, nsamples: int, actorder: bool) -> torch.Tensor:
    """The per-layer stages on a CUDA-resident [N,K] weight; returns the quantized weight."""
    H, perm, factor = _prepare(input_feat, W.shape[1], W.device, perp_damp, nsamples, actorder)
    if factor is not None:
        _dist.broadcast(factor, 0)
    return _column_stage(W, n_bit, q_group_size, blocksize, H, perm, factor)


def _column_stage(W: torch.Tensor, n_bit: int, q_group_size: int, blocksize: int, H, perm,
                  factor) -> torch.Tensor:
    if MODE == "compensated":
        from b200q import tensor_ops as _tops
        out = _tops.gptq_compensated(W, None, n_bit, q_group_size, blocksize, perm, U=factor)
    elif MODE == "parity":
        # The reference's loop rounds column j with s_j = clamp(max_i |W[i,j]| / (2^b-1), 1e-5) and
        # writes q*s back; permuting and un-permuting independent columns is the identity.
        if _dist.is_sharded():
            # the column scale spans ALL rows (:182): combine the shards' column maxima first
            colmax = _dist.allreduce_max(_ops.col_absmax(W))
            out = _ops.gptq_parity_quant(W, n_bit, colmax)
        else:
            out = _ops.gptq_parity_layer(W, n_bit)          # both kernels behind one host call
    else:
        raise ValueError(f"gptq_quantizer.MODE must be 'parity' or 'compensated', got {MODE!r}")
    return out


# ==================================================================================================
# calibration capture
# ==================================================================================================
@torch.no_grad()
def gptq_calibrate_hessian(
    model: nn.Module,
    calib_samples: List[torch.Tensor],
    nsamples: int = 128,
    verbose: bool = True,
) -> Dict[str, List[torch.Tensor]]:
    """Capture, per Linear, the [tokens, in_features] input of every calibration batch, kept on the
    device it was produced on (reference: gptq_quantizer.py:210-264).  These lists are the 2-D
    `input_feat` layout `_gptq_quantize_layer` turns into H with the tensor-core kernel."""
    import tqdm

    captured: Dict[str, List[torch.Tensor]] = {}

    def make_hook(name: str):
        def hook(_m, inputs, _out):
            x = inputs[0] if isinstance(inputs, tuple) else inputs
            if x.dim() > 2:
                x = x.reshape(-1, x.shape[-1])
            captured.setdefault(name, []).append(x.detach())
        return hook

    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if verbose:
        print("Pre-computing GPTQ Hessian matrices...")
    handles = [m.register_forward_hook(make_hook(n)) for n, m in model.named_modules()
               if isinstance(m, nn.Linear)]
    try:
        for sample in tqdm.tqdm(calib_samples[:nsamples], disable=not verbose,
                                desc="hessian calibration"):
            model(sample.to(device))
    finally:
        for h in handles:
            h.remove()
    return captured
