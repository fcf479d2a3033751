"""smooth_quant_quantizer — drop-in for the reference module of the same name.

y = W x = (W diag(s)^-1)(diag(s) x) with s_k = max|x_k|^alpha / max_i|W_ik|^(1-alpha)
(reference: smooth_quant_quantizer.py:112-199).  On the B200:

  * the per-channel activation |max| of `collect_act_scales` is the b200q act_maxabs kernel,
  * the weight column |max|, the scale vector, the migration W / s and — in
    `smoothquant_quantize_model_weight` — the group fake-quant that follows are fused into
    col_absmax -> smooth_scale -> ONE group_fakequant pass with the DIV column op.

`m.smoothing_scale` and the forward-pre-hook that multiplies the inputs by s are installed exactly
as the reference does, so a model evaluated after the call computes the same function.
`smoothquant_search_alpha` is a stub in the reference (returns the midpoint); here it evaluates the
weight-side reconstruction error per alpha with the same kernels (SEARCH_STUB restores the stub).
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from b200q import ops as _ops  # noqa: E402
from b200q import dist as _dist  # noqa: E402
from b200q import pipeline as _pipeline  # noqa: E402
from quantization_utils import pseudo_quantize_tensor  # noqa: E402,F401

SEARCH_STUB = False


# ==================================================================================================
# calibration
# ==================================================================================================
@torch.no_grad()
def collect_act_scales(model: nn.Module, calib_samples: List[torch.Tensor],
                       verbose: bool = True) -> Dict[str, torch.Tensor]:
    """Running per-channel max|x| of every Linear's input over the calibration batches, returned
    as CPU tensors in the activations' dtype (reference: :33-104)."""
    import tqdm

    running: Dict[str, torch.Tensor] = {}   # fp32 on the device while collecting
    dtypes: Dict[str, torch.dtype] = {}

    def make_hook(name: str):
        def hook(_m, inputs, _out):
            x = inputs[0] if isinstance(inputs, tuple) else inputs
            if x.dim() == 1:
                x = x.unsqueeze(0)
            x = _ops.to_device(x.detach())
            if name in running:
                _ops.act_maxabs(x, out=running[name])
            else:
                running[name] = _ops.act_maxabs(x)
                dtypes[name] = x.dtype
        return hook

    handles = [m.register_forward_hook(make_hook(n)) for n, m in model.named_modules()
               if isinstance(m, nn.Linear)]
    model_device = next(model.parameters()).device
    if verbose:
        print("Collecting activation scales from calibration data...")
    try:
        for input_ids in tqdm.tqdm(calib_samples, disable=not verbose, desc="collecting act scales"):
            with torch.no_grad():
                model(input_ids.to(model_device))
    finally:
        for h in handles:
            h.remove()
    if verbose:
        print(f"  -> Collected scales for {len(running)} layers")
    return {n: v.to(dtypes[n]).cpu() for n, v in running.items()}


# ==================================================================================================
# smoothing
# ==================================================================================================
def _scale_activations_hook(mod: nn.Module, inputs):
    """forward-pre-hook: x <- x * s, keeping (W/s)(s x) = W x."""
    s = getattr(mod, "smoothing_scale", None)
    if s is None:
        return None
    x = inputs[0] if isinstance(inputs, tuple) else inputs
    x = x * s.to(x.device, dtype=x.dtype)
    return (x,) + tuple(inputs[1:]) if isinstance(inputs, tuple) else (x,)


def _layer_smoothing_scale(W: torch.Tensor, act_scale: torch.Tensor, alpha: float, colmax=None):
    """(s as fp32 on W's device, dtype s has in torch's promotion rules).  `colmax` = the column
    |max| of W over all row shards when the caller already has it (the alpha sweep)."""
    if colmax is None:
        colmax = _dist.allreduce_max(_ops.col_absmax(W))
    s = _ops.smooth_scale(act_scale, colmax, alpha, act_scale.dtype if act_scale.dtype in
                          _ops.DTYPE_CODE else torch.float32, W.dtype)
    return s, torch.promote_types(act_scale.dtype, W.dtype)


def _attach(m: nn.Linear, s: torch.Tensor) -> None:
    m.smoothing_scale = s.detach()
    if getattr(m, "_smooth_pre_hook_handle", None) is None:
        m._smooth_pre_hook_handle = m.register_forward_pre_hook(_scale_activations_hook)


@torch.no_grad()
def smooth_weights(model: nn.Module, act_scales: Dict[str, torch.Tensor], alpha: float = 0.5,
                   verbose: bool = True) -> None:
    """W <- W / s per input channel for every Linear present in `act_scales`."""
    if verbose:
        print(f"Applying weight smoothing with alpha={alpha}...")
    for name, m in model.named_modules():
        if not isinstance(m, nn.Linear):
            continue
        if name not in act_scales:
            if verbose:
                print(f"  warning: {name} not in act_scales, skipping")
            continue
        src = m.weight.device
        W = _ops.to_device(m.weight.data)
        s, s_dtype = _layer_smoothing_scale(W, act_scales[name], alpha)
        out = _ops.col_scale(W.to(s_dtype), s)
        m.weight.data = out if out.device == src else out.to(src)
        _attach(m, s.to(s_dtype).to(src))


@torch.no_grad()
def smooth_activations(model: nn.Module, calib_samples: List[torch.Tensor], alpha: float = 0.5,
                       verbose: bool = True) -> None:
    """No-op kept for API parity: the activation side is applied by the forward-pre-hook."""
    if verbose:
        print(f"Activation smoothing (inverse transformation) - alpha={alpha}")
        print("  note: in practice, this is fused into next layer's weights")


@torch.no_grad()
def reverse_weight_smoothing(model: nn.Module, verbose: bool = True) -> None:
    """W <- W * s and remove the hook (reference: :230-260)."""
    if verbose:
        print("Reversing weight smoothing...")
    for _name, m in model.named_modules():
        if isinstance(m, nn.Linear) and hasattr(m, "smoothing_scale"):
            src = m.weight.device
            W = _ops.to_device(m.weight.data)
            s = m.smoothing_scale
            out = _ops.col_scale(W.to(torch.promote_types(W.dtype, s.dtype)), s, mul=True)
            m.weight.data = out if out.device == src else out.to(src)
            del m.smoothing_scale
            handle = getattr(m, "_smooth_pre_hook_handle", None)
            if handle is not None:
                try:
                    handle.remove()
                finally:
                    m._smooth_pre_hook_handle = None


# ==================================================================================================
# quantization
# ==================================================================================================
@torch.no_grad()
def smoothquant_quantize_model_weight(model: nn.Module, w_bit: int, q_group_size: int,
                                      act_scales: Dict[str, torch.Tensor], alpha: float = 0.5,
                                      verbose: bool = True) -> None:
    """Smooth every Linear found in `act_scales`, then fake-quantize EVERY Linear (reference:
    :268-323).  Smoothing and quantization of a layer are one pass over its weights."""
    if verbose:
        print(f"Applying SmoothQuant quantization (w_bit={w_bit}, alpha={alpha})...")
        print(f"Applying weight smoothing with alpha={alpha}...")
    def compute(name, m, W):
        if q_group_size > 0:
            assert W.shape[-1] % q_group_size == 0
        if name in act_scales:
            act = act_scales[name]
            if not _dist.is_sharded() and torch.promote_types(act.dtype, W.dtype) == W.dtype \
                    and act.dtype in _ops.DTYPE_CODE:
                # column |max| -> s -> fused (W / s, group fake-quant): one host call
                out, s = _ops.smoothquant_layer(W, act, alpha, w_bit, q_group_size)
                _attach(m, s.to(W.dtype).to(m.weight.device))
                return out
            s, s_dtype = _layer_smoothing_scale(W, act, alpha)
            _attach(m, s.to(s_dtype).to(m.weight.device))
            return _ops.group_fakequant(W.to(s_dtype), w_bit, q_group_size, colop=_ops.COLOP_DIV,
                                        colvec=s)
        if verbose:
            print(f"  warning: {name} not in act_scales, skipping")
        return _ops.group_fakequant(W, w_bit, q_group_size)

    _pipeline.run_layers([(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)],
                         compute)
    if verbose:
        print("  Quantizing weights...")
        print("  Done! (Activation scaling is fused into next layer in real inference)")


@torch.no_grad()
def smoothquant_search_alpha(model: nn.Module, calib_samples: List[torch.Tensor],
                             act_scales: Dict[str, torch.Tensor], w_bit: int = 8,
                             q_group_size: int = -1, alpha_range: Tuple[float, float] = (0.0, 1.0),
                             n_grid: int = 20, verbose: bool = True) -> float:
    """Grid-search alpha.  The reference returns the midpoint without measuring anything
    (:363-371); this evaluates, for each of `n_grid` alphas, the reconstruction error of the
    smoothed-then-quantized weights mapped back to the original basis,
        sum_layers || (Q(W / s) * s - W) diag(a) ||_F^2 ,   a = calibration max|x| per channel,
    and returns the minimiser.  The model is not modified.  (PARITY UNPINNED: no reference body.)"""
    if verbose:
        print("Searching for optimal alpha value...")
    lo, hi = alpha_range
    if SEARCH_STUB:
        best = (lo + hi) / 2.0
    else:
        alphas = torch.linspace(float(lo), float(hi), int(n_grid), dtype=torch.float64).tolist()
        totals = None
        for name, m in model.named_modules():
            if not isinstance(m, nn.Linear) or name not in act_scales:
                continue
            W = _ops.to_device(m.weight.data)
            a = act_scales[name].to(W.device, torch.float32).clamp(min=1e-5)
            # one smoothing-scale vector per alpha ([n_grid, K], tiny), then ONE kernel sweeps all
            # alphas over the weight: W is read once, the errors accumulate on the device
            colmax = _dist.allreduce_max(_ops.col_absmax(W))
            scales = [_layer_smoothing_scale(W, act_scales[name], alpha, colmax) for alpha in alphas]
            S = torch.stack([s for s, _ in scales])
            totals = _ops.smooth_alpha_errors(W.to(scales[0][1]), S, a, w_bit, q_group_size, totals)
        if totals is not None:
            totals = _dist.allreduce_sum(totals)
        best = alphas[int(torch.argmin(totals).item())] if totals is not None and \
            float(totals.sum().item()) > 0 else (lo + hi) / 2.0
    if verbose:
        print(f"  -> Using alpha: {best:.2f}")
    return float(best)


@torch.no_grad()
def smoothquant_quantize_and_calibrate(model: nn.Module, w_bit: int, q_group_size: int,
                                       calib_samples: List[torch.Tensor],
                                       alpha: Optional[float] = None, search_alpha: bool = False,
                                       verbose: bool = True) -> Dict[str, torch.Tensor]:
    """collect_act_scales -> (optional) alpha search -> smoothquant_quantize_model_weight."""
    act_scales = collect_act_scales(model, calib_samples, verbose)
    if search_alpha:
        alpha = smoothquant_search_alpha(model, calib_samples, act_scales, w_bit, q_group_size,
                                         verbose=verbose)
    elif alpha is None:
        alpha = 0.5
    smoothquant_quantize_model_weight(model, w_bit, q_group_size, act_scales, alpha, verbose)
    return act_scales
