"""pot_apot_quantizer — drop-in for the reference module of the same name.

POT:  w_q = s * sign(w) * 2^E, E in [0, 2^(b-1)-1], per-group scale s = s0 * b searched over the
      200-point grid torch.arange(0.01, 2.01, 0.01)            (reference: pot_apot_quantizer.py:25-115)
APOT: w_q = s * level, level from the additive power-of-two set, s = max|w| * b searched over a
      20- or 40-point grid chosen from the tensor's element count      (reference: :192-351)

Both searches run entirely inside one kernel per tensor (b200q_pot_quant / b200q_apot_quant): a
group stays in registers while every candidate is evaluated, the strict first-minimum rule and
torch's summation order are reproduced, and the level selection is bit-identical to the reference.
The grids and the level table are materialised on the host with the same torch calls the reference
uses and handed to the kernel as data.
"""
from __future__ import annotations

import itertools
import sys
from pathlib import Path

import torch
import torch.nn as nn

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from b200q import ops as _ops  # noqa: E402
from b200q import dist as _dist  # noqa: E402
from b200q import pipeline as _pipeline  # noqa: E402


def _as_groups(w: torch.Tensor, q_group_size: int) -> torch.Tensor:
    if q_group_size > 0:
        assert w.shape[-1] % q_group_size == 0
        w = w.reshape(-1, q_group_size)
    assert w.dim() == 2
    return w


# ==================================================================================================
# POT
# ==================================================================================================
@torch.no_grad()
def pot_quantize_tensor(w: torch.Tensor, n_bit: int = 4, q_group_size: int = -1) -> torch.Tensor:
    """Power-of-two fake-quantization; returns a new tensor with w's shape and dtype."""
    shape, src = w.shape, w.device
    groups = _as_groups(_ops.to_device(w), q_group_size)
    grid = torch.arange(0.01, 2.01, 0.01)          # host tensor, as the reference builds it (:75)
    out = _ops.pot_quant(groups, n_bit, grid).reshape(shape)
    assert torch.isnan(out).sum() == 0
    return out if out.device == src else out.to(src)


@torch.no_grad()
def pot_quantize_model_weight(model: nn.Module, w_bit: int, q_group_size: int) -> None:
    _pipeline.run_layers([(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)],
                         lambda _n, _m, W: pot_quantize_tensor(W, n_bit=w_bit,
                                                               q_group_size=q_group_size))


# ==================================================================================================
# APOT
# ==================================================================================================
def generate_apot_levels(n: int, k: int, device: torch.device = torch.device("cpu")) -> torch.Tensor:
    """Sorted unique sums of n terms, term i drawn from {0} U {2^-(i + j*n) : j = 0 .. 2^k - 2}
    (reference: :138-188).  A host-side table; fp32."""
    choices = [[0.0] + [2.0 ** -(i + j * n) for j in range(2 ** k - 1)] for i in range(n)]
    sums = torch.tensor([sum(c) for c in itertools.product(*choices)], dtype=torch.float32,
                        device=device)
    return torch.sort(torch.unique(sums))[0]


def _apot_signed_levels(n_bit: int, k: int) -> torch.Tensor:
    """{-levels, 0, +levels} normalised to max 1, thinned to 32 entries when larger (:224-247)."""
    levels = generate_apot_levels(max(1, n_bit // k), k)
    top = levels.max()
    if top > 0:
        levels = levels / top
    pos = levels[levels > 0]
    full = torch.cat([-pos.flip(0), torch.zeros(1), pos])
    if full.numel() > 32:
        full = full[torch.linspace(0, full.numel() - 1, 32, dtype=torch.long)]
    return full


@torch.no_grad()
def apot_quantize_tensor(w: torch.Tensor, n_bit: int = 4, q_group_size: int = -1,
                         k: int = 2) -> torch.Tensor:
    """Additive power-of-two fake-quantization; returns a new tensor with w's shape and dtype."""
    shape, src = w.shape, w.device
    groups = _as_groups(_ops.to_device(w), q_group_size)
    # the grid depends on the element count of the WHOLE tensor (:258-262); under row sharding
    # that is the sum over ranks, not this shard's
    total = _dist.global_numel(groups.numel(), groups.device)
    grid = torch.arange(0.01, 2.01, 0.1 if total > 500000 else 0.05)
    out = _ops.apot_quant(groups, _apot_signed_levels(n_bit, k), grid).reshape(shape)
    assert torch.isnan(out).sum() == 0
    return out if out.device == src else out.to(src)


@torch.no_grad()
def apot_quantize_model_weight(model: nn.Module, w_bit: int, q_group_size: int, k: int = 2) -> None:
    _pipeline.run_layers([(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)],
                         lambda _n, _m, W: apot_quantize_tensor(W, n_bit=w_bit,
                                                                q_group_size=q_group_size, k=k))
