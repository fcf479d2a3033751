"""quantization_utils — drop-in for the reference module of the same name.

In scope here: `pseudo_quantize_tensor` and the activation-statistics hook
of `get_calib_feat`, both running as sm_100a kernels through libb200quant.  The small host helpers
(`load_config`, `get_model_size`, ...) are re-stated because callers import them from this module.
Model/dataset loading and perplexity evaluation are outside the hot path: they are resolved lazily
from a reference checkout (env LLMQ_REFERENCE_DIR) and fail with a clear message if none is present.
"""
from __future__ import annotations

import gc
import importlib.util
import json
import sys
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from b200q import ops as _ops  # noqa: E402

# size units, in bits (reference: quantization_utils.py:38-41)
Byte = 8
KiB = 1024 * Byte
MiB = 1024 * KiB
GiB = 1024 * MiB


# ==================================================================================================
# hot path
# ==================================================================================================
@torch.no_grad()
def pseudo_quantize_tensor(w: Tensor, n_bit: int = 4, q_group_size: int = -1) -> Tensor:
    """Asymmetric min/max uniform fake-quantization per group of `q_group_size` elements of the
    last dimension (per row when q_group_size <= 0).

    Same contract as the reference (quantization_utils.py:362-413): new tensor, input's shape and
    dtype, AssertionError when the last dim is not divisible by the group size or the tensor is not
    2-D in per-row mode.  Results are bit-identical to the reference's ON CPU TENSORS (fp32, fp16 and
    bf16): torch's CPU kernels divide by the Python scalar `max_int` with a true division, which is
    what the kernel does; torch's CUDA kernels turn `x / scalar` into `x * (1 / scalar)`, so the
    reference run on a GPU can differ from both in the last bit of a scale.
    """
    if q_group_size > 0:
        assert w.shape[-1] % q_group_size == 0
    else:
        assert w.dim() == 2
    src = w.device
    wd = _ops.to_device(w)
    out = _ops.group_fakequant(wd, n_bit, q_group_size)
    # The reference asserts that neither the scales nor the result contain NaN (:398-399, :407).
    # The kernel's min/max and clamps drop NaNs (a NaN weight would come out as a finite grid
    # point), so the contract is kept by testing the INPUT: NaN in <=> NaN scales in the reference.
    assert torch.isnan(wd).sum() == 0
    return out if src == out.device else out.to(src)


CALIB_ON_DEVICE = False   # True: get_calib_feat returns one CUDA [n_batches, C] matrix per Linear


def get_calib_feat(model: nn.Module, tokenizer: Any, calib_samples: List[Tensor],
                   verbose: bool = True) -> Dict[str, List[Tensor]]:
    """Per-Linear list of mean|x| vectors, one per calibration batch (reference:
    quantization_utils.py:204-262).  The reduction over tokens runs in b200q's act_meanabs kernel.

    The reference's hook moves every vector to the host as it is produced (`.cpu()`, :231): one
    device synchronisation per Linear per batch.  Here the rows stay on the device while the
    calibration batches run and each layer's [n_batches, C] matrix crosses
    to the host ONCE at the end; the returned dict has the reference's layout (lists of CPU [C]
    tensors in the activations' dtype).  With CALIB_ON_DEVICE = True nothing is copied back: the
    values are CUDA [n_batches, C] matrices, which the AWQ / GPTQ walkers accept wherever they
    accept the lists (a matrix iterates and sums row by row like the list)."""
    import tqdm

    rows: Dict[str, List[Tensor]] = {}

    def make_hook(name: str):
        def hook(_m, inputs, _out):
            x = inputs[0] if isinstance(inputs, tuple) else inputs
            rows.setdefault(name, []).append(_ops.act_meanabs(_ops.to_device(x.detach())).to(x.dtype))
        return hook

    handles = [m.register_forward_hook(make_hook(n)) for n, m in model.named_modules()
               if isinstance(m, nn.Linear)]
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if verbose:
        print("Collecting activation scales from calibration data...")
    try:
        for input_ids in tqdm.tqdm(calib_samples, disable=not verbose, desc="calibration"):
            with torch.no_grad():
                model(input_ids.to(device))
    finally:
        for h in handles:
            h.remove()
    if CALIB_ON_DEVICE:
        return {n: torch.stack(v) for n, v in rows.items()}
    return {n: list(torch.stack(v).cpu().unbind(0)) for n, v in rows.items()}


# ==================================================================================================
# small host helpers
# ==================================================================================================
def load_config(config_path: str) -> Dict[str, Any]:
    with open(config_path, "r") as fh:
        return json.load(fh)


def save_config(config: Dict[str, Any], config_path: str) -> None:
    with open(config_path, "w") as fh:
        json.dump(config, fh, indent=2)


def unload_model(model: Optional[nn.Module] = None) -> None:
    del model
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def get_model_size(model: nn.Module, data_width: int = 16, group_size: int = -1,
                   use_zero_point: bool = True) -> int:
    """Model size in bits: parameters x (data_width + per-group fp16 scale [+ 4-bit zero])."""
    bits = data_width
    if group_size != -1:
        bits += 16 / group_size
        if use_zero_point:
            bits += 4 / group_size
    return sum(p.numel() for p in model.parameters()) * bits


def get_linear_layers(model: nn.Module) -> List[Tuple[str, nn.Linear]]:
    return [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear)]


# ==================================================================================================
# out-of-scope pass-throughs (HF model / dataset I/O, perplexity evaluation)
# ==================================================================================================
_PASSTHROUGH = ("load_model_and_tokenizer", "get_calibration_dataset", "get_test_dataset",
                "evaluate_perplexity")
_ref_module = None


def _reference_utils():
    global _ref_module
    if _ref_module is None:
        from b200q.build import reference_dir
        root = reference_dir()
        if root is None:
            raise ImportError(
                f"{', '.join(_PASSTHROUGH)} are not part of the B200 hot path and are taken from a "
                f"reference checkout; none found (set LLMQ_REFERENCE_DIR)")
        path = root / "quantization_utils.py"
        spec = importlib.util.spec_from_file_location("_llmq_reference_quantization_utils", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_module = mod
    return _ref_module


def __getattr__(name: str):
    if name in _PASSTHROUGH:
        return getattr(_reference_utils(), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
