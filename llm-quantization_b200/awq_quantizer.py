"""awq_quantizer — drop-in for the reference module of the same name.

`awq_quantize_model_weight` keeps the reference's semantics exactly (awq_quantizer.py:22-84):
importance = Python-sum of the per-batch mean|x| vectors, top-k salient input channels, scale them
up, asymmetric group fake-quant, scale them back.  On the B200 the three weight passes are ONE
kernel (b200q_group_fakequant with the MUL_DIV column op) and the importance sum is one kernel that
reproduces sum()'s left-to-right order.

`awq_search_scale_factor` is a stub in the reference (returns the midpoint).  Here it performs the
search its docstring describes (awq_quantizer.py:116-119) with the fused quantize + output-MSE
tensor-core kernel; set SEARCH_STUB = True to get the reference's midpoint back.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

_HERE = Path(__file__).resolve().parent
if str(_HERE) not in sys.path:
    sys.path.insert(0, str(_HERE))

from b200q import ops as _ops  # noqa: E402
from b200q import pipeline as _pipeline  # noqa: E402
from quantization_utils import pseudo_quantize_tensor  # noqa: E402,F401  (re-exported like the reference)

SEARCH_STUB = False
LOOKAHEAD = 2      # under row sharding: Gram matrices (and their exchanges) started ahead of the search
# What the search runs on a (high-priority) side stream WHILE the Gram GEMM executes: 0 nothing,
# 1 the activation statistics + salient-channel selection, 2 also the candidate quantisation (dW_c
# for every c).  Measured on one B200 (Llama-2-7B shapes, same box, interleaved): the kernels do
# overlap (statistics 111 -> 221 ms, candidates 205 -> 252 ms elapsed) but the GEMM stretches from
# 1.55 s to 1.93 s -- the step sits on the 1 kW power cap, so concurrent work just lowers the clock
# -- and the step does not get shorter (2.60 s vs 2.63 s).  Hence off by default.
SEARCH_OVERLAP = int(__import__("os").environ.get("B200Q_AWQ_OVERLAP", "0"))


def _feat_matrix(feats, device) -> torch.Tensor:
    """The per-batch statistics of one layer as an [n, K] matrix on `device`.  The reference takes
    a list of [K] tensors; an [n, K] tensor iterates (and sums) row by row and is accepted as is."""
    from b200q.streaming import ActivationStream
    if isinstance(feats, ActivationStream):
        stacked = feats.stats_matrix(device)
    elif isinstance(feats, torch.Tensor):
        stacked = feats.reshape(feats.shape[0], -1)
    else:
        stacked = torch.stack([f.reshape(-1) for f in feats])
    if stacked.dtype not in _ops.DTYPE_CODE:
        stacked = stacked.float()
    return stacked.to(device, non_blocking=True)


def _stat_rows(feats: List[torch.Tensor], device) -> torch.Tensor:
    """[n, K] matrix of per-batch mean|x| rows from a feature list whose entries are either such
    vectors already (1-D) or raw [tokens, K] activations (reduced by the act_meanabs kernel)."""
    if isinstance(feats, torch.Tensor) and feats.dim() == 3:
        # [n, tokens, K] batch of raw activations: one batched reduction (under row sharding the
        # batches are dealt to the ranks and the [n, K] rows all-gathered)
        from b200q import dist as _dist
        x = _ops.to_device(feats)
        return _dist.gather_rows(lambda lo, hi: _ops.act_meanabs_batched(x[lo:hi]), x.shape[0],
                                 (x.shape[2],), torch.float32, x.device).to(feats.dtype)
    rows = []
    for f in feats:
        if f.dim() == 1:
            rows.append(f.to(device))
        else:
            rows.append(_ops.act_meanabs(_ops.to_device(f)).to(f.dtype))
    return torch.stack(rows)


_side_streams = {}


def _side_stream(device) -> torch.cuda.Stream:
    s = _side_streams.get(device.index)
    if s is None:
        # high priority: the block scheduler hands SM slots to this stream's (small) kernels ahead
        # of the thousands of pending GEMM blocks, otherwise they would only start at the GEMM's tail
        s = _side_streams[device.index] = torch.cuda.Stream(device, priority=-1)
    return s


def _importance(feats, device) -> torch.Tensor:
    """sum(list of [K] tensors).float(), evaluated left to right in the tensors' own dtype."""
    return _ops.seq_sum_rows(_feat_matrix(feats, device))


def _salient_channels(importance: torch.Tensor, protect_ratio: float) -> torch.Tensor:
    n_protect = max(1, int(importance.numel() * protect_ratio))
    return torch.topk(importance, n_protect)[1]


@torch.no_grad()
def awq_quantize_model_weight(
    model: nn.Module,
    w_bit: int,
    q_group_size: int,
    input_feat: Dict[str, List[torch.Tensor]],
    protect_ratio: float = 0.01,
    scale_factor: float = 1.0,
) -> None:
    """AWQ-protect and fake-quantize every calibrated nn.Linear in place; Linears without
    calibration features are left untouched (reference: awq_quantizer.py:50-54)."""
    def compute(name, _module, W):
        K = W.shape[-1]
        if q_group_size > 0:
            assert K % q_group_size == 0
        # importance sum, top-k and the fused scale/quantize/unscale pass: one host call
        n_protect = max(1, int(K * protect_ratio))
        return _ops.awq_layer(W, _feat_matrix(input_feat[name], W.device), w_bit, q_group_size,
                              n_protect, scale_factor)

    _pipeline.run_layers([(n, m) for n, m in model.named_modules()
                          if isinstance(m, nn.Linear) and n in input_feat], compute)


@torch.no_grad()
def awq_search_scale_factor(
    model: nn.Module,
    w_bit: int,
    q_group_size: int,
    input_feat: Dict[str, List[torch.Tensor]],
    protect_ratio: float = 0.01,
    scale_search_range: Tuple[float, float] = (1.0, 2.0),
    n_grid: int = 20,
) -> float:
    """Grid-search the AWQ scale factor: for each of `n_grid` candidates in `scale_search_range`
    quantize every calibrated Linear and add up the output reconstruction error
    ||(Q(W) - W) X^T||^2 against the calibration features; return the candidate with the smallest
    total.  The model is not modified."""
    print("Searching for optimal scale factor...")
    lo, hi = scale_search_range
    if SEARCH_STUB:
        best = (lo + hi) / 2.0
        print(f"  -> Using scale factor: {best:.3f}")
        return best
    from b200q import tensor_ops as _tops
    from b200q import dist as _dist
    candidates = torch.linspace(float(lo), float(hi), int(n_grid), dtype=torch.float64).tolist()
    totals = []
    items = [(n, m) for n, m in model.named_modules()
             if isinstance(m, nn.Linear) and n in input_feat]
    index = {n: i for i, (n, _) in enumerate(items)}
    grams = {}

    def supported(K: int) -> bool:
        # TMA needs a 16-byte row pitch (K % 8); a quantisation group must divide the row
        return K % 8 == 0 and not (0 < q_group_size < K and K % q_group_size != 0)

    def start_gram(i, device):
        n, m = items[i]
        if supported(m.weight.shape[1]):
            grams[n] = _tops.gram_matrix_begin(input_feat[n], m.weight.shape[1], device, want_folded=True)

    skipped = []

    def compute(name, _module, W):
        K = W.shape[1]
        feats = input_feat[name]
        # Shapes the kernels cannot take do not raise here: the reference's search returns a float for ANY model
        # (awq_quantizer.py:88-126), so such a layer is left out of the total with a warning.
        if not supported(K):
            skipped.append(name)
            return None
        # The Gram GEMM of this layer (and, under row sharding, of the NEXT one, whose exchange then
        # hides behind this layer's search) is queued first; with SEARCH_OVERLAP the per-batch
        # mean|x| statistics, the salient-channel selection and the candidate quantisation run on a
        # side stream while it executes (see SEARCH_OVERLAP for why that is not the default).
        main = torch.cuda.current_stream(W.device)
        before = torch.cuda.Event()
        before.record(main)
        i = index[name]
        if name not in grams:
            start_gram(i, W.device)
        if _dist.is_sharded():
            # two layers ahead: the exchange of layer i+1 (reduce-scatter, fold, all-gather on the
            # communication stream) then has the Gram GEMM of layer i+2 AND this layer's search to
            # hide behind -- at 8 GPUs one search alone (~0.5 ms) is shorter than the exchange
            for j in range(i + 1, min(i + 1 + LOOKAHEAD, len(items))):
                if items[j][0] not in grams:
                    start_gram(j, W.device)
        side = _side_stream(W.device) if SEARCH_OVERLAP >= 1 else main
        side.wait_event(before)
        with torch.cuda.stream(side):
            # 2-D [tokens, K] features are raw activations: their per-batch mean|x| is the statistic
            # the quantizer ranks channels by (quantization_utils.py:231); 1-D features already are it
            from b200q.streaming import ActivationStream
            if isinstance(feats, ActivationStream) or (isinstance(feats, torch.Tensor) and feats.dim() == 2):
                stat_rows = feats
            else:
                stat_rows = _stat_rows(feats, W.device)
            n_protect = max(1, int(K * protect_ratio))
            mask = _ops.salient_mask(_feat_matrix(stat_rows, W.device), n_protect)
            # ... and so does the candidate quantisation (dW_c = Q_c(W) - W for every candidate): it
            # needs W and the mask only and is FP32-issue bound, i.e. it uses what the GEMM leaves idle
            prepared = None
            if len(candidates) <= 32 and SEARCH_OVERLAP >= 2:
                W.record_stream(side)
                prepared = _tops.awq_search_prepare(W, mask, w_bit, q_group_size, candidates)
            selected = torch.cuda.Event()
            selected.record(side)
        mask.record_stream(main)
        main.wait_event(selected)
        # the loss is linear in H: search in the plain sum X^T X and divide the n_grid losses by
        # the row count instead of rescaling the [K, K] matrix
        gram = grams.pop(name)
        H = _tops.gram_matrix_end(gram, normalise=False)
        if prepared is not None:
            loss = _tops.awq_search_finish(prepared, H)
        else:
            # (the kernel takes up to 32 candidates per call: longer grids go in slices)
            loss = torch.cat([_tops.awq_search_losses(W, H, mask, w_bit, q_group_size, candidates[c:c + 32])
                              for c in range(0, len(candidates), 32)])
        loss.mul_(1.0 / gram.rows_total)
        if totals:
            totals[0] += loss
        else:
            totals.append(loss)
        return None                       # the search leaves the model untouched

    # host-resident weights are prefetched one layer ahead; nothing is written back
    _pipeline.run_layers(items, compute)
    if skipped:
        import warnings
        warnings.warn(f"awq_search_scale_factor: {len(skipped)} layer(s) left out of the search "
                      f"(in_features not a multiple of 8 or of the group size): {skipped[:4]}")
    if not totals:
        best = (lo + hi) / 2.0
    else:
        total = _dist.allreduce_sum(totals[0])
        best = candidates[int(torch.argmin(total).item())]
    print(f"  -> Using scale factor: {best:.3f}")
    return float(best)
