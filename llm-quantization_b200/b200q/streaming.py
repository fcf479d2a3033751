"""On-device streaming calibration capture.

The reference's hooks keep every calibration batch: `gptq_calibrate_hessian` stores the full
[tokens, in_features] input of every Linear for every batch (gptq_quantizer.py:236-247, 2-7 GB per
Linear at Llama-7B sizes) and builds H from the list afterwards.  An `ActivationStream` folds each
batch into the running statistics the quantizers actually consume as soon as the hook sees it:

    * the GPTQ Hessian term  x^T x / (||x|| + 1e-5)^2      (normalize=True,  gptq_quantizer.py:142-148)
      or the plain Gram term x^T x for the AWQ search       (normalize=False)
    * the per-batch mean|x| row AWQ ranks channels by        (quantization_utils.py:231)

so the activations themselves are never kept.  A stream can be put wherever the drop-in modules
accept a feature list: `gptq_quantize_model_weight`, `awq_search_scale_factor` and
`awq_quantize_model_weight` recognise it.  Under `b200q.dist.row_sharded()` batches are dealt to
the ranks round-robin and the sums all-reduced when the matrix is asked for.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import dist as _dist
from . import ops as _ops


class ActivationStream:
    def __init__(self, in_features: int, normalize: bool, max_batches: Optional[int] = None,
                 keep_stats: bool = True):
        self.in_features = int(in_features)
        self.normalize = bool(normalize)
        self.max_batches = max_batches      # batches beyond this are counted but not added (:140 `[:nsamples]`)
        self.keep_stats = keep_stats
        self.batches_seen = 0
        self.rows_added = 0
        self.H: Optional[torch.Tensor] = None
        self.stat_rows: List[torch.Tensor] = []

    def add(self, x: torch.Tensor) -> None:
        """Fold one calibration batch ([..., in_features], any float dtype, any device) in."""
        from . import tensor_ops as _tops
        index = self.batches_seen
        self.batches_seen += 1
        x = _ops.to_device(x.detach().reshape(-1, self.in_features))
        if x.dtype not in _ops.DTYPE_CODE:
            x = x.float()
        if self.keep_stats:
            self.stat_rows.append(_ops.act_meanabs(x).to(x.dtype))
        if self.max_batches is not None and index >= self.max_batches:
            return
        if _dist.is_sharded() and index % _dist.world_size() != _dist.rank():
            return
        self.H = _tops.hessian_accum(x, x.shape[0], self.H, normalize=self.normalize)
        self.rows_added += x.shape[0]

    # ---- what the quantizers ask for ----------------------------------------------------------
    def matrix_sum(self, device) -> torch.Tensor:
        """The running sum over all ranks' batches (a fresh tensor; the stream can keep growing)."""
        K = self.in_features
        H = self.H.clone() if self.H is not None else \
            torch.zeros((K, K), dtype=torch.float32, device=torch.device(device))
        from . import tensor_ops as _tops
        return _tops.allreduce_symmetric(H)

    def total_rows(self, device) -> int:
        if not _dist.is_sharded():
            return self.rows_added
        n = torch.tensor([self.rows_added], dtype=torch.int64, device=torch.device(device))
        return int(_dist.allreduce_sum(n).item())

    def stats_matrix(self, device) -> torch.Tensor:
        """[batches, K] per-batch mean|x| rows (every rank sees every batch's row)."""
        assert self.keep_stats and self.stat_rows, "this stream was created without statistics"
        return torch.stack([r.to(device) for r in self.stat_rows])
