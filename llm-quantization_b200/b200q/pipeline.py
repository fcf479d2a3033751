"""Layer streaming for the model walkers.

The walkers visit nn.Linear modules one at a time.  When the weights already live on the GPU the
per-layer kernels run back to back on the current stream.  When they live on the HOST (the
reference's tests and CPU-offloaded models) the walker would otherwise serialise
copy-in / compute / copy-out per layer; `run_layers` overlaps them instead:

    h2d stream :  W[i+1] host -> device            (prefetch, one layer ahead)
    compute    :  quantize W[i]                     (current stream)
    d2h stream :  out[i-1] device -> host

Pinned host tensors make both copies asynchronous; pageable ones still work (torch stages them).
Results are written back INTO the module's host tensor when shape and dtype allow, so no host
allocation happens per layer.  All arithmetic stays on the GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops as _ops

Compute = Callable[[str, nn.Linear, torch.Tensor], Optional[torch.Tensor]]


def _assign(module: nn.Linear, out: torch.Tensor) -> None:
    module.weight.data = out


def run_layers(items: Sequence[Tuple[str, nn.Linear]], compute: Compute) -> None:
    """For every (name, linear): out = compute(name, linear, W_on_gpu); linear.weight.data <- out
    on the weight's original device.  `compute` must launch on the current stream and may return
    None to leave the layer untouched (read-only passes such as the AWQ search)."""
    items = list(items)
    if not items:
        return
    if all(m.weight.is_cuda for _, m in items):
        for name, m in items:
            out = compute(name, m, m.weight.data)
            if out is not None:
                _assign(m, out)
        return
    _ops.require_cuda()
    _run_host_layers(items, compute)


def _run_host_layers(items: List[Tuple[str, nn.Linear]], compute: Compute) -> None:
    dev = torch.device("cuda", torch.cuda.current_device())
    cur = torch.cuda.current_stream(dev)
    h2d = torch.cuda.Stream(dev)
    d2h = torch.cuda.Stream(dev)
    n = len(items)
    staged = [None] * n
    ready = [None] * n

    def prefetch(i: int) -> None:
        w = items[i][1].weight.data
        if w.is_cuda:
            staged[i] = w.contiguous()
            return
        if i == 0:
            h2d.wait_stream(cur)
        with torch.cuda.stream(h2d):
            buf = torch.empty(w.shape, dtype=w.dtype, device=dev)
            buf.copy_(w, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(h2d)
        buf.record_stream(cur)
        staged[i], ready[i] = buf, ev

    pending = []   # device results kept alive until their D2H copy has been issued and finished
    prefetch(0)
    for i, (name, m) in enumerate(items):
        if i + 1 < n:
            prefetch(i + 1)
        if ready[i] is not None:
            cur.wait_event(ready[i])
        W = staged[i]
        staged[i] = None
        out = compute(name, m, W)
        if out is None:
            continue
        host = m.weight.data
        if host.is_cuda:
            _assign(m, out)
            continue
        done = torch.cuda.Event()
        done.record(cur)
        d2h.wait_event(done)
        out.record_stream(d2h)
        with torch.cuda.stream(d2h):
            if host.shape == out.shape and host.dtype == out.dtype and host.is_contiguous():
                host.copy_(out, non_blocking=True)      # in place: no host allocation per layer
            else:
                new = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                new.copy_(out, non_blocking=True)
                _assign(m, new)
        pending.append(out)
    d2h.synchronize()
    cur.wait_stream(d2h)
    pending.clear()
