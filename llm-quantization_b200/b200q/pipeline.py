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

# ---- device copies kept between passes over HOST-resident weights --------------------------------
# A search pass followed by a quantize pass (awq_search_scale_factor -> awq_quantize_model_weight)
# would stream every weight host -> device twice.  Inside `with keep_resident():` the device copy a
# READ-ONLY pass made of a host weight is kept (up to a byte budget) and the next pass over the
# same, unmodified host tensor uses it instead of copying again; a pass that writes a layer drops
# that layer's copy.  Opt-in and scoped: nothing is cached outside the block.
_resident = None     # {key: device tensor} while a keep_resident() block is active
_resident_budget = 0
_resident_bytes = 0


def _key(w: torch.Tensor):
    return (w.data_ptr(), tuple(w.shape), w.dtype, w._version)


class keep_resident:
    def __init__(self, max_bytes: Optional[int] = None):
        self.max_bytes = max_bytes

    def __enter__(self):
        global _resident, _resident_budget, _resident_bytes
        self._prev = (_resident, _resident_budget, _resident_bytes)
        budget = self.max_bytes
        if budget is None and torch.cuda.is_available():
            free, _total = torch.cuda.mem_get_info()
            # memory sitting unused in torch's allocator cache is as good as free
            free += torch.cuda.memory_reserved() - torch.cuda.memory_allocated()
            budget = int(free * 0.5)          # leave half of it to the kernels' scratch
        _resident, _resident_budget, _resident_bytes = {}, int(budget or 0), 0
        return self

    def __exit__(self, *exc):
        global _resident, _resident_budget, _resident_bytes
        _resident, _resident_budget, _resident_bytes = self._prev
        return False



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


def _remember(key, W: torch.Tensor) -> None:
    global _resident_bytes
    if _resident is None or key is None or key in _resident:
        return
    nbytes = W.numel() * W.element_size()
    if _resident_bytes + nbytes <= _resident_budget:
        _resident[key] = W
        _resident_bytes += nbytes


def _forget(key) -> None:
    global _resident_bytes
    if _resident is not None and key is not None:
        W = _resident.pop(key, None)
        if W is not None:
            _resident_bytes -= W.numel() * W.element_size()


def _run_host_layers(items: List[Tuple[str, nn.Linear]], compute: Compute) -> None:
    dev = torch.device("cuda", torch.cuda.current_device())
    cur = torch.cuda.current_stream(dev)
    h2d = torch.cuda.Stream(dev)
    d2h = torch.cuda.Stream(dev)
    n = len(items)
    staged = [None] * n
    ready = [None] * n
    keys = [None] * n

    def prefetch(i: int) -> None:
        w = items[i][1].weight.data
        if w.is_cuda:
            staged[i] = w.contiguous()
            return
        if _resident is not None:
            hit = _resident.get(_key(w))
            if hit is not None:
                staged[i], keys[i] = hit, _key(w)
                return
        keys[i] = _key(w)
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
            _remember(keys[i], W)           # read-only pass: the copy may serve the next pass
            continue
        _forget(keys[i])                    # the host tensor is about to change
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
            copied = torch.cuda.Event()
            copied.record(d2h)
        pending.append((copied, out))
        while pending and pending[0][0].query():      # results whose copy has finished can go
            pending.pop(0)
    d2h.synchronize()
    cur.wait_stream(d2h)
    pending.clear()
