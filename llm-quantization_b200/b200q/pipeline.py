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

from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops as _ops

Compute = Callable[[str, nn.Linear, torch.Tensor], torch.Tensor]


def _assign(module: nn.Linear, out: torch.Tensor) -> None:
    module.weight.data = out


def run_layers(items: Sequence[Tuple[str, nn.Linear]], compute: Compute) -> None:
    """For every (name, linear): out = compute(name, linear, W_on_gpu); linear.weight.data <- out
    on the weight's original device.  `compute` must launch on the current stream and may return
    None to leave the layer untouched."""
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
        h2d.wait_stream(cur) if i == 0 else None
        with torch.cuda.stream(h2d):
            buf = torch.empty(w.shape, dtype=w.dtype, device=dev)
            buf.copy_(w, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(h2d)
        buf.record_stream(cur)
        staged[i], ready[i] = buf, ev

    pending = []   # (event, device tensor) kept alive until their D2H finished
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


# --------------------------------------------------------------------------------------------------
# bench helper: the same walker call, model weights in pinned host memory
# --------------------------------------------------------------------------------------------------
def bench_host_roundtrip(method, model, originals, feats_by_K, act_by_K, w_bit, group, steps, world,
                         barrier):
    """Time `steps` walker passes over a host-resident copy of `model` (pinned memory): every step
    copies every weight host->device and every result device->host inside the timed region."""
    import time

    import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer
    from . import dist as bdist

    names = list(originals)
    total = sum(originals[n].numel() for n in names)
    dtype = originals[names[0]].dtype
    flat = torch.empty(total, dtype=dtype, pin_memory=True)
    host_model = type(model)()
    off = 0
    for n in names:
        w = originals[n]
        view = flat[off:off + w.numel()].view(w.shape)
        view.copy_(w)
        off += w.numel()
        lin = nn.Linear(w.shape[1], 1, bias=False)
        lin.weight = nn.Parameter(view, requires_grad=False)
        host_model.layers[n] = lin
    torch.cuda.synchronize()
    # calibration statistics arrive on the host too, as the reference's hooks produce them
    feats_host = {K: v.cpu() for K, v in feats_by_K.items()}
    act_host = {K: v.cpu() for K, v in act_by_K.items()}

    def fd(src):
        return {n: src[m.in_features] for n, m in host_model.named_modules() if isinstance(m, nn.Linear)}

    def call():
        if method == "awq":
            awq_quantizer.awq_quantize_model_weight(host_model, w_bit, group, fd(feats_host), 0.01, 2.0)
        elif method == "gptq":
            gptq_quantizer.gptq_quantize_model_weight(host_model, w_bit, group, fd(feats_host),
                                                      verbose=False)
        elif method == "pot":
            pot_apot_quantizer.pot_quantize_model_weight(host_model, w_bit, group)
        elif method == "apot":
            pot_apot_quantizer.apot_quantize_model_weight(host_model, w_bit, group, k=2)
        else:
            smooth_quant_quantizer.smoothquant_quantize_model_weight(host_model, 8, group, fd(act_host),
                                                                     alpha=0.5, verbose=False)

    def step():
        if world > 1:
            with bdist.row_sharded():
                call()
        else:
            call()

    step()          # warm-up (allocator, pinned staging)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    dev_ms = e0.elapsed_time(e1) / steps
    nbytes = total * flat.element_size()
    feat_bytes = sum(feats_host[m.in_features].numel() * 4 for m in host_model.layers.values()) \
        if method in ("awq", "gptq") else 0
    return {"ms_per_step": max(wall, dev_ms), "h2d_bytes": nbytes + feat_bytes, "d2h_bytes": nbytes,
            "how": "walker on a pinned-host model: per Linear H2D prefetch / kernels / D2H overlapped "
                   "on three streams; wall clock and CUDA events, max of the two"}
