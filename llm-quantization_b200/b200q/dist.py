"""Row-sharded execution over several GPUs (one process per GPU, torch.distributed / NCCL).

The quantizers shard a Linear's OUTPUT rows: each rank holds W[r0:r1, :].  Rows are independent in
every method except for three small exchanges, which are the only collectives on the path:

  * GPTQ (parity) and SmoothQuant use a per-input-column |max| over ALL rows
    (gptq_quantizer.py:182, smooth_quant_quantizer.py:156)           -> all-reduce MAX of fp32 [K]
  * the GPTQ Hessian is accumulated over calibration samples, which are dealt round-robin to the
    ranks                                                            -> all-reduce SUM of fp32 [K,K]
    and its inverse is computed once and shared                      -> broadcast of fp32 [K,K]
  * the AWQ scale search adds per-candidate losses over row shards   -> all-reduce SUM of [n_cand]
  * APOT picks its grid from the element count of the WHOLE tensor (pot_apot_quantizer.py:258)
                                                                     -> all-reduce SUM of one int

Outside a `row_sharded()` block every helper is the identity, so the single-GPU path pays nothing.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.distributed as td

_active_group = None
_active = False

# bench.py sets this to a list to learn how long the compute stream sat waiting on collectives:
# every collective below is then bracketed by two CUDA events on the current stream and recorded as
# (start, end, payload bytes).  None (the default) costs nothing.
WAIT_EVENTS = None
_on_comm_stream = False   # collectives issued on the communication stream stall THAT stream, not compute


class _Timed:
    __slots__ = ("nbytes", "e0")

    def __init__(self, t):
        self.nbytes = int(t.numel() * t.element_size()) if t is not None else 0
        self.e0 = None

    def __enter__(self):
        if WAIT_EVENTS is not None and not _on_comm_stream and torch.cuda.is_available():
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and WAIT_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            WAIT_EVENTS.append((self.e0, e1, self.nbytes))
        return False


class on_comm_stream:
    """Marks collectives issued inside the block as running on a communication stream: their
    duration is not compute-stream exposure (the consumer's timed_wait is)."""

    def __enter__(self):
        global _on_comm_stream
        self._prev, _on_comm_stream = _on_comm_stream, True
        return self

    def __exit__(self, *exc):
        global _on_comm_stream
        _on_comm_stream = self._prev
        return False


class _TimedWork:
    """Async collective handle whose wait() is bracketed like the synchronous collectives."""
    __slots__ = ("work", "t")

    def __init__(self, work, t):
        self.work, self.t = work, t

    def wait(self):
        with _Timed(self.t):
            self.work.wait()


@contextlib.contextmanager
def row_sharded(group: Optional["td.ProcessGroup"] = None):
    """Mark the enclosed quantizer calls as operating on this rank's row shard of every weight."""
    global _active, _active_group
    if not (td.is_available() and td.is_initialized()):
        raise RuntimeError("row_sharded() needs an initialised torch.distributed process group")
    prev = (_active, _active_group)
    _active, _active_group = True, group
    try:
        yield
    finally:
        _active, _active_group = prev


def is_sharded() -> bool:
    return _active and td.get_world_size(_active_group) > 1


def world_size() -> int:
    return td.get_world_size(_active_group) if _active else 1


def rank() -> int:
    return td.get_rank(_active_group) if _active else 0


def allreduce_max(t: torch.Tensor) -> torch.Tensor:
    if is_sharded():
        with _Timed(t):
            td.all_reduce(t, op=td.ReduceOp.MAX, group=_active_group)
    return t


def allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    if is_sharded():
        with _Timed(t):
            td.all_reduce(t, op=td.ReduceOp.SUM, group=_active_group)
    return t


def allreduce_sum_async(t: torch.Tensor):
    """Start the sum over ranks of `t` (in place) and return the work handle, or None when not
    sharded.  handle.wait() orders the caller's current stream after the collective."""
    if is_sharded():
        return _TimedWork(td.all_reduce(t, op=td.ReduceOp.SUM, group=_active_group, async_op=True), t)
    return None


def backend_is_nccl() -> bool:
    return is_sharded() and td.get_backend(_active_group) == "nccl"


def reduce_scatter_sum(out: torch.Tensor, inp: torch.Tensor) -> torch.Tensor:
    """out (this rank's 1/world slice) = sum over ranks of the matching slice of inp."""
    with _Timed(inp):
        td.reduce_scatter_tensor(out, inp, op=td.ReduceOp.SUM, group=_active_group)
    return out


def all_gather_into(out: torch.Tensor, inp: torch.Tensor) -> torch.Tensor:
    with _Timed(out):
        td.all_gather_into_tensor(out, inp, group=_active_group)
    return out


class timed_wait:
    """Brackets a wait of the CURRENT stream on communication work (an event of the communication
    stream) the way the collectives above are bracketed, so that bench.py's nccl_exposed counts it."""

    def __init__(self, nbytes: int = 0):
        self.nbytes, self.e0 = nbytes, None

    def __enter__(self):
        if WAIT_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and WAIT_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            WAIT_EVENTS.append((self.e0, e1, self.nbytes))
        return False


def broadcast(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if is_sharded():
        with _Timed(t):
            td.broadcast(t, src=td.get_global_rank(_active_group, src) if _active_group else src,
                         group=_active_group)
    return t


def gather_rows(compute, n_rows: int, row_shape, dtype, device) -> torch.Tensor:
    """[n_rows, *row_shape] tensor whose rows are produced by `compute(lo, hi)` (-> rows lo..hi-1):
    under sharding each rank computes an equal share and the shares are all-gathered; otherwise (or
    when n_rows does not divide evenly) the caller's rank computes everything."""
    if not is_sharded() or n_rows % world_size() != 0:
        return compute(0, n_rows)
    w, r = world_size(), rank()
    per = n_rows // w
    local = compute(r * per, (r + 1) * per).contiguous()
    out = torch.empty((n_rows, *row_shape), dtype=dtype, device=device)
    with _Timed(out):
        td.all_gather_into_tensor(out, local, group=_active_group)
    return out


def global_numel(local_numel: int, device) -> int:
    """Element count of the whole (unsharded) tensor."""
    if not is_sharded():
        return local_numel
    n = torch.tensor([local_numel], dtype=torch.int64, device=device)
    td.all_reduce(n, op=td.ReduceOp.SUM, group=_active_group)
    return int(n.item())


def shard_rows(n_rows: int, world: Optional[int] = None, r: Optional[int] = None):
    """Contiguous, near-equal row range [r0, r1) of rank r."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    base, extra = divmod(n_rows, world)
    r0 = r * base + min(r, extra)
    return r0, r0 + base + (1 if r < extra else 0)
