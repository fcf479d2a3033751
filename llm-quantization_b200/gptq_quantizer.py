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
from typing import Dict, List

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
FACTOR_STREAMS = 16 # Hessian inverses in flight at a time on one GPU (each on its own CUDA stream)
LOCAL_GROUP = 16    # layers prepared together PER RANK (Hessians, then their inverses side by side)
TIMINGS = None      # set to a list to collect (phase, ms) CUDA-event pairs from the model walker
TRACE = None        # set to a list: (group start, begin, end, K) events of every concurrent inverse
HOST_LAPS = None    # set to a dict to add up host seconds per walker phase (launch-side cost)


class _Prepared:
    """What the column stage of one layer needs, plus the bookkeeping of an inverse that may still
    be running on a side stream: `done` (event on that stream), `info_host` (pinned copy of the
    factorisation's status flag, valid once `done` has completed)."""
    __slots__ = ("name", "H", "perm", "factor", "done", "info", "info_host", "K")

    def __init__(self, name, H, perm, factor, K, done=None, info=None, info_host=None):
        self.name, self.H, self.perm, self.factor, self.K = name, H, perm, factor, K
        self.done, self.info, self.info_host = done, info, info_host

    def check(self):
        """Wait for the factorisation (host side) and act on its status flag."""
        if self.done is not None:
            self.done.synchronize()
            self.done = None
        if self.info_host is not None:
            _report_pivot(int(self.info_host.item()), self.K, self.name)
            self.info_host = None
        elif self.info is not None:
            _report_pivot(int(self.info.item()), self.K, self.name)
        self.info = None


def _report_pivot(j: int, K: int, name: str) -> None:
    """A non-positive Cholesky pivot means H + 1e-6 I was not positive definite.  The reference
    hides that behind `except: pinv` (gptq_quantizer.py:161-165) and its output never reads H^-1;
    parity mode therefore warns and carries on, compensated mode (which multiplies the factor into
    the weights) raises."""
    if j == 0:
        return
    from b200q import tensor_ops as _tops
    if MODE == "compensated":
        _tops.raise_if_not_spd(j, K, f"gptq layer {name!r}")
    import warnings
    warnings.warn(f"gptq layer {name!r}: damped Hessian not positive definite (pivot {j}); H^-1 is "
                  f"not used by the reference-parity output, continuing")


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
    ready: Dict[str, _Prepared] = {}
    retiring: List[_Prepared] = []          # status flags not yet looked at (checked one group late)
    side_streams: List[torch.cuda.Stream] = []
    # one pinned buffer for all status flags (a pinned allocation per layer would call
    # cudaHostAlloc, which synchronises the device and serialises the concurrent inverses)
    flags_host = torch.zeros(max(1, len(calibrated)), dtype=torch.int32).pin_memory() \
        if torch.cuda.is_available() else None
    # per-slot matrices of the concurrent inverses (slot = position inside a group): allocated on
    # the main stream once per size and reused by every group, so that the side streams never touch
    # the allocator (a cudaMalloc / cudaFree there synchronises the device and serialises them)
    slots: Dict[int, dict] = {}

    def slot_buffers(k, K, device):
        b = slots.get(k)
        if b is None or b["out"].shape[0] != K:
            b = {"out": torch.empty((K, K), dtype=torch.float32, device=device),
                 "scratch": torch.empty((K, K), dtype=torch.float32, device=device),
                 "info": torch.zeros(1, dtype=torch.int32, device=device)}
            slots[k] = b
        return b

    in_flight: List[torch.cuda.Event] = []  # broadcasts of the previous group (parity mode)

    group_status: List[tuple] = []          # (event, pinned flags, [(name, K)]) of sharded groups
    colmax_ready: Dict[str, torch.Tensor] = {}   # all-reduced column |max| of a group's resident layers

    def check_group_status():
        for seen, host, layers in group_status:
            seen.synchronize()
            for (name, K), j in zip(layers, host.tolist()):
                _report_pivot(int(j), K, name)
        group_status.clear()

    def retire():
        for ev in in_flight:
            torch.cuda.current_stream().wait_event(ev)
        in_flight.clear()
        check_group_status()
        for p in retiring:
            p.check()
        retiring.clear()

    def factor_concurrently(jobs, device):
        """jobs: [(name, H, owner)].  A factorisation is a chain of ~500 small dependent kernels
        (64-column diagonal blocks, leaf GEMMs, operand splits: linalg.cu) that keeps only a few
        SMs busy, and the layers of a group are independent: the ones this rank owns run
        CONCURRENTLY, each on its own CUDA stream, with the GPU to themselves (the Hessian GEMMs
        occupy every SM's shared memory, so the two phases are not interleaved); the main stream
        waits for all of them.  Status flags travel to pinned host memory and are read one group
        later (no host sync on the critical path)."""
        main = torch.cuda.current_stream(device)
        mine = [j for j in jobs if j[1] is not None and (not _dist.is_sharded() or _dist.rank() == j[2])]
        n_streams = _factor_stream_count(max((j[1].shape[0] for j in mine), default=0), len(mine), device)
        while len(side_streams) < n_streams:
            side_streams.append(torch.cuda.Stream(device))
        if TRACE is not None:
            trace_start = torch.cuda.Event(enable_timing=True)
            trace_start.record(main)
        # (slot buffers first, on the main stream, before anything is queued on the side streams)
        bufs = {j[0]: slot_buffers(i, j[1].shape[0], device) for i, j in enumerate(mine)} \
            if n_streams > 1 else {}
        start = torch.cuda.Event()
        start.record(main)
        out, k = [], 0
        for name, H, owner in jobs:
            if H is None or n_streams <= 1 or (_dist.is_sharded() and _dist.rank() != owner):
                out.append(_factor_stage(name, H, actorder, owner))     # inline (or a placeholder)
                continue
            side = side_streams[k % n_streams]
            k += 1
            if k <= n_streams:
                side.wait_event(start)
            with torch.cuda.stream(side):
                if TRACE is not None:
                    b = torch.cuda.Event(enable_timing=True)
                    b.record(side)
                p = _factor_stage(name, H, actorder, owner, bufs[name])
                if TRACE is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record(side)
                    TRACE.append((trace_start, b, e, H.shape[0]))
                p.info_host = flags_host[position[name]:position[name] + 1]
                p.info_host.copy_(p.info, non_blocking=True)
                p.done = torch.cuda.Event()
                p.done.record(side)
            # memory handed across streams: H is read by the side stream (the factor and the flag
            # live in slot buffers owned by the main stream; the main stream waits for `done`)
            H.record_stream(side)
            if p.perm is not None:
                p.perm.record_stream(main)
            out.append(p)
        for p in out:
            if p.done is not None:
                main.wait_event(p.done)
        return out

    def prepare(group, device):
        # A GROUP of layers is prepared at once.  Under row sharding every rank adds its calibration
        # samples to each layer's Hessian (all-reduced), then the ranks factor DIFFERENT layers of
        # the group at the same time (dealt longest-first by K^3, so a rank that draws an
        # 11008-wide layer gets fewer 4096-wide ones) and the factors are broadcast.  Three phases,
        # so that no collective sits between two ranks' factorisations (an all-reduce there would
        # make everybody wait for whoever is busy inverting).
        world = _dist.world_size()
        owner = _deal_layers([(n, m.weight.shape[1]) for n, m in group], world)
        t0 = _mark()
        # (under sharding each layer's exchange -- packed all-reduce, unpack, damping -- runs on the
        # communication stream while the next layer's partial Hessian is computed)
        main = torch.cuda.current_stream(device)
        hessians, pending = [], []
        for n, m in group:
            H, ev = _hessian_stage(input_feat[n], m.weight.shape[1], device, perp_damp, nsamples,
                                   defer_exchange=True)
            hessians.append(H)
            if ev is not None:
                pending.append((ev, H))
        if pending:
            with _dist.timed_wait(sum(2 * H.numel() for _e, H in pending)):
                for ev, _H in pending:
                    main.wait_event(ev)
        t1 = _mark()
        import time as _time
        h0 = _time.perf_counter()
        retire()                      # the previous group's flags: long since on the host
        h1 = _time.perf_counter()
        prepared = factor_concurrently([(n, H, owner[n]) for (n, _m), H in zip(group, hessians)], device)
        if HOST_LAPS is not None:
            HOST_LAPS["retire"] = HOST_LAPS.get("retire", 0.0) + (h1 - h0)
            HOST_LAPS["factor launch"] = HOST_LAPS.get("factor launch", 0.0) + (_time.perf_counter() - h1)
        t2 = _mark()
        if world > 1:
            # The factors go out on the communication stream.  Parity mode never reads H^-1
            # (gptq_quantizer.py:189-194): the column stages and the next group's Hessians run
            # while the broadcasts are in flight and the group is joined when it retires;
            # compensated mode waits before its column stages.
            from b200q import tensor_ops as _tops
            owned = [p for p in prepared if p.factor is not None]
            if owned:
                # one status exchange for the whole group, BEFORE the broadcasts are queued (NCCL runs
                # a communicator's collectives in order): every rank gets every owner's flag, so all
                # ranks warn / raise together -- one group later, from pinned memory, without a host
                # sync here
                status = _dist.allreduce_max(torch.cat([p.info for p in owned]))
                lo = position[owned[0].name]
                host = flags_host[lo:lo + len(owned)] if lo + len(owned) <= flags_host.numel() else \
                    torch.empty(len(owned), dtype=torch.int32).pin_memory()
                host.copy_(status, non_blocking=True)
                seen = torch.cuda.Event()
                seen.record(main)
                group_status.append((seen, host, [(p.name, p.K) for p in owned]))
            comm = _tops.comm_stream(device)
            comm.wait_stream(main)
            with torch.cuda.stream(comm), _dist.on_comm_stream():
                for p in owned:
                    _dist.broadcast(p.factor, owner[p.name])
                    p.factor.record_stream(comm)
                    p.done = p.info_host = p.info = None
                sent = torch.cuda.Event()
                sent.record(comm)
            if MODE == "compensated":
                with _dist.timed_wait(sum(4 * p.factor.numel() for p in owned)):
                    main.wait_event(sent)
                check_group_status()          # the column stages multiply by the factors
            else:
                in_flight.append(sent)
        if world > 1 and MODE == "parity":
            # The reference's column scale spans ALL rows (:182): the shards' column maxima are
            # combined with an all-reduce MAX.  For the device-resident layers of the group that is
            # ONE collective over the concatenated [sum K] vector instead of one per layer (at 8
            # GPUs the per-layer launches and waits were a fifth of the step).
            resident = [(n, m) for n, m in group if m.weight.is_cuda]
            if resident:
                sizes = [m.weight.shape[1] for _n, m in resident]
                padded = [(k + 3) // 4 * 4 for k in sizes]           # 16-byte aligned slices
                flat = torch.zeros(sum(padded), dtype=torch.float32, device=device)
                views = [c[:k] for c, k in zip(flat.split(padded), sizes)]
                for (_n, m), v in zip(resident, views):
                    _ops.col_absmax(m.weight.data, out=v)
                _dist.allreduce_max(flat)
                for (n, _m), v in zip(resident, views):
                    colmax_ready[n] = v
        for p in prepared:
            ready[p.name] = p
        _lap("hessians", t0, t1)
        _lap("factors", t1, t2)
        _lap("broadcast", t2, _mark())

    def compute(name, _module, W):
        if name not in input_feat:
            return _symmetric_groups(W, w_bit, q_group_size)
        if name not in ready:
            i = position[name]
            world = _dist.world_size()
            prepare(calibrated[i:i + max(1, LOCAL_GROUP * max(1, world))],
                    W.device)
        p = ready.pop(name)
        if p.done is not None or p.info_host is not None:
            if MODE == "compensated":
                p.check()             # the column stage multiplies by the factor: it must be sound
            else:
                retiring.append(p)    # parity output does not read H^-1: look at the flag later
        t2 = _mark()
        out = _column_stage(W, w_bit, q_group_size, blocksize, p.H, p.perm, p.factor,
                            colmax=colmax_ready.pop(name, None))
        _lap("columns", t2, _mark())
        return out

    try:
        _pipeline.run_layers(items, compute)
    finally:
        retiring.extend(ready.values())
        ready.clear()
        retire()


def _factor_stream_count(K: int, n_jobs: int, device) -> int:
    """Factorisations in flight at once on this GPU: FACTOR_STREAMS, fewer when their workspaces
    (3 K^2 floats + fp16 operand planes each, cached per stream) and matrices would not fit in 40 %
    of the memory that is free or sitting unused in torch's allocator cache."""
    if n_jobs <= 1 or K <= 0:
        return min(1, n_jobs)
    from b200q import _lib as _l
    per = _l.load().b200q_spd_inverse_workspace(K) + 3 * 4 * K * K
    free, _total = torch.cuda.mem_get_info(device)
    free += torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return int(max(1, min(FACTOR_STREAMS, n_jobs, int(free * 0.4) // max(per, 1))))


def _deal_layers(layers, world: int) -> Dict[str, int]:
    """{name: rank} for a group of (name, in_features) pairs: longest factorisation first (cost
    ~ in_features^3), each to the rank with the least work so far.  Deterministic, so every rank
    derives the same assignment without talking."""
    load = [0.0] * max(1, world)
    owner: Dict[str, int] = {}
    for name, k in sorted(layers, key=lambda nk: (-nk[1], nk[0])):
        j = min(range(len(load)), key=lambda r: (load[r], r))
        owner[name] = j
        load[j] += float(k) ** 3
    return owner


def _mark():
    if TIMINGS is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _lap(phase, e0, e1):
    if TIMINGS is not None:
        TIMINGS.append((phase, e0, e1))


# ==================================================================================================
# per-layer stages
# ==================================================================================================
@torch.no_grad()
def _simple_quantize_layer(layer: nn.Linear, n_bit: int, q_group_size: int) -> None:
    """Symmetric |max| group quantization, codes in [-2^b, 2^b-1] (reference: :79-108)."""
    w = layer.weight.data
    src = w.device
    out = _symmetric_groups(_ops.to_device(w), n_bit, q_group_size)
    layer.weight.data = out if out.device == src else out.to(src)


def _symmetric_groups(W: torch.Tensor, n_bit: int, q_group_size: int) -> torch.Tensor:
    """The reference groups with `w.reshape(-1, q_group_size)` (:88-91): consecutive elements of
    the FLATTENED weight, so a group size that does not divide in_features is still legal as long
    as it divides the element count -- groups then run across row boundaries."""
    if q_group_size > 0 and W.shape[-1] % q_group_size != 0:
        flat = W.reshape(-1, q_group_size)          # raises like the reference if numel % G != 0
        return _ops.group_fakequant(flat, n_bit, -1, symmetric=True).reshape(W.shape)
    return _ops.group_fakequant(W, n_bit, q_group_size, symmetric=True)


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


def _hessian_stage(input_feat, K: int, device, perp_damp: float, nsamples: int,
                   defer_exchange: bool = False):
    """The damped Hessian of one layer (all-reduced under row sharding), or None when parity mode
    is told to skip the work the reference's output does not depend on.  defer_exchange: returns
    (H, event) with the cross-rank exchange queued on the communication stream (see
    tensor_ops.gptq_hessian); event is None when there is nothing to wait for."""
    if MODE not in ("parity", "compensated"):
        raise ValueError(f"gptq_quantizer.MODE must be 'parity' or 'compensated', got {MODE!r}")
    skip = MODE == "parity" and not BUILD_HESSIAN
    if not skip and K % 8 != 0:
        # the tensor-core Hessian kernel reads the activations by TMA (16-byte row pitch)
        if MODE == "compensated":
            raise NotImplementedError(f"compensated GPTQ needs in_features % 8 == 0, got {K}")
        import warnings
        warnings.warn(f"gptq: in_features = {K} is not a multiple of 8; H and H^-1 (which the "
                      f"reference-parity output does not read) are not built for this layer")
        skip = True
    if skip:
        return (None, None) if defer_exchange else None
    from b200q import tensor_ops as _tops
    return _tops.gptq_hessian(input_feat, K, device, perp_damp, nsamples, defer_exchange=defer_exchange)


def _factor_stage(name: str, H, actorder: bool, owner: int = 0, buffers=None) -> _Prepared:
    """The act-order permutation (compensated mode only) and what the column stage needs from the
    inverse -- H^-1 in parity mode (built like the reference builds it; its output does not depend
    on it), U = chol(H^-1) in compensated mode -- launched on the CURRENT stream, status flag left
    on the device (the caller checks it: _Prepared.check).  Under row sharding only rank `owner`
    computes the factor and the caller broadcasts it; no collective happens here."""
    from b200q import tensor_ops as _tops
    if H is None:
        return _Prepared(name, None, None, None, 0)
    K = H.shape[0]
    if MODE == "compensated":
        perm = torch.argsort(torch.diag(H), descending=True) if actorder else None
        U, info = _tops.compensation_factor(H, perm, owner=owner, broadcast=False, check=False,
                                            return_info=True, buffers=buffers)
        return _Prepared(name, H, perm, U, K, info=info)
    Hinv, info = _tops.spd_inverse(H, ridge=1e-6, owner=owner, broadcast=False, check=False,
                                   return_info=True, buffers=buffers)
    return _Prepared(name, H, None, Hinv, K, info=info)


def _gptq_device(W: torch.Tensor, n_bit: int, q_group_size: int, input_feat, perp_damp: float,
                 blocksize: int, nsamples: int, actorder: bool) -> torch.Tensor:
    """The per-layer stages on a CUDA-resident [N,K] weight; returns the quantized weight."""
    H = _hessian_stage(input_feat, W.shape[1], W.device, perp_damp, nsamples)
    p = _factor_stage("layer", H, actorder)
    if p.factor is not None:
        _dist.broadcast(p.factor, 0)
        _dist.broadcast(p.info, 0)
        p.check()
    return _column_stage(W, n_bit, q_group_size, blocksize, p.H, p.perm, p.factor)


def _column_stage(W: torch.Tensor, n_bit: int, q_group_size: int, blocksize: int, H, perm,
                  factor, colmax=None) -> torch.Tensor:
    if MODE == "compensated":
        from b200q import tensor_ops as _tops
        out = _tops.gptq_compensated(W, None, n_bit, q_group_size, blocksize, perm, U=factor)
    elif MODE == "parity":
        # The reference's loop rounds column j with s_j = clamp(max_i |W[i,j]| / (2^b-1), 1e-5) and
        # writes q*s back; permuting and un-permuting independent columns is the identity.
        if _dist.is_sharded():
            # the column scale spans ALL rows (:182): combine the shards' column maxima first
            if colmax is None:
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
    return _calibrate(model, calib_samples, nsamples, verbose, streaming=False)


@torch.no_grad()
def gptq_calibrate_hessian_streaming(
    model: nn.Module,
    calib_samples: List[torch.Tensor],
    nsamples: int = 128,
    verbose: bool = True,
):
    """Not in the reference: like gptq_calibrate_hessian, but every batch is
    folded into a running Hessian inside the hook (b200q.streaming.ActivationStream), so no
    activations are kept.  The returned dict goes to gptq_quantize_model_weight like the lists."""
    return _calibrate(model, calib_samples, nsamples, verbose, streaming=True)


def _calibrate(model, calib_samples, nsamples, verbose, streaming):
    import tqdm
    from b200q.streaming import ActivationStream

    captured: Dict[str, List[torch.Tensor]] = {}

    def make_hook(name: str):
        def hook(_m, inputs, _out):
            x = inputs[0] if isinstance(inputs, tuple) else inputs
            if x.dim() > 2:
                x = x.reshape(-1, x.shape[-1])
            if streaming:
                if name not in captured:
                    captured[name] = ActivationStream(x.shape[-1], normalize=True,
                                                      max_batches=nsamples, keep_stats=False)
                captured[name].add(x)
            else:
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
