"""Tensor-level wrappers for the tensor-core stages (Hessian, SPD inverse, AWQ search loss,
compensated GPTQ).  Same rules as b200q.ops: torch owns memory and streams, libb200quant does the
arithmetic, nothing falls back to torch or the CPU."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, dist as _dist
from .ops import DTYPE_CODE, _on, _stream, dtype_code, require_cuda

_scratch = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch.pop(key, None)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def release_workspace() -> None:
    """Drop the cached Hessian scratch (it holds a scaled fp16 copy of the activations)."""
    _scratch.clear()


def hessian_accum(X: torch.Tensor, rows_per_sample: int, H: Optional[torch.Tensor] = None,
                  return_norms: bool = False, normalize: bool = True):
    """H (+)= sum_i X_i^T X_i / (||X_i||_F + 1e-5)^2 over equal-length samples of the CUDA [T,K]
    matrix X (fp32/fp16/bf16).  Returns H (fp32 [K,K]) or (H, norms)."""
    assert X.is_cuda and X.dim() == 2
    X = X.contiguous()
    T, K = X.shape
    assert T % rows_per_sample == 0
    n = T // rows_per_sample
    accumulate = H is not None
    if H is None:
        H = torch.empty((K, K), dtype=torch.float32, device=X.device)
    norms = torch.empty(n, dtype=torch.float32, device=X.device) if return_norms else None
    lib = _lib.load()
    with _on(X.device):
        work = _workspace(X.device, lib.b200q_hessian_workspace(T, K, n))
        rc = lib.b200q_hessian_accum(X.data_ptr(), n, rows_per_sample, K, dtype_code(X),
                                     int(normalize), H.data_ptr(), int(accumulate),
                                     None if norms is None else norms.data_ptr(), work.data_ptr(),
                                     _stream())
    _lib.check(rc, "hessian_accum")
    return (H, norms) if return_norms else H


def hessian_finalize(H: torch.Tensor, scale: float, damp: float) -> torch.Tensor:
    with _on(H.device):
        rc = _lib.load().b200q_hessian_finalize(H.data_ptr(), H.shape[0], float(scale), float(damp),
                                                _stream())
    _lib.check(rc, "hessian_finalize")
    return H


class FactorizationError(_lib.B200QuantError):
    """The matrix handed to spd_inverse was not positive definite (a Cholesky pivot <= 0)."""


def spd_inverse(H: torch.Tensor, ridge: float = 0.0, want_inverse: bool = True,
                want_upper: bool = False, owner: int = 0, broadcast: bool = True,
                check: bool = True, return_info: bool = False, buffers: Optional[dict] = None):
    """inv(H + ridge I) (and/or the upper Cholesky factor U of it, U^T U = inverse) for a CUDA
    fp32 SPD matrix.  Returns Hinv, U or (Hinv, U).  Under row sharding rank `owner` computes and
    the result is broadcast (the inverse is shared by all row shards); with broadcast=False the
    other ranks get an uninitialised buffer and the caller broadcasts later, which lets different
    ranks invert different layers' matrices at the same time.

    The kernels report a non-positive pivot through a device flag (0 = ok, j = pivot of column j).
    check=True reads it (one host sync) and raises FactorizationError -- on every rank under
    sharding -- instead of handing back garbage; check=False leaves that to the caller, who gets
    the int32 flag tensor with return_info=True (the model walker checks a whole group of layers
    with one sync).

    buffers: optional {"out": [K,K] fp32, "scratch": [K,K] fp32, "info": int32[1]} preallocated by
    the caller; nothing is then allocated here (the walker runs many inverses on side streams,
    where allocator traffic would synchronise them)."""
    assert H.is_cuda and H.dtype == torch.float32 and H.dim() == 2 and H.shape[0] == H.shape[1]
    K = H.shape[0]
    lib = _lib.load()
    buffers = buffers or {}
    out = buffers.get("out")
    Hinv = (out if out is not None else torch.empty_like(H)) if want_inverse else None
    U = (out if (out is not None and not want_inverse) else torch.empty_like(H)) if want_upper else None
    info = buffers.get("info")
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=H.device)
    else:
        info.zero_()
    if not _dist.is_sharded() or _dist.rank() == owner:
        A = H.contiguous()
        if ridge != 0.0:
            scratch = buffers.get("scratch")
            if scratch is not None:
                scratch.copy_(A)
                A = hessian_finalize(scratch, 1.0, ridge)
            else:
                A = hessian_finalize(A.clone(), 1.0, ridge)
        with _on(H.device):
            work = _workspace(H.device, lib.b200q_spd_inverse_workspace(K))
            rc = lib.b200q_spd_inverse(A.data_ptr(), None if Hinv is None else Hinv.data_ptr(),
                                       None if U is None else U.data_ptr(), K, work.data_ptr(),
                                       info.data_ptr(), _stream())
        _lib.check(rc, "spd_inverse")
    if broadcast:
        for t in (Hinv, U):
            if t is not None:
                _dist.broadcast(t, owner)
        if check or return_info:
            _dist.broadcast(info, owner)
    if check:
        raise_if_not_spd(info, K)
    out = (Hinv, U) if (Hinv is not None and U is not None) else (Hinv if Hinv is not None else U)
    return (out, info) if return_info else out


def raise_if_not_spd(info, K: int, what: str = "spd_inverse") -> None:
    """info: the device flag of spd_inverse (or its int value)."""
    j = int(info.item()) if isinstance(info, torch.Tensor) else int(info)
    if j != 0:
        raise FactorizationError(
            f"{what}: the {K} x {K} matrix is not positive definite (pivot {j} <= 0); the reference "
            f"would fall back to torch.linalg.pinv here (gptq_quantizer.py:161-165) -- raise the "
            f"damping (perp_damp) instead")


def compensation_factor(H: torch.Tensor, perm: Optional[torch.Tensor] = None, owner: int = 0,
                        broadcast: bool = True, check: bool = True, return_info: bool = False,
                        buffers: Optional[dict] = None):
    """U = chol(inv(H_perm + 1e-6 I), upper): what the compensated column loop multiplies by."""
    if perm is not None:
        H = H[perm][:, perm]
    return spd_inverse(H.contiguous(), ridge=1e-6, want_inverse=False, want_upper=True, owner=owner,
                       broadcast=broadcast, check=check, return_info=return_info, buffers=buffers)


def gptq_compensated(W: torch.Tensor, H: Optional[torch.Tensor], n_bit: int, group: int,
                     blocksize: int = 128, perm: Optional[torch.Tensor] = None,
                     U: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Error-compensated GPTQ of a CUDA [N,K] weight given its damped Hessian H (fp32 [K,K]) or the
    factor U = compensation_factor(H, perm) computed earlier.  `perm` (act-order) permutes the
    columns before and restores them after, as GPTQ does."""
    assert W.is_cuda and W.dim() == 2
    N, K = W.shape
    Wf = W.float()
    if perm is not None:
        Wf = Wf[:, perm]
    Wf = Wf.contiguous().clone() if Wf.data_ptr() == W.data_ptr() else Wf.contiguous()
    if U is None:
        assert H is not None and H.shape == (K, K)
        U = compensation_factor(H, perm)
    Q = torch.empty_like(Wf)
    lib = _lib.load()
    with _on(W.device):
        work = torch.empty(lib.b200q_gptq_compensated_workspace(N, K), dtype=torch.uint8,
                           device=W.device)
        rc = lib.b200q_gptq_compensated(Wf.data_ptr(), Q.data_ptr(), U.data_ptr(), N, K, group, n_bit,
                                        blocksize, work.data_ptr(), _stream())
    _lib.check(rc, "gptq_compensated")
    if perm is not None:
        Q = Q[:, torch.argsort(perm)]
    return Q.to(W.dtype)


# ---- symmetric matrices between GPUs: the packed lower triangle ---------------------------------
# False (or B200Q_PACKED_EXCHANGE=0): plain fp32 [K,K] all-reduce (A/B timing, non-NCCL backends)
PACKED_EXCHANGE = __import__("os").environ.get("B200Q_PACKED_EXCHANGE", "1") != "0"
_comm_streams = {}


def comm_stream(device) -> torch.cuda.Stream:
    """The per-device stream collectives are queued on when they should not block compute."""
    return _comm_stream(torch.device(device))


def _comm_stream(device) -> torch.cuda.Stream:
    s = _comm_streams.get(device.index)
    if s is None:
        # (high priority: its small pack / fold / unpack kernels must not queue behind a GEMM's grid)
        s = _comm_streams[device.index] = torch.cuda.Stream(device, priority=-1)
    return s


def sym_pack_lower(H: torch.Tensor, pad_to: int = 1) -> torch.Tensor:
    """fp32 [K,K] symmetric -> its lower triangle as a vector (row n = columns 0..n at offset
    n(n+1)/2), length rounded up to a multiple of pad_to (padding uninitialised)."""
    K = H.shape[0]
    L = K * (K + 1) // 2
    P = torch.empty((L + pad_to - 1) // pad_to * pad_to, dtype=torch.float32, device=H.device)
    with _on(H.device):
        rc = _lib.load().b200q_sym_pack_lower(H.data_ptr(), K, P.data_ptr(), _stream())
    _lib.check(rc, "sym_pack_lower")
    return P


def sym_unpack_lower(P: torch.Tensor, K: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """the packed triangle back into a full symmetric fp32 [K,K]."""
    H = out if out is not None else torch.empty((K, K), dtype=torch.float32, device=P.device)
    with _on(P.device):
        rc = _lib.load().b200q_sym_unpack_lower(P.data_ptr(), K, H.data_ptr(), _stream())
    _lib.check(rc, "sym_unpack_lower")
    return H


def allreduce_symmetric(H: torch.Tensor) -> torch.Tensor:
    """In-place sum over the ranks of a SYMMETRIC fp32 [K,K] matrix; only the lower triangle
    travels (half the bytes of all-reducing the square)."""
    if not _dist.is_sharded():
        return H
    if not (PACKED_EXCHANGE and H.is_cuda):
        return _dist.allreduce_sum(H)
    P = sym_pack_lower(H)
    _dist.allreduce_sum(P)
    return sym_unpack_lower(P, H.shape[0], out=H)


class FoldedGram:
    """The AWQ search operand already folded onto its lower triangle: bf16 [K,K],
    Hb[n][k] = H[n][k] + H[k][n] for k < n, H[n][n] on the diagonal, 0 above."""
    __slots__ = ("Hb",)

    def __init__(self, Hb):
        self.Hb = Hb


class PendingGram:
    """A Gram matrix whose cross-rank sum may still be in flight (see gram_matrix_begin)."""
    __slots__ = ("H", "work", "rows_total", "folded", "event", "nbytes")

    def __init__(self, H, work, rows_total, folded=None, event=None, nbytes=0):
        self.H, self.work, self.rows_total = H, work, rows_total
        self.folded, self.event, self.nbytes = folded, event, nbytes


def _exchange_folded(H: torch.Tensor) -> PendingGram:
    """Partial X^T X of this rank -> the search operand summed over all ranks, with the exchange on
    the communication stream:  pack lower triangle (fp32) -> reduce-scatter -> every rank folds ITS
    slice to bf16 -> all-gather (bf16) -> unpack into the [K,K] operand.  Per rank that moves
    (w-1)/w * 3 K^2 bytes against 8 K^2 for the fp32 all-reduce of the square, and the fold pass is
    shared out over the ranks."""
    K = H.shape[0]
    dev = H.device
    w, r = _dist.world_size(), _dist.rank()
    lib = _lib.load()
    main = torch.cuda.current_stream(dev)
    comm = _comm_stream(dev)
    P = sym_pack_lower(H, pad_to=8 * w)
    Lp = P.numel()
    per = Lp // w
    comm.wait_stream(main)
    with torch.cuda.stream(comm), _dist.on_comm_stream():
        mine = torch.empty(per, dtype=torch.float32, device=dev)
        _dist.reduce_scatter_sum(mine, P)
        mine16 = torch.empty(per, dtype=torch.bfloat16, device=dev)
        rc = lib.b200q_sym_fold_packed_bf16(mine.data_ptr(), K, r * per, (r + 1) * per,
                                            mine16.data_ptr(), _stream())
        _lib.check(rc, "sym_fold_packed_bf16")
        Pb = torch.empty(Lp, dtype=torch.bfloat16, device=dev)
        _dist.all_gather_into(Pb, mine16)
        Hb = torch.empty((K, K), dtype=torch.bfloat16, device=dev)
        rc = lib.b200q_sym_unpack_folded_bf16(Pb.data_ptr(), K, Hb.data_ptr(), _stream())
        _lib.check(rc, "sym_unpack_folded_bf16")
        ev = torch.cuda.Event()
        ev.record(comm)
    P.record_stream(comm)
    Hb.record_stream(main)
    return PendingGram(None, None, 0, folded=FoldedGram(Hb), event=ev, nbytes=Lp * 4 + Lp * 2)


def gram_matrix_begin(input_feat: Sequence, in_features: int, device,
                      want_folded: bool = False) -> PendingGram:
    """Launch X^T X over this rank's share of the calibration rows and, under row sharding, START
    the all-reduce of the partial sums without waiting for it: the caller can queue the next
    layer's kernels behind this one and pick the result up later with gram_matrix_end.
    want_folded: the caller is the AWQ search, which only needs the matrix folded onto its lower
    triangle in bf16 -- under sharding the exchange then runs packed (see _exchange_folded) and
    gram_matrix_end returns a FoldedGram."""
    device = torch.device(device)
    K = in_features
    from .streaming import ActivationStream
    if isinstance(input_feat, ActivationStream):
        assert not input_feat.normalize and input_feat.in_features == K
        H = input_feat.matrix_sum(device)            # (already all-reduced: nothing left in flight)
        return PendingGram(H, None, max(1, input_feat.total_rows(device)))
    if isinstance(input_feat, torch.Tensor):
        X = input_feat.reshape(-1, K)
    else:
        X = torch.cat([f.reshape(-1, K) for f in input_feat])
    X = X.to(device)
    if X.dtype not in DTYPE_CODE:
        X = X.float()
    rows_total = X.shape[0]
    # samples only matter for the fp16 pre-scaling of fp32 input here; use runs of up to 2048 rows
    rows = 1
    for cand in (2048, 1024, 512, 256, 128, 64, 32, 16, 8, 4, 2):
        if rows_total % cand == 0:
            rows = cand
            break
    work = None
    if _dist.is_sharded():
        n = rows_total // rows
        lo, hi = _dist.shard_rows(n, _dist.world_size(), _dist.rank())
        H = hessian_accum(X[lo * rows:hi * rows], rows, normalize=False) if hi > lo else \
            torch.zeros((K, K), dtype=torch.float32, device=device)
        if want_folded and PACKED_EXCHANGE and _dist.backend_is_nccl():
            p = _exchange_folded(H)
            p.rows_total = rows_total
            return p
        work = _dist.allreduce_sum_async(H)
    else:
        H = hessian_accum(X, rows, normalize=False)
    return PendingGram(H, work, rows_total)


def gram_matrix_end(p: PendingGram, normalise: bool = True) -> torch.Tensor:
    """The finished matrix: X^T X / rows, or the plain sum X^T X with normalise=False (a caller
    whose result is linear in H can apply 1 / p.rows_total to its own, much smaller, output)."""
    if p.folded is not None:
        assert not normalise, "a folded Gram matrix comes as the plain sum"
        with _dist.timed_wait(p.nbytes):
            torch.cuda.current_stream(p.folded.Hb.device).wait_event(p.event)
        return p.folded
    if p.work is not None:
        p.work.wait()                       # orders the current stream after the all-reduce
        p.work = None
    return hessian_finalize(p.H, 1.0 / p.rows_total, 0.0) if normalise else p.H


def gram_matrix(input_feat: Sequence, in_features: int, device) -> torch.Tensor:
    """X^T X / (number of rows) over all calibration features of a layer, fp32 [K,K]: the matrix
    in which the AWQ search measures output reconstruction error.  Feature lists follow the same
    conventions as gptq_hessian (1-D entries are single rows)."""
    return gram_matrix_end(gram_matrix_begin(input_feat, in_features, device))


def awq_search_losses(W: torch.Tensor, H: torch.Tensor, salient_mask: torch.Tensor, n_bit: int,
                      group: int, candidates: Sequence[float]) -> torch.Tensor:
    """fp32 [n_cand]: sum_rows dW_c H dW_c^T for every candidate scale factor, for the CUDA [N,K]
    weight W (this rank's row shard under sharding; the caller all-reduces)."""
    folded = isinstance(H, FoldedGram)
    Hm = H.Hb if folded else H
    assert W.is_cuda and W.dim() == 2 and Hm.shape == (W.shape[1], W.shape[1])
    import ctypes as C
    W = W.contiguous()
    N, K = W.shape
    n_cand = len(candidates)
    lib = _lib.load()
    loss = torch.zeros(n_cand, dtype=torch.float32, device=W.device)
    sf = (C.c_float * n_cand)(*[float(c) for c in candidates])
    mask = salient_mask.to(device=W.device, dtype=torch.uint8).contiguous()
    with _on(W.device):
        work = _workspace(W.device, lib.b200q_awq_search_workspace(N, K, n_cand))
        fn = lib.b200q_awq_search_loss_folded if folded else lib.b200q_awq_search_loss
        rc = fn(W.data_ptr(), N, K, group, n_bit, mask.data_ptr(), sf, n_cand,
                Hm.contiguous().data_ptr(), dtype_code(W), work.data_ptr(), loss.data_ptr(), _stream())
    _lib.check(rc, "awq_search_loss")
    return loss


class PreparedSearch:
    """dW_c for every candidate of one layer, sitting in a workspace (awq_search_prepare)."""
    __slots__ = ("work", "N", "K", "n_cand", "device")

    def __init__(self, work, N, K, n_cand, device):
        self.work, self.N, self.K, self.n_cand, self.device = work, N, K, n_cand, device


_search_work = {}


def awq_search_prepare(W: torch.Tensor, salient_mask: torch.Tensor, n_bit: int, group: int,
                       candidates: Sequence[float]) -> PreparedSearch:
    """First half of awq_search_losses: quantise W for every candidate and keep dW_c (bf16) in a
    per-device workspace.  Needs no Gram matrix, so it can run on a side stream while the Gram GEMM
    is still busy; awq_search_finish consumes it (the caller orders the two with events)."""
    assert W.is_cuda and W.dim() == 2
    import ctypes as C
    W = W.contiguous()
    N, K = W.shape
    n_cand = len(candidates)
    lib = _lib.load()
    nbytes = lib.b200q_awq_search_workspace(N, K, n_cand)
    buf = _search_work.get(W.device.index)
    if buf is None or buf.numel() < nbytes:
        buf = _search_work[W.device.index] = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    sf = (C.c_float * n_cand)(*[float(c) for c in candidates])
    mask = salient_mask.to(device=W.device, dtype=torch.uint8).contiguous()
    with _on(W.device):
        rc = lib.b200q_awq_search_delta(W.data_ptr(), N, K, group, n_bit, mask.data_ptr(), sf, n_cand,
                                        dtype_code(W), buf.data_ptr(), _stream())
    _lib.check(rc, "awq_search_delta")
    return PreparedSearch(buf, N, K, n_cand, W.device)


def awq_search_finish(p: PreparedSearch, H) -> torch.Tensor:
    """Second half: fp32 [n_cand] losses sum_rows dW_c H dW_c^T from a prepared workspace and the
    Gram matrix (fp32 [K,K], or a FoldedGram)."""
    folded = isinstance(H, FoldedGram)
    Hm = (H.Hb if folded else H).contiguous()
    assert Hm.shape == (p.K, p.K)
    loss = torch.zeros(p.n_cand, dtype=torch.float32, device=p.device)
    with _on(p.device):
        rc = _lib.load().b200q_awq_search_loss_prepared(p.N, p.K, p.n_cand, None if folded else Hm.data_ptr(),
                                                        Hm.data_ptr() if folded else None,
                                                        p.work.data_ptr(), loss.data_ptr(), _stream())
    _lib.check(rc, "awq_search_loss_prepared")
    return loss


def gptq_hessian(input_feat: Sequence, in_features: int, device, perp_damp: float = 0.01,
                 nsamples: int = 128, defer_exchange: bool = False):
    """The damped Hessian of gptq_quantizer.py:133-150 as fp32 [K,K] on `device`:
    sum over input_feat[:nsamples] of x^T x / (||x|| + 1e-5)^2, divided by len(input_feat) (the
    FULL list length, as the reference does), plus perp_damp * I.  A [n, K] tensor stands for n
    one-row samples (the reference iterates its rows); non-tensor features give I.

    defer_exchange (row sharding only): returns (H, event) instead of H -- the cross-rank sum of the
    partial matrices (packed lower triangle), the unpacking and the scale / damping are queued on
    the communication stream and `event` marks their completion, so the caller can go on with the
    next layer's partial Hessian while this one is exchanged.  event is None when nothing was
    deferred."""
    require_cuda()
    device = torch.device(device)
    K = in_features
    from .streaming import ActivationStream
    if isinstance(input_feat, ActivationStream):
        # batches were folded in as they were captured (streaming.py); `[:nsamples]` was applied by
        # the stream's max_batches, the divisor is every batch seen, as in the reference
        assert input_feat.normalize and input_feat.in_features == K
        Hs = hessian_finalize(input_feat.matrix_sum(device), 1.0 / max(1, input_feat.batches_seen),
                              perp_damp)
        return (Hs, None) if defer_exchange else Hs
    if isinstance(input_feat, torch.Tensor):
        feats_total = input_feat.shape[0]
        runs = [(input_feat[:nsamples].reshape(-1, K), 1)] if input_feat.dim() == 2 else \
            [(input_feat[:nsamples].reshape(-1, K), input_feat.shape[1])]
    else:
        feats_total = len(input_feat)
        if feats_total == 0 or not isinstance(input_feat[0], torch.Tensor):
            H = torch.eye(K, dtype=torch.float32, device=device)
            H = hessian_finalize(H, 1.0 / max(1, feats_total), perp_damp)
            return (H, None) if defer_exchange else H
        # group consecutive samples with the same number of rows into one call
        runs: List = []
        cur: List[torch.Tensor] = []
        cur_rows = None
        for f in input_feat[:nsamples]:
            f2 = f.reshape(1, -1) if f.dim() == 1 else f.reshape(-1, f.shape[-1])
            if cur_rows is not None and f2.shape[0] != cur_rows:
                runs.append((cur, cur_rows))
                cur = []
            cur_rows = f2.shape[0]
            cur.append(f2)
        if cur:
            runs.append((cur, cur_rows))
        runs = [(torch.cat([t.to(device, non_blocking=True) for t in ts]) if len(ts) > 1
                 else ts[0].to(device), r) for ts, r in runs]
    H = None
    # under row sharding the calibration samples are dealt to the ranks and the partial sums
    # all-reduced (b200q.dist); every rank ends with the same H
    world, rank = _dist.world_size(), _dist.rank()
    for X, rows in runs:
        X = X.to(device)
        if X.dtype not in DTYPE_CODE:
            X = X.float()
        n = X.shape[0] // rows
        if _dist.is_sharded():
            lo, hi = _dist.shard_rows(n, world, rank)
            X = X[lo * rows:hi * rows]
            if X.shape[0] == 0:
                continue
        H = hessian_accum(X, rows, H)
    if H is None:
        H = torch.zeros((K, K), dtype=torch.float32, device=device)
    if defer_exchange and _dist.is_sharded() and PACKED_EXCHANGE:
        main = torch.cuda.current_stream(device)
        comm = _comm_stream(device)
        P = sym_pack_lower(H)
        comm.wait_stream(main)
        with torch.cuda.stream(comm), _dist.on_comm_stream():
            _dist.allreduce_sum(P)
            sym_unpack_lower(P, K, out=H)
            hessian_finalize(H, 1.0 / feats_total, perp_damp)
            ev = torch.cuda.Event()
            ev.record(comm)
        P.record_stream(comm)
        H.record_stream(comm)
        return H, ev
    allreduce_symmetric(H)
    H = hessian_finalize(H, 1.0 / feats_total, perp_damp)
    return (H, None) if defer_exchange else H
