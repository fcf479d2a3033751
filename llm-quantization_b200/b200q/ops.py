"""Tensor-level wrappers over the C ABI: torch owns the memory and the stream, the library does the
arithmetic.  Every function here launches sm_100a kernels; none has a torch/CPU fallback.

Inputs that live on the host are copied to the current CUDA device, processed there, and the
result is returned on the input's device (the reference's own tests hand CPU tensors to the
tensor-level functions) — the arithmetic still runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
COLOP_NONE, COLOP_MUL_DIV, COLOP_DIV = 0, 1, 2


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "b200q: no CUDA device visible. This framework runs its quantization arithmetic in "
            "sm_100a kernels only; there is no CPU path.")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"b200q: unsupported dtype {t.dtype} (float32/float16/bfloat16 only)")


def to_device(t: torch.Tensor) -> torch.Tensor:
    """Contiguous CUDA view/copy of t."""
    require_cuda()
    if not t.is_cuda:
        t = t.cuda()
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def _on(device):
    """Device guard that costs nothing when `device` is already current (the per-layer calls of a
    model walker are host-overhead sensitive: ~80 us/layer of Python dwarfs a 20 us kernel)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


# --------------------------------------------------------------------------------------------------
def col_absmax(W: torch.Tensor, out: Optional[torch.Tensor] = None,
               accumulate: bool = False) -> torch.Tensor:
    """max_i |W[i,k]| as fp32 [K].  W: CUDA [N,K] (row stride may exceed K)."""
    assert W.dim() == 2 and W.is_cuda and W.stride(1) == 1
    N, K = W.shape
    if out is None:
        out = torch.empty(K, dtype=torch.float32, device=W.device)
        accumulate = False
    with _on(W.device):
        rc = _lib.load().b200q_col_absmax(W.data_ptr(), N, K, W.stride(0) if N > 1 else K,
                                          dtype_code(W), out.data_ptr(), int(accumulate), _stream())
    _lib.check(rc, "col_absmax")
    return out


def gptq_parity_quant(W: torch.Tensor, n_bit: int, colmax: Optional[torch.Tensor] = None,
                      return_codes: bool = False):
    """Reference-parity GPTQ column stage on a CUDA [N,K] matrix.  Returns out or
    (out, codes int8 [N,K], scales fp32 [K])."""
    assert W.dim() == 2 and W.is_cuda
    W = W.contiguous()
    N, K = W.shape
    if colmax is None:
        colmax = col_absmax(W)
    out = torch.empty_like(W)
    codes = torch.empty((N, K), dtype=torch.int8, device=W.device) if return_codes else None
    scales = torch.empty(K, dtype=torch.float32, device=W.device) if return_codes else None
    with _on(W.device):
        rc = _lib.load().b200q_gptq_parity_quant(W.data_ptr(), out.data_ptr(), _ptr(codes),
                                                 colmax.data_ptr(), _ptr(scales), N, K, K, n_bit,
                                                 dtype_code(W), _stream())
    _lib.check(rc, "gptq_parity_quant")
    return (out, codes, scales) if return_codes else out


def group_fakequant(W: torch.Tensor, n_bit: int, group: int, symmetric: bool = False,
                    colop: int = COLOP_NONE, colvec: Optional[torch.Tensor] = None,
                    return_codes: bool = False):
    """Uniform group fake-quant of a CUDA tensor viewed as [-1, K] with K = last dim."""
    assert W.is_cuda
    W = W.contiguous()
    K = W.shape[-1]
    N = W.numel() // K if K > 0 else 0
    G = group if group > 0 else K
    out = torch.empty_like(W)
    n_groups = N * (K // G) if K % G == 0 else 0
    codes = scales = zeros = None
    if return_codes:
        codes = torch.empty(W.shape, dtype=torch.int8 if symmetric else torch.uint8, device=W.device)
        scales = torch.empty(n_groups, dtype=torch.float32, device=W.device)
        zeros = torch.empty(n_groups, dtype=torch.float32, device=W.device)
    if colvec is not None:
        colvec = _f32(colvec, W.device)
        assert colvec.numel() == K
    with _on(W.device):
        rc = _lib.load().b200q_group_fakequant(W.data_ptr(), out.data_ptr(), _ptr(codes),
                                               _ptr(scales), _ptr(zeros), N, K, group, n_bit,
                                               int(symmetric), colop, _ptr(colvec), dtype_code(W),
                                               _stream())
    _lib.check(rc, "group_fakequant")
    return (out, codes, scales, zeros) if return_codes else out


def smooth_scale(act_scale: torch.Tensor, wmax: torch.Tensor, alpha: float, act_dtype: torch.dtype,
                 w_dtype: torch.dtype) -> torch.Tensor:
    """SmoothQuant s[k] as fp32 (values already rounded to the promoted dtype)."""
    K = wmax.numel()
    a = _f32(act_scale, wmax.device)
    assert a.numel() == K, "act_scales length does not match in_features"
    s = torch.empty(K, dtype=torch.float32, device=wmax.device)
    with _on(wmax.device):
        rc = _lib.load().b200q_smooth_scale(a.data_ptr(), wmax.data_ptr(), s.data_ptr(), K,
                                            float(alpha), DTYPE_CODE[act_dtype], DTYPE_CODE[w_dtype],
                                            _stream())
    _lib.check(rc, "smooth_scale")
    return s


def col_scale(W: torch.Tensor, s: torch.Tensor, mul: bool = False) -> torch.Tensor:
    """W / s[k] (or W * s[k]) per input column, CUDA [N,K]."""
    assert W.dim() == 2 and W.is_cuda
    W = W.contiguous()
    N, K = W.shape
    s = _f32(s, W.device)
    out = torch.empty_like(W)
    with _on(W.device):
        rc = _lib.load().b200q_col_scale(W.data_ptr(), out.data_ptr(), s.data_ptr(), N, K, int(mul),
                                         dtype_code(W), _stream())
    _lib.check(rc, "col_scale")
    return out


def act_meanabs(X: torch.Tensor) -> torch.Tensor:
    """mean over tokens of |x| per channel: X CUDA [..., K] -> fp32 [K]."""
    assert X.is_cuda
    X = X.contiguous()
    K = X.shape[-1]
    T = X.numel() // K
    lib = _lib.load()
    out = torch.empty(K, dtype=torch.float32, device=X.device)
    work = torch.empty(lib.b200q_act_stat_workspace(T, K), dtype=torch.uint8, device=X.device)
    with _on(X.device):
        rc = lib.b200q_act_meanabs(X.data_ptr(), T, K, dtype_code(X), out.data_ptr(),
                                   work.data_ptr(), _stream())
    _lib.check(rc, "act_meanabs")
    return out


def act_meanabs_batched(X: torch.Tensor) -> torch.Tensor:
    """Per-sample mean|x| of a CUDA [n_samples, tokens, K] batch -> fp32 [n_samples, K]."""
    assert X.is_cuda and X.dim() == 3
    X = X.contiguous()
    n, T, K = X.shape
    lib = _lib.load()
    out = torch.empty((n, K), dtype=torch.float32, device=X.device)
    work = torch.empty(n * lib.b200q_act_stat_workspace(T, K), dtype=torch.uint8, device=X.device)
    with _on(X.device):
        rc = lib.b200q_act_meanabs_batched(X.data_ptr(), n, T, K, dtype_code(X), out.data_ptr(),
                                           work.data_ptr(), _stream())
    _lib.check(rc, "act_meanabs_batched")
    return out


def act_maxabs(X: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """max over tokens of |x| per channel; with `out` given, a running max into it."""
    assert X.is_cuda
    X = X.contiguous()
    K = X.shape[-1]
    T = X.numel() // K
    accumulate = out is not None
    if out is None:
        out = torch.empty(K, dtype=torch.float32, device=X.device)
    with _on(X.device):
        rc = _lib.load().b200q_act_maxabs(X.data_ptr(), T, K, dtype_code(X), out.data_ptr(),
                                          int(accumulate), _stream())
    _lib.check(rc, "act_maxabs")
    return out


def seq_sum_rows(V: torch.Tensor) -> torch.Tensor:
    """Python-sum() order: ((0+V[0])+V[1])+... over the rows of a CUDA [n,K] matrix, every partial
    sum rounded to V's dtype; returned as fp32."""
    assert V.is_cuda and V.dim() == 2
    V = V.contiguous()
    n, K = V.shape
    out = torch.empty(K, dtype=torch.float32, device=V.device)
    with _on(V.device):
        rc = _lib.load().b200q_seq_sum_rows(V.data_ptr(), n, K, dtype_code(V), out.data_ptr(),
                                            _stream())
    _lib.check(rc, "seq_sum_rows")
    return out


def _host_floats(t: torch.Tensor):
    t = t.detach().to("cpu", torch.float32).contiguous()
    arr = (C.c_float * t.numel())(*t.tolist())
    return arr, t.numel()


def pot_quant(w: torch.Tensor, n_bit: int, grid: torch.Tensor, return_codes: bool = False):
    """POT quantisation of CUDA [n_groups, G]; grid = the host-materialised candidate multipliers."""
    assert w.is_cuda and w.dim() == 2
    w = w.contiguous()
    n_groups, G = w.shape
    out = torch.empty_like(w)
    exps = scale = idx = None
    if return_codes:
        exps = torch.empty((n_groups, G), dtype=torch.uint8, device=w.device)
        scale = torch.empty(n_groups, dtype=torch.float32, device=w.device)
        idx = torch.empty(n_groups, dtype=torch.int32, device=w.device)
    garr, n_grid = _host_floats(grid)
    with _on(w.device):
        rc = _lib.load().b200q_pot_quant(w.data_ptr(), out.data_ptr(), _ptr(exps), _ptr(scale),
                                         _ptr(idx), n_groups, G, n_bit, garr, n_grid,
                                         dtype_code(w), _stream())
    _lib.check(rc, "pot_quant")
    return (out, exps, scale, idx) if return_codes else out


def apot_quant(w: torch.Tensor, levels: torch.Tensor, grid: torch.Tensor,
               return_codes: bool = False):
    """APOT quantisation of CUDA [n_groups, G] against the signed level set `levels`."""
    assert w.is_cuda and w.dim() == 2
    w = w.contiguous()
    n_groups, G = w.shape
    out = torch.empty_like(w)
    lidx = scale = idx = None
    if return_codes:
        lidx = torch.empty((n_groups, G), dtype=torch.uint8, device=w.device)
        scale = torch.empty(n_groups, dtype=torch.float32, device=w.device)
        idx = torch.empty(n_groups, dtype=torch.int32, device=w.device)
    larr, n_levels = _host_floats(levels)
    garr, n_grid = _host_floats(grid)
    with _on(w.device):
        rc = _lib.load().b200q_apot_quant(w.data_ptr(), out.data_ptr(), _ptr(lidx), _ptr(scale),
                                          _ptr(idx), n_groups, G, larr, n_levels, garr, n_grid,
                                          dtype_code(w), _stream())
    _lib.check(rc, "apot_quant")
    return (out, lidx, scale, idx) if return_codes else out


# --------------------------------------------------------------------------------------------------
# whole-layer calls: one host call per nn.Linear
# --------------------------------------------------------------------------------------------------
_scratch = {}


def _work(device, n_floats: int) -> torch.Tensor:
    """Per-device fp32 scratch, grown on demand (stream-ordered reuse on the current stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < n_floats:
        buf = torch.empty(max(n_floats, 1 << 16), dtype=torch.float32, device=device)
        _scratch[key] = buf
    return buf


def awq_layer(W: torch.Tensor, feats: torch.Tensor, n_bit: int, group: int, n_protect: int,
              scale_factor: float, return_mask: bool = False):
    """importance -> top-k -> fused scale / asymmetric group fake-quant / unscale of a CUDA [N,K]
    weight; feats is a CUDA [n,K] matrix of per-batch mean|x| rows."""
    assert W.is_cuda and W.dim() == 2 and feats.is_cuda and feats.dim() == 2
    W = W.contiguous()
    feats = feats.contiguous()
    N, K = W.shape
    assert feats.shape[1] == K
    out = torch.empty_like(W)
    mask = torch.empty(K, dtype=torch.uint8, device=W.device) if return_mask else None
    with _on(W.device):
        work = _work(W.device, 3 * K)
        rc = _lib.load().b200q_awq_layer(W.data_ptr(), out.data_ptr(), N, K, group, n_bit,
                                         feats.data_ptr(), feats.shape[0], dtype_code(feats),
                                         n_protect, float(scale_factor), work.data_ptr(), _ptr(mask),
                                         dtype_code(W), _stream())
    _lib.check(rc, "awq_layer")
    return (out, mask) if return_mask else out


def salient_mask(feats: torch.Tensor, n_protect: int) -> torch.Tensor:
    """uint8 [K] mask of the n_protect channels with the largest summed statistic (the selection of
    awq_quantizer.py:57-61) for a CUDA [n,K] matrix of per-batch rows."""
    assert feats.is_cuda and feats.dim() == 2
    K = feats.shape[1]
    imp = seq_sum_rows(feats)
    colmul = torch.empty(K, dtype=torch.float32, device=feats.device)
    mask = torch.empty(K, dtype=torch.uint8, device=feats.device)
    with _on(feats.device):
        rc = _lib.load().b200q_topk_colmul(imp.data_ptr(), K, n_protect, 2.0, colmul.data_ptr(),
                                           mask.data_ptr(), _stream())
    _lib.check(rc, "topk_colmul")
    return mask


def gptq_parity_layer(W: torch.Tensor, n_bit: int) -> torch.Tensor:
    assert W.is_cuda and W.dim() == 2
    W = W.contiguous()
    N, K = W.shape
    out = torch.empty_like(W)
    with _on(W.device):
        work = _work(W.device, 2 * K)
        rc = _lib.load().b200q_gptq_parity_layer(W.data_ptr(), out.data_ptr(), N, K, n_bit,
                                                 work.data_ptr(), dtype_code(W), _stream())
    _lib.check(rc, "gptq_parity_layer")
    return out


def smoothquant_layer(W: torch.Tensor, act_scale: torch.Tensor, alpha: float, n_bit: int,
                      group: int):
    """(quantized smoothed weight, s fp32 [K]) for a CUDA [N,K] weight and fp32/16 act_scale [K]
    of the same dtype family as W (mixed dtypes go through the separate calls)."""
    assert W.is_cuda and W.dim() == 2
    W = W.contiguous()
    N, K = W.shape
    a = _f32(act_scale, W.device)
    assert a.numel() == K, "act_scales length does not match in_features"
    out = torch.empty_like(W)
    s = torch.empty(K, dtype=torch.float32, device=W.device)
    with _on(W.device):
        work = _work(W.device, 2 * K)
        rc = _lib.load().b200q_smoothquant_layer(W.data_ptr(), out.data_ptr(), N, K, group, n_bit,
                                                 a.data_ptr(), float(alpha),
                                                 DTYPE_CODE.get(act_scale.dtype, 0), s.data_ptr(),
                                                 work.data_ptr(), dtype_code(W), _stream())
    _lib.check(rc, "smoothquant_layer")
    return out, s


def smooth_alpha_errors(W: torch.Tensor, S: torch.Tensor, act_weight: torch.Tensor, n_bit: int,
                        group: int, err: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp64 [n_alpha] on the device: for every row S[a] of smoothing scales, the squared error of
    the smoothed-then-quantized weight mapped back and weighted by act_weight (see the header).
    Adds to `err` when given, so a model walk never synchronises."""
    assert W.is_cuda and W.dim() == 2 and S.dim() == 2 and S.shape[1] == W.shape[1]
    W = W.contiguous()
    N, K = W.shape
    S = _f32(S, W.device)
    a = _f32(act_weight, W.device)
    n_alpha = S.shape[0]
    accumulate = err is not None
    if err is None:
        err = torch.empty(n_alpha, dtype=torch.float64, device=W.device)
    lib = _lib.load()
    nbytes = lib.b200q_smooth_alpha_workspace(N, K, group, n_alpha)
    assert nbytes > 0, "in_features not divisible by the group size"
    work = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    with _on(W.device):
        rc = lib.b200q_smooth_alpha_errors(W.data_ptr(), N, K, group, n_bit, S.data_ptr(), n_alpha,
                                           a.data_ptr(), dtype_code(W), work.data_ptr(),
                                           err.data_ptr(), int(accumulate), _stream())
    _lib.check(rc, "smooth_alpha_errors")
    return err
