"""CPU oracle for the weight-quantization hot path of vimarsh244/llm-quantization.

TEST INFRASTRUCTURE ONLY.  This module restates, with torch CPU ops, the arithmetic of the
reference functions named in each docstring, and additionally returns the integer codes / scales /
zero points the reference keeps as temporaries.  It exists so that the CUDA path can be checked
bit-for-bit; nothing under llm-quantization_b200/ may import it.  Allowed importers: tests/,
__graft_entry__.smoke(), and the cpu_baseline / --impl reference legs of bench.py.

Pinning: oracle/gen_golden.py imports the unmodified reference from /root/reference, runs both on
the same seeded inputs, asserts torch.equal on every dequantized output and stores the vectors in
tests/golden/*.npz; tests/test_oracle_golden.py re-checks the oracle against those files wherever
it runs.  Rows whose reference body is a stub (awq_search_scale_factor, smoothquant_search_alpha)
or absent (error-compensated GPTQ) are marked PARITY UNPINNED below.

All functions take and return CPU tensors and follow torch's type promotion exactly as the
reference would for the given dtype (they are torch programs, not re-derivations).
"""
from __future__ import annotations

import itertools
from typing import Dict, Optional, Sequence

import torch

# --------------------------------------------------------------------------------------------------
# a7  pseudo_quantize_tensor            ref: quantization_utils.py:362-413
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def uniform_group_quant(w: torch.Tensor, n_bit: int = 4, group: int = -1) -> Dict[str, torch.Tensor]:
    shape, dtype = w.shape, w.dtype
    if group > 0:
        assert shape[-1] % group == 0
        w = w.reshape(-1, group)
    assert w.dim() == 2
    hi = w.amax(dim=1, keepdim=True)
    lo = w.amin(dim=1, keepdim=True)
    qmax = 2 ** n_bit - 1
    scales = (hi - lo).clamp(min=1e-5) / qmax
    zeros = (-torch.round(lo / scales)).clamp_(0, qmax)
    codes = torch.clamp(torch.round(w / scales) + zeros, 0, qmax)
    deq = (codes - zeros) * scales
    return {
        "out": deq.reshape(shape).to(dtype),
        "codes": codes.reshape(shape).to(torch.int32),
        "scales": scales.reshape(-1).float(),
        "zeros": zeros.reshape(-1).float(),
    }


# --------------------------------------------------------------------------------------------------
# a6  _simple_quantize_layer            ref: gptq_quantizer.py:79-108
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def symmetric_group_quant(w: torch.Tensor, n_bit: int, group: int) -> Dict[str, torch.Tensor]:
    shape, dtype = w.shape, w.dtype
    if group > 0:
        w = w.reshape(-1, group)
    qmax = 2 ** n_bit - 1
    scales = torch.clamp(w.abs().amax(dim=1, keepdim=True) / qmax, min=1e-5)
    codes = torch.clamp(torch.round(w / scales), -qmax - 1, qmax)
    return {
        "out": (codes * scales).reshape(shape).to(dtype),
        "codes": codes.reshape(shape).to(torch.int32),
        "scales": scales.reshape(-1).float(),
    }


# --------------------------------------------------------------------------------------------------
# a4  _gptq_quantize_layer column loop, closed form      ref: gptq_quantizer.py:167-206
# The reference quantizes column j with s_j = clamp(max_i|W[i,j]|/(2^b-1), 1e-5) and applies no
# error compensation, so the loop order, blocksize, permutation and H never reach the output.
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def gptq_parity_quant(W: torch.Tensor, n_bit: int) -> Dict[str, torch.Tensor]:
    dtype = W.dtype
    qmax = 2 ** n_bit - 1
    scales = torch.clamp(W.abs().amax(dim=0, keepdim=True) / qmax, min=1e-5)
    codes = torch.clamp(torch.round(W / scales), -qmax - 1, qmax)
    return {
        "out": (codes * scales).to(dtype),
        "codes": codes.to(torch.int32),
        "scales": scales.reshape(-1).float(),
    }


# --------------------------------------------------------------------------------------------------
# a2/a3  Hessian, damping, act-order, inverse            ref: gptq_quantizer.py:133-165
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def gptq_hessian(feats: Sequence[torch.Tensor], K: int, dtype: torch.dtype = torch.float32,
                 nsamples: int = 128, perp_damp: float = 0.01) -> torch.Tensor:
    H = torch.zeros(K, K, dtype=dtype)
    for x in feats[:nsamples]:
        if x.dim() == 1:
            x = x.unsqueeze(0)
        xn = x / (x.norm() + 1e-5)
        H += xn.T @ xn
    # note: divides by the FULL list length, damping is absolute (not x mean diag)   :150
    return H / len(feats) + perp_damp * torch.eye(K)


@torch.no_grad()
def gptq_perm(H: torch.Tensor, actorder: bool) -> torch.Tensor:
    if actorder:
        return torch.argsort(torch.diag(H), descending=True)
    return torch.arange(H.shape[0])


@torch.no_grad()
def gptq_hinv(H: torch.Tensor) -> torch.Tensor:
    return torch.linalg.inv(H + 1e-6 * torch.eye(H.shape[0], dtype=H.dtype))


# --------------------------------------------------------------------------------------------------
# a8  awq_quantize_model_weight, one Linear               ref: awq_quantizer.py:56-84
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def awq_layer(W: torch.Tensor, feats: Sequence[torch.Tensor], n_bit: int, group: int,
              protect_ratio: float = 0.01, scale_factor: float = 1.0) -> Dict[str, torch.Tensor]:
    importance = sum(feats).float()
    n_protect = max(1, int(len(importance) * protect_ratio))
    salient = torch.topk(importance, n_protect)[1]
    Ws = W.clone()
    Ws[:, salient] *= scale_factor
    q = uniform_group_quant(Ws, n_bit, group)
    out = q["out"]
    out[:, salient] /= scale_factor
    return {"out": out.to(W.dtype), "salient": salient, "importance": importance,
            "codes": q["codes"], "scales": q["scales"], "zeros": q["zeros"]}


# --------------------------------------------------------------------------------------------------
# a10 / a16  activation statistics     ref: quantization_utils.py:231, smooth_quant_quantizer.py:68-74
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def act_meanabs(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-1]).abs().mean(dim=0)


@torch.no_grad()
def act_maxabs(x: torch.Tensor, running: Optional[torch.Tensor] = None) -> torch.Tensor:
    m = x.reshape(-1, x.shape[-1]).abs().max(dim=0)[0]
    return m if running is None else torch.max(running, m)


# --------------------------------------------------------------------------------------------------
# a11  pot_quantize_tensor                                ref: pot_apot_quantizer.py:25-115
# --------------------------------------------------------------------------------------------------
def pot_grid() -> torch.Tensor:
    return torch.arange(0.01, 2.01, 0.01)  # :75 — 200 points, materialised by torch


@torch.no_grad()
def pot_quant(w: torch.Tensor, n_bit: int = 4, group: int = -1) -> Dict[str, torch.Tensor]:
    shape, dtype = w.shape, w.dtype
    if group > 0:
        assert shape[-1] % group == 0
        w = w.reshape(-1, group)
    assert w.dim() == 2
    top = 2 ** (n_bit - 1) - 1                      # E_max_idx
    tiny = torch.finfo(dtype).tiny
    aw = w.abs()
    sgn = torch.sign(w)
    peak = torch.clamp(aw.amax(dim=1, keepdim=True), min=1e-12)
    e_lo = torch.floor(torch.log2(peak)) - top
    s0 = torch.clamp(torch.pow(torch.tensor(2.0, dtype=dtype), e_lo.to(dtype)), min=tiny)

    def exponents(scale):
        return torch.clamp(torch.round(torch.log2(torch.clamp(aw / scale, min=1e-10))), 0, top)

    best_err = torch.full((w.size(0), 1), float("inf"))
    best_s = s0.clone()
    best_i = torch.full((w.size(0), 1), -1, dtype=torch.int32)
    for i, b in enumerate(pot_grid()):
        s = torch.clamp(s0 * b, min=tiny)
        wq = s * sgn * torch.pow(2.0, exponents(s))
        err = ((w - wq) ** 2).sum(dim=1, keepdim=True)
        better = err < best_err                      # strict: the first minimum wins
        best_err = torch.where(better, err, best_err)
        best_s = torch.where(better, s, best_s)
        best_i = torch.where(better, torch.tensor(i, dtype=torch.int32), best_i)
    best_s = torch.clamp(best_s, min=tiny)
    E = exponents(best_s)
    out = best_s * sgn * torch.pow(2.0, E)
    return {
        "out": out.reshape(shape).to(dtype),
        "exps": E.reshape(shape).to(torch.int32),
        "scale": best_s.reshape(-1).float(),
        "best_idx": best_i.reshape(-1),
    }


# --------------------------------------------------------------------------------------------------
# a12  generate_apot_levels                               ref: pot_apot_quantizer.py:138-188
# --------------------------------------------------------------------------------------------------
def apot_levels(n: int, k: int) -> torch.Tensor:
    terms = []
    for i in range(n):
        terms.append([0.0] + [2.0 ** (-(i + (j - 1) * n)) for j in range(1, 2 ** k)])
    sums = [sum(c) for c in itertools.product(*terms)]
    return torch.sort(torch.unique(torch.tensor(sums, dtype=torch.float32)))[0]


def apot_level_set(n_bit: int, k: int) -> torch.Tensor:
    """Signed, normalised level set incl. the 32-level cap.   ref: :224-247"""
    lv = apot_levels(max(1, n_bit // k), k)
    if lv.max() > 0:
        lv = lv / lv.max()
    pos = lv[lv > 0]
    full = torch.cat([-pos.flip(0), torch.tensor([0.0]), pos])
    if full.numel() > 32:
        full = full[torch.linspace(0, full.numel() - 1, 32, dtype=torch.long)]
    return full


def apot_grid(total_elements: int) -> torch.Tensor:
    """ref: :258-262 — the step depends on the WHOLE tensor's element count."""
    return torch.arange(0.01, 2.01, 0.1 if total_elements > 500000 else 0.05)


# --------------------------------------------------------------------------------------------------
# a13  apot_quantize_tensor                               ref: pot_apot_quantizer.py:192-351
# (the reference's column chunking / empty_cache calls have no numeric effect and are dropped)
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def apot_quant(w: torch.Tensor, n_bit: int = 4, group: int = -1, k: int = 2,
               total_elements: Optional[int] = None) -> Dict[str, torch.Tensor]:
    shape, dtype = w.shape, w.dtype
    if group > 0:
        assert shape[-1] % group == 0
        w = w.reshape(-1, group)
    assert w.dim() == 2
    levels = apot_level_set(n_bit, k)
    s0 = torch.clamp(w.abs().amax(dim=1, keepdim=True), min=1e-5)
    grid = apot_grid(w.numel() if total_elements is None else total_elements)

    def nearest(x):
        idx = torch.empty(x.shape, dtype=torch.long)
        step = max(1, (1 << 22) // max(1, x.shape[0] * levels.numel()))
        for c0 in range(0, x.shape[1], step):
            blk = x[:, c0:c0 + step]
            idx[:, c0:c0 + step] = torch.argmin(
                torch.abs(blk.unsqueeze(-1) - levels.view(1, 1, -1)), dim=-1)
        return idx

    def dequant_norm(x):
        idx = nearest(x)
        qn = torch.zeros_like(x)
        qn[:] = levels[idx]
        return qn, idx

    best_err = torch.full((w.size(0), 1), float("inf"))
    best_s = s0.clone()
    best_i = torch.full((w.size(0), 1), -1, dtype=torch.int32)
    for i, b in enumerate(grid):
        s = s0 * b
        qn, _ = dequant_norm(w / s)
        err = ((w - s * qn) ** 2).sum(dim=1, keepdim=True)
        better = err < best_err
        best_err = torch.where(better, err, best_err)
        best_s = torch.where(better, s, best_s)
        best_i = torch.where(better, torch.tensor(i, dtype=torch.int32), best_i)
    qn, idx = dequant_norm(w / best_s)
    return {
        "out": (best_s * qn).reshape(shape).to(dtype),
        "level_idx": idx.reshape(shape).to(torch.int32),
        "scale": best_s.reshape(-1).float(),
        "best_idx": best_i.reshape(-1),
        "levels": levels,
    }


# --------------------------------------------------------------------------------------------------
# a15  smooth_weights, one Linear                         ref: smooth_quant_quantizer.py:150-174
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def smooth_scale(act_scale: torch.Tensor, W: torch.Tensor, alpha: float) -> torch.Tensor:
    a = torch.clamp(act_scale, min=1e-5)
    wmax = torch.clamp(W.abs().max(dim=0)[0], min=1e-5)
    return torch.clamp(torch.pow(a, alpha) / torch.pow(wmax, 1.0 - alpha), min=1e-5)


@torch.no_grad()
def smooth_layer(W: torch.Tensor, act_scale: torch.Tensor, alpha: float) -> Dict[str, torch.Tensor]:
    s = smooth_scale(act_scale, W, alpha)
    return {"out": W / s, "s": s}


# a17  smoothquant_quantize_model_weight, one Linear      ref: smooth_quant_quantizer.py:301-320
@torch.no_grad()
def smoothquant_layer(W: torch.Tensor, act_scale: Optional[torch.Tensor], alpha: float, n_bit: int,
                      group: int, s: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """`s` overrides the computed smoothing scale (lets a test feed the CUDA path's own s, whose
    powf may differ from SLEEF's in the last ulp, and still demand bit-equality downstream)."""
    if act_scale is not None or s is not None:
        if s is None:
            s = smooth_scale(act_scale, W, alpha)
        W = W / s
    q = uniform_group_quant(W, n_bit, group)
    q["s"] = s
    return q


# --------------------------------------------------------------------------------------------------
# PARITY UNPINNED rows — the reference body is a stub; these restate the docstrings / papers.
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def awq_search_losses(W: torch.Tensor, H: torch.Tensor, salient: torch.Tensor, n_bit: int, group: int,
                      candidates: Sequence[float]) -> torch.Tensor:
    """awq_quantizer.py:116-119 ("for each scale factor, quantize and measure reconstruction
    error"): loss_c = || (Q_c(W) - W) X^T ||_F^2 = tr(dW H dW^T), H = X^T X.  float64 accumulate."""
    Hd = H.double()
    losses = []
    for sf in candidates:
        Ws = W.clone()
        Ws[:, salient] *= sf
        q = uniform_group_quant(Ws, n_bit, group)["out"]
        q[:, salient] /= sf
        d = (q - W).double()
        losses.append(((d @ Hd) * d).sum())
    return torch.stack(losses)


@torch.no_grad()
def gptq_compensated(W: torch.Tensor, H: torch.Tensor, n_bit: int, group: int,
                     blocksize: int = 128, perm: Optional[torch.Tensor] = None,
                     return_margin: bool = False):
    """GPTQ (Frantar et al. 2022, Alg. 1) with the asymmetric per-group grid of
    pseudo_quantize_tensor: the loop gptq_quantizer.py:173-197 sketches and then skips.
    H must already include the damping.  float64 reference arithmetic.
    return_margin: also return, per element, the distance of the (compensated) value from the
    nearest rounding boundary in code units, 0 = exactly on a tie -- lets a test tell a legitimate
    tie flip of an fp32 implementation from an error."""
    Wd = W.double().clone()
    N, K = Wd.shape
    G = group if group > 0 else K
    if perm is not None:
        inv = torch.argsort(perm)
        res = gptq_compensated(W[:, perm], H[perm][:, perm], n_bit, group, blocksize,
                               return_margin=return_margin)
        return (res[0][:, inv], res[1][:, inv]) if return_margin else res[:, inv]
    Hinv = torch.linalg.inv(H.double())
    U = torch.linalg.cholesky(Hinv, upper=True)
    qmax = 2 ** n_bit - 1
    Q = torch.zeros_like(Wd)
    margin = torch.zeros_like(Wd)
    scale = zero = None
    for c0 in range(0, K, blocksize):
        c1 = min(c0 + blocksize, K)
        Err = torch.zeros(N, c1 - c0, dtype=torch.float64)
        for j in range(c0, c1):
            if j % G == 0:
                blk = Wd[:, j:j + G]
                hi, lo = blk.amax(dim=1), blk.amin(dim=1)
                scale = (hi - lo).clamp(min=1e-5) / qmax
                zero = (-torch.round(lo / scale)).clamp(0, qmax)
            col = Wd[:, j]
            x = col / scale
            margin[:, j] = (x - torch.floor(x) - 0.5).abs()
            qc = (torch.clamp(torch.round(x) + zero, 0, qmax) - zero) * scale
            Q[:, j] = qc
            e = (col - qc) / U[j, j]
            Wd[:, j + 1:c1] -= e.unsqueeze(1) * U[j, j + 1:c1].unsqueeze(0)
            Err[:, j - c0] = e
        Wd[:, c1:] -= Err @ U[c0:c1, c1:]
    return (Q.to(W.dtype), margin) if return_margin else Q.to(W.dtype)


# --------------------------------------------------------------------------------------------------
# packed integer export.  The reference stores no codes, so there is
# nothing to pin against: this is the specification of the layout (PARITY UNPINNED by construction).
# --------------------------------------------------------------------------------------------------
def pack_codes(codes, n_bit: int):
    """uint8 codes [N, K] -> uint32 words [N, ceil(K*n_bit/32)]: each row is a little-endian bit
    stream with code k at bits [k*n_bit, (k+1)*n_bit)."""
    import numpy as np
    c = np.asarray(codes, dtype=np.uint64) & ((1 << n_bit) - 1)
    N, K = c.shape
    words = (K * n_bit + 31) // 32
    out = np.zeros((N, words), dtype=np.uint64)
    for k in range(K):
        bit = k * n_bit
        w, off = bit // 32, bit % 32
        out[:, w] |= (c[:, k] << off) & 0xFFFFFFFF
        if off + n_bit > 32:
            out[:, w + 1] |= c[:, k] >> (32 - off)
    return out.astype(np.uint32)


def unpack_codes(words, K: int, n_bit: int):
    import numpy as np
    w = np.asarray(words, dtype=np.uint64)
    N = w.shape[0]
    out = np.zeros((N, K), dtype=np.uint8)
    for k in range(K):
        bit = k * n_bit
        i, off = bit // 32, bit % 32
        v = w[:, i] >> off
        if off + n_bit > 32:
            v = v | (w[:, i + 1] << (32 - off))
        out[:, k] = (v & ((1 << n_bit) - 1)).astype(np.uint8)
    return out


# --------------------------------------------------------------------------------------------------
# a18  smoothquant_search_alpha's measure      ref: smooth_quant_quantizer.py:327-371 (a stub there)
# PARITY UNPINNED: the reference returns the midpoint; this restates the measure the drop-in uses.
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def smooth_alpha_errors(W: torch.Tensor, S: torch.Tensor, act_weight: torch.Tensor, n_bit: int,
                        group: int) -> torch.Tensor:
    """fp64 [n_alpha]: sum ((Q(W / S[a]) * S[a] - W) * act_weight)^2 with Q = uniform_group_quant."""
    out = []
    for s in S:
        smoothed = W / s.to(W.dtype)
        q = uniform_group_quant(smoothed, n_bit, group)["out"]
        back = q * s.to(W.dtype)
        err = (back.double() - W.double()) * act_weight.double()
        out.append((err ** 2).sum())
    return torch.stack(out)
