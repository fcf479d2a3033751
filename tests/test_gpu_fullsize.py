"""Size-independent properties at BASELINE.json's full layer shapes (the oracle is too slow there):
idempotence of the quantizers, code ranges, group-wise invariants, row-shard consistency."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(4096, 4096), (11008, 4096), (4096, 11008)]


@pytest.mark.parametrize("N,K", SHAPES)
def test_uniform_fakequant_properties(N, K):
    from b200q import ops
    g = torch.Generator(device="cuda").manual_seed(N + K)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    out, codes, scales, zeros = ops.group_fakequant(w, 4, 128, return_codes=True)
    assert codes.min() >= 0 and codes.max() <= 15
    # dequantisation identity, exactly as the reference writes it
    deq = (codes.float().view(-1, 128) - zeros[:, None]) * scales[:, None]
    assert torch.equal(deq.view(N, K), out)
    # every group's extremes are representable: error bounded by scale/2 (+ rounding)
    assert ((out - w).abs().view(-1, 128) <= scales[:, None] * 0.5001).all()
    # row-shard consistency: quantising a row block alone gives the same rows
    part = ops.group_fakequant(w[1000:1500].contiguous(), 4, 128)
    assert torch.equal(part, out[1000:1500])


@pytest.mark.parametrize("N,K", SHAPES)
def test_gptq_parity_properties(N, K):
    from b200q import ops
    g = torch.Generator(device="cuda").manual_seed(N * 3 + K)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    out, codes, scales = ops.gptq_parity_quant(w, 4, return_codes=True)
    assert codes.min() >= -16 and codes.max() <= 15
    # (torch-on-CUDA turns `x / 15` into `x * (1/15)`; the reference semantics are the CPU's true division)
    assert torch.equal(scales.cpu(), (w.cpu().abs().amax(0) / 15).clamp(min=1e-5))
    assert torch.equal(codes.float() * scales, out)
    # idempotent: the quantized matrix is a fixed point (column max is a grid point)
    assert torch.equal(ops.gptq_parity_quant(out, 4), out)
    # sharded: two row halves with a max-combined column statistic reproduce the full result
    cm = ops.col_absmax(w[: N // 2].contiguous())
    ops.col_absmax(w[N // 2:].contiguous(), out=cm, accumulate=True)
    top = ops.gptq_parity_quant(w[: N // 2].contiguous(), 4, colmax=cm)
    assert torch.equal(top, out[: N // 2])


@pytest.mark.parametrize("N,K", [(4096, 4096)])
def test_pot_apot_properties_full_matrix(N, K):
    from pot_apot_quantizer import pot_quantize_tensor, apot_quantize_tensor, _apot_signed_levels
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    q = pot_quantize_tensor(w, 4, 128)
    # every value is +-scale*2^E: per group at most 8 distinct magnitudes (+0), ratios powers of two
    mags = q.abs().view(-1, 128)
    top = mags.amax(1, keepdim=True)
    ratio = torch.where(mags > 0, top / mags, torch.ones_like(mags))
    assert torch.equal(ratio, torch.exp2(torch.round(torch.log2(ratio))))
    assert ratio.max() <= 128
    assert (torch.sign(q) == torch.sign(w)).all()
    from b200q import ops
    lv = _apot_signed_levels(4, 2)
    a = apot_quantize_tensor(w, 4, 128, 2)
    out, lidx, scale, idx = ops.apot_quant(w.view(-1, 128), lv, torch.arange(0.01, 2.01, 0.1),
                                           return_codes=True)
    assert torch.equal(out.view(N, K), a)
    assert lidx.max() < lv.numel() and idx.min() >= 0 and idx.max() < 20
    assert torch.equal(scale[:, None] * lv.cuda()[lidx.long()], out)
    # row-shard consistency for POT (APOT's grid depends on the global element count)
    assert torch.equal(pot_quantize_tensor(w[100:164].contiguous(), 4, 128), q[100:164])


@pytest.mark.parametrize("K", [4096, 11008])
def test_inverse_defining_identities_at_llama_sizes(K):
    """BASELINE sizes (Llama-2-7B in_features): several levels of the recursive factorisation, the
    tensor-core products with chunked accumulation, the CUDA-graph replay.  Size-independent
    properties: H^-1 H = I, symmetry, U upper triangular with U^T U = H^-1."""
    from b200q import tensor_ops as T
    g = torch.Generator(device="cuda").manual_seed(K)
    X = torch.randn(K + 512, K, device="cuda", generator=g)
    H = (X.T @ X) / X.shape[0]
    H += 0.01 * torch.diag(H).mean() * torch.eye(K, device="cuda")
    del X
    for _ in range(2):                                   # capture, then replay
        Hinv, U = T.spd_inverse(H, want_inverse=True, want_upper=True)
    eye = torch.eye(K, device="cuda")
    assert (Hinv @ H - eye).abs().max().item() < 2e-3    # fp32 check product; fp64 residual is ~1e-5
    assert torch.equal(Hinv, Hinv.T)
    assert torch.count_nonzero(U.tril(-1)).item() == 0
    rel = ((U.T @ U) - Hinv).abs().max() / Hinv.abs().max()
    assert rel.item() < 1e-4


def test_hessian_and_gram_at_llama_token_counts():
    """128 samples x 2048 tokens of bf16 activations (the bench's calibration set) at K = 4096:
    linearity in the sample set and agreement of the two tensor-core paths where they must agree."""
    from b200q import tensor_ops as T
    K, n, rows = 4096, 128, 2048
    g = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn(n * rows, K, device="cuda", generator=g, dtype=torch.bfloat16)
    G_all = T.hessian_accum(X, rows, normalize=False)
    half = n // 2 * rows
    G_two = T.hessian_accum(X[half:], rows, T.hessian_accum(X[:half], rows, normalize=False),
                            normalize=False)
    # The tensor core truncates every product at the accumulator's ulp, so a same-sign sum (the
    # diagonal) comes out low by ~N * 2^-24 per accumulation chain; the kernel folds 4096-token
    # chunks into a round-to-nearest running total, which bounds that at 2.4e-4 (one long chain of
    # 37k tokens measured 2.3e-3 low, and the two ways of splitting the samples 4.4e-4 apart).
    assert ((G_all - G_two).abs().max() / G_all.abs().max()).item() < 3e-4
    exact_diag = (X[:, :8].double() ** 2).sum(0)
    assert ((torch.diag(G_all)[:8].double() - exact_diag).abs() / exact_diag).max().item() < 4e-4
    assert torch.equal(G_all, G_all.T)
    # normalised Hessian: every sample has trace 1 (up to the 1e-5 in the denominator)
    H, norms = T.hessian_accum(X, rows, return_norms=True)
    assert abs(torch.diag(H).sum().item() - n) < 1e-2 * n
    # a sample's block of the Gram matrix, scaled by its norm, is its term of H
    H0 = T.hessian_accum(X[:rows], rows)
    G0 = T.hessian_accum(X[:rows], rows, normalize=False)
    want = G0 / (norms[0] + 1e-5) ** 2
    assert ((H0 - want).abs().max() / want.abs().max()).item() < 1e-5
