"""GPTQ Hessian on the tensor cores (tcgen05) vs the oracle (gptq_quantizer.py:133-150) and the
golden H the reference built.  fp32 accumulation of exactly representable fp16 products: the only
error sources are the fp16 rounding of the normalised activations (2^-12 relative per element) and
summation order, so H is compared at 2e-3 relative to its largest entry (measured: ~1e-4)."""
import pytest
import torch

from conftest import case_dtype
from oracle import quant_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max()).item()


def make_feats(seed, n, rows, K, dtype=torch.float32, outliers=True):
    g = torch.Generator().manual_seed(seed)
    chan = torch.ones(K)
    if outliers:
        chan[torch.randperm(K, generator=g)[: max(1, K // 100)]] = 20.0
    return [(torch.randn(rows, K, generator=g) * chan).to(dtype) for _ in range(n)]


@pytest.mark.parametrize("n,rows,K", [(4, 64, 256), (3, 128, 128), (8, 200, 384), (5, 72, 1000),
                                      (16, 512, 768), (2, 2048, 1024), (128, 1, 512)])
def test_hessian_sum_vs_fp64(n, rows, K):
    from b200q import tensor_ops as T
    feats = make_feats(n * 1000 + K, n, rows, K)
    X = torch.cat(feats).cuda()
    H, norms = T.hessian_accum(X, rows, return_norms=True)
    want = torch.zeros(K, K, dtype=torch.float64)
    for f in feats:
        fn = f.double() / (f.double().norm() + 1e-5)
        want += fn.T @ fn
    assert rel_err(H.cpu(), want) < 5e-4
    assert torch.allclose(H, H.T, rtol=0, atol=0), "H must be exactly symmetric"
    torch.testing.assert_close(norms.cpu(), torch.stack([f.double().norm() for f in feats]).float(), rtol=2e-6, atol=0)
    # accumulate=True adds to the existing contents
    H2 = T.hessian_accum(X, rows, H.clone())
    assert rel_err(H2.cpu(), 2 * want) < 5e-4


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_hessian_16bit_inputs(dtype):
    from b200q import tensor_ops as T
    feats = make_feats(77, 6, 256, 512, dtype)
    H = T.hessian_accum(torch.cat(feats).cuda(), 256)
    want = torch.zeros(512, 512, dtype=torch.float64)
    for f in feats:
        fn = f.double() / (f.double().norm() + 1e-5)
        want += fn.T @ fn
    assert rel_err(H.cpu(), want) < 5e-4


def test_gptq_hessian_matches_reference_golden(golden):
    """Same feature lists the reference saw (2-D and 1-D, nsamples truncation, full-length divisor)."""
    import gptq_quantizer as gq
    g = golden("gptq")
    for case in g.cases("gptq"):
        dt = case_dtype(case)
        if dt != torch.float32:
            continue            # the reference accumulates H in fp16/bf16 there: not a target
        b, ns, act = (int(v) for v in g.arr(f"gptq/{case}/meta"))
        feats = list(g.tensor(f"gptq/{case}/feats"))
        K = feats[0].shape[-1]
        H = gq.gptq_hessian(feats, K, "cuda", 0.01, ns)
        want = g.tensor(f"gptq/{case}/H_reg") - 1e-6 * torch.eye(K)
        assert rel_err(H.cpu(), want) < 1e-3, case
        if act:
            # act-order permutation from diag(H): same ordering wherever the gaps exceed the error
            perm = torch.argsort(torch.diag(H), descending=True).cpu()
            want_perm = g.tensor(f"gptq/{case}/perm")
            assert (perm[:4] == want_perm[:4]).all()


def test_gptq_hessian_ragged_and_identity():
    import gptq_quantizer as gq
    K = 256
    feats = make_feats(5, 3, 64, K) + make_feats(6, 2, 96, K)
    H = gq.gptq_hessian(feats, K, "cuda", 0.01, 128)
    want = O.gptq_hessian(feats, K, torch.float32, 128, 0.01)
    assert rel_err(H.cpu(), want) < 1e-3
    H = gq.gptq_hessian(["not a tensor"] * 4, K, "cuda", 0.01, 128)
    assert torch.equal(H.cpu(), torch.eye(K) / 4 + 0.01 * torch.eye(K))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,rows,K", [(4, 256, 512), (3, 100, 392), (1, 4096, 1024)])
def test_gram_of_16bit_activations_reads_them_directly(dtype, n, rows, K):
    """normalize=False on fp16/bf16 input skips the staging pass: products of 16-bit inputs are
    exact in fp32, so the only error left is the fp32 summation order (1e-5 of the largest entry)."""
    from b200q import tensor_ops as T
    from b200q import _lib
    feats = make_feats(5 + K, n, rows, K, dtype)
    X = torch.cat(feats).cuda()
    _lib.profile_enable(True)
    H = T.hessian_accum(X, rows, normalize=False)
    torch.cuda.synchronize()
    ran = {k: _lib.profile_query(k)["launches"] for k in ("hessian_gemm", "hessian_prescale")}
    _lib.profile_enable(False)
    assert ran["hessian_gemm"] >= 1 and ran["hessian_prescale"] == 0, ran
    want = X.cpu().double().T @ X.cpu().double()
    assert rel_err(H.cpu(), want) < 1e-5
    assert torch.equal(H, H.T)
    H2 = T.hessian_accum(X, rows, H.clone(), normalize=False)
    assert rel_err(H2.cpu(), 2 * want) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n,rows,K", [(4, 512, 512), (3, 2048, 384), (9, 1024, 1000), (1, 512, 256)])
def test_per_sample_path_for_16bit_activations(dtype, n, rows, K):
    """Normalised Hessian of 16-bit samples of >= 512 rows: no fp16 staging copy -- the kernel
    accumulates one sample at a time and folds it in with weight 1/(||x||+1e-5)^2.  Products are
    exact, so only fp32 summation error is left."""
    from b200q import tensor_ops as T
    from b200q import _lib
    feats = make_feats(31 + K + n, n, rows, K, dtype)
    X = torch.cat(feats).cuda()
    H, norms = T.hessian_accum(X, rows, return_norms=True)
    want = torch.zeros(K, K, dtype=torch.float64)
    for f in feats:
        fn = f.double() / (f.double().norm() + 1e-5)
        want += fn.T @ fn
    assert rel_err(H.cpu(), want) < 2e-5
    assert torch.equal(H, H.T)
    torch.testing.assert_close(norms.cpu(), torch.stack([f.double().norm() for f in feats]).float(),
                               rtol=2e-6, atol=0)
    H2 = T.hessian_accum(X, rows, H.clone())
    assert rel_err(H2.cpu(), 2 * want) < 2e-5
