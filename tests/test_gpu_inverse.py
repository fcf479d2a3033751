"""Damped SPD inverse (blocked Cholesky route) vs fp64 LAPACK and vs the reference's golden H^-1
(gptq_quantizer.py:160-165).  cond(H) <= ~100 by construction; fp32 arithmetic -> 1e-4 of the
largest entry (measured ~1e-6)."""
import pytest
import torch

from conftest import case_dtype
from oracle import quant_oracle as O

pytestmark = pytest.mark.gpu


def spd(K, seed, damp=0.01):
    g = torch.Generator().manual_seed(seed)
    chan = torch.ones(K)
    chan[torch.randperm(K, generator=g)[: max(1, K // 50)]] = 20.0
    feats = [torch.randn(96, K, generator=g) * chan for _ in range(6)]
    return O.gptq_hessian(feats, K, torch.float32, 128, damp)


def rel(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max()).item()


@pytest.mark.parametrize("K", [64, 128, 200, 256, 384, 1000, 1001, 1544, 2048])
def test_inverse_and_upper_factor_vs_fp64(K):
    from b200q import tensor_ops as T
    H = spd(K, K)
    Hinv, U = T.spd_inverse(H.cuda(), want_inverse=True, want_upper=True)
    want = torch.linalg.inv(H.double())
    assert rel(Hinv.cpu(), want) < 1e-4
    want_U = torch.linalg.cholesky(want, upper=True)
    assert rel(U.cpu(), want_U) < 1e-4
    assert torch.equal(U.cpu().tril(-1), torch.zeros(K, K)), "U must be upper triangular"
    # defining identities
    eye = torch.eye(K, dtype=torch.float64)
    assert ((H.double() @ Hinv.cpu().double()) - eye).abs().max() < 1e-3
    assert rel(U.cpu().double().T @ U.cpu().double(), want) < 1e-4


def test_gptq_inverse_matches_reference_golden(golden):
    import gptq_quantizer as gq
    g = golden("gptq")
    for case in g.cases("gptq"):
        if case_dtype(case) != torch.float32:
            continue
        H_reg = g.tensor(f"gptq/{case}/H_reg")
        K = H_reg.shape[0]
        H = (H_reg - 1e-6 * torch.eye(K)).cuda()
        Hinv = gq.gptq_inverse(H)                      # adds the 1e-6 ridge itself (:160)
        assert rel(Hinv.cpu(), g.tensor(f"gptq/{case}/H_inv")) < 2e-4, case


def test_non_spd_is_reported():
    import ctypes
    from b200q import _lib
    K = 128
    H = torch.eye(K, device="cuda")
    H[5, 5] = -1.0
    lib = _lib.load()
    work = torch.empty(lib.b200q_spd_inverse_workspace(K), dtype=torch.uint8, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.empty_like(H)
    rc = lib.b200q_spd_inverse(H.data_ptr(), out.data_ptr(), None, K, work.data_ptr(), info.data_ptr(),
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and int(info.item()) == 6


def test_replayed_graph_sees_new_data():
    """The factorisation of a given size is recorded into a CUDA graph on first use and replayed:
    a second, different matrix of the same size must give ITS inverse, and a non-positive pivot
    deep inside the recursion (past the 512-column leaves) must still be reported."""
    from b200q import tensor_ops as T
    K = 1200
    for seed in (1, 2, 3):
        H = spd(K, seed)
        Hinv = T.spd_inverse(H.cuda())
        assert rel(Hinv.cpu(), torch.linalg.inv(H.double())) < 1e-4, seed
    import ctypes
    from b200q import _lib
    lib = _lib.load()
    H = torch.eye(K, device="cuda")
    H[900, 900] = -2.0
    work = torch.empty(lib.b200q_spd_inverse_workspace(K), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(H)
    for _ in range(2):                                   # capture, then replay
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = lib.b200q_spd_inverse(H.data_ptr(), out.data_ptr(), None, K, work.data_ptr(), info.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
        assert rc == 0 and int(info.item()) == 901
