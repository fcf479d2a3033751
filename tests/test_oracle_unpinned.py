"""The oracle rows the reference cannot pin (its bodies are stubs: awq_quantizer.py:116-126,
gptq_quantizer.py:189-194, smooth_quant_quantizer.py:363-371) are checked against INDEPENDENT
formulations of the same definitions instead, so a mistake in the restatement does not silently
become the truth the CUDA kernels are compared with."""
import torch

from oracle import quant_oracle as O


def _case(N, K, seed, n=6, rows=64):
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(N, K, generator=g) * 0.05
    chan = torch.ones(K)
    hot = torch.randperm(K, generator=g)[: max(1, K // 50)]
    chan[hot] = 12.0
    X = torch.cat([torch.randn(rows, K, generator=g) * chan for _ in range(n)])
    return W, X, hot


def test_awq_search_loss_is_the_output_error_on_the_calibration_rows():
    """tr(dW H dW^T) with H = X^T X must equal ||dW X^T||_F^2 computed row by row from the
    activations, with dW built from the PINNED awq_layer restatement (a8) for each candidate."""
    W, X, hot = _case(48, 256, 1)
    H = X.T @ X
    cands = [1.0, 1.25, 1.5, 2.0]
    got = O.awq_search_losses(W, H, hot, 4, 128, cands)
    # importance vector that makes exactly `hot` the top-k, so awq_layer protects the same channels
    imp = torch.zeros(256)
    imp[hot] = 1.0
    for c, sf in enumerate(cands):
        q = O.awq_layer(W, [imp], 4, 128, len(hot) / 256 + 1e-9, sf)["out"]
        direct = (((q - W).double() @ X.double().T) ** 2).sum()
        assert abs(got[c].item() - direct.item()) <= 2e-6 * direct.item()     # H itself is fp32


def test_compensated_gptq_blocking_is_only_a_schedule():
    """Algorithm 1 with the lazy block update must give the same codes for any block size (the
    blocks only reorder when the error of a column reaches the later columns), and must beat
    round-to-nearest on the objective it minimises."""
    W, X, _ = _case(32, 256, 2)
    H = O.gptq_hessian([X], 256, torch.float32, 128, 0.01)
    q128 = O.gptq_compensated(W, H, 4, 128, blocksize=128)
    q32 = O.gptq_compensated(W, H, 4, 128, blocksize=32)
    q256 = O.gptq_compensated(W, H, 4, 128, blocksize=256)
    assert (q128 == q32).float().mean() > 0.999 and (q128 == q256).float().mean() > 0.999
    rtn = O.uniform_group_quant(W, 4, 128)["out"]

    def objective(q):
        d = (q - W).double()
        return ((d @ H.double()) * d).sum().item()
    assert objective(q128) < objective(rtn)
    # act-order is a column permutation applied before and undone after
    perm = torch.argsort(torch.diag(H), descending=True)
    qp = O.gptq_compensated(W, H, 4, 128, perm=perm)
    manual = O.gptq_compensated(W[:, perm], H[perm][:, perm], 4, 128)[:, torch.argsort(perm)]
    assert torch.equal(qp, manual)


def test_smooth_alpha_measure_is_built_from_pinned_pieces():
    """The alpha-sweep measure must equal: smooth with the PINNED smooth_layer restatement (a15),
    quantize with the PINNED uniform quantizer (a7), scale back, weight by the activation scale."""
    W, X, _ = _case(40, 256, 3)
    act = X.abs().amax(0)
    alphas = [0.0, 0.3, 0.5, 0.85, 1.0]
    S = torch.stack([O.smooth_scale(act, W, a).float() for a in alphas])
    got = O.smooth_alpha_errors(W, S, act, 8, -1)
    for i, a in enumerate(alphas):
        sm = O.smooth_layer(W, act, a)
        q = O.uniform_group_quant(sm["out"], 8, -1)["out"]
        err = ((q * sm["s"].to(W.dtype)).double() - W.double()) * act.double()
        want = (err ** 2).sum().item()
        assert abs(got[i].item() - want) <= 1e-9 * want
    # alpha = 0 leaves s = 1 / wmax^1 ... the measure must differ across alphas (not degenerate)
    assert len({round(v, 6) for v in got.tolist()}) == len(alphas)


def test_pack_layout_is_a_plain_little_endian_bit_stream():
    import numpy as np
    rng = np.random.default_rng(1)
    for b in (2, 3, 4, 5, 8):
        codes = rng.integers(0, 1 << b, size=(2, 50), dtype=np.uint8)
        words = O.pack_codes(codes, b)
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, : 50 * b]
        want = ((codes[:, :, None] >> np.arange(b)) & 1).reshape(2, -1)
        assert np.array_equal(bits, want)
