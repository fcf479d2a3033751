"""Streaming calibration capture (b200q.streaming): folding batches into the
running Hessian / Gram matrix as the hook sees them must give what the reference's list-of-batches
layout gives (same kernels, different fp32 summation order: 1e-5)."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def batches(seed, n, rows, K, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    chan = torch.ones(K)
    chan[torch.randperm(K, generator=g)[: max(1, K // 64)]] = 15.0
    return [(torch.randn(rows, K, generator=g) * chan).to(dtype) for _ in range(n)]


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.parametrize("nsamples", [128, 4])
def test_streamed_hessian_equals_list_hessian(nsamples):
    from b200q import tensor_ops as T
    from b200q.streaming import ActivationStream
    K = 384
    feats = batches(3, 6, 96, K)
    want = T.gptq_hessian(feats, K, "cuda", 0.01, nsamples)
    s = ActivationStream(K, normalize=True, max_batches=nsamples, keep_stats=False)
    for f in feats:
        s.add(f)
    got = T.gptq_hessian(s, K, "cuda", 0.01, nsamples)
    assert rel(got, want) < 1e-5
    assert s.batches_seen == 6


def test_calibration_hook_streams_and_the_walker_accepts_streams():
    import gptq_quantizer as gq
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(256, 384, bias=False), nn.Linear(384, 256, bias=False)).cuda()
    samples = [torch.randn(4, 32, 256) for _ in range(5)]
    lists = gq.gptq_calibrate_hessian(net, samples, nsamples=128, verbose=False)
    streams = gq.gptq_calibrate_hessian_streaming(net, samples, nsamples=128, verbose=False)
    assert set(lists) == set(streams) == {"0", "1"}
    assert streams["0"].batches_seen == 5 and streams["0"].stat_rows == []
    from b200q import tensor_ops as T
    for mode in ("parity", "compensated"):
        gq.MODE = mode
        try:
            a, b = copy.deepcopy(net), copy.deepcopy(net)
            gq.gptq_quantize_model_weight(a, 4, 128, lists, verbose=False)
            gq.gptq_quantize_model_weight(b, 4, 128, streams, verbose=False)
        finally:
            gq.MODE = "parity"
        for name, (la, lb, l0) in zip(("0", "1"), zip(a, b, net)):
            same = (la.weight == lb.weight).float().mean().item()
            if mode == "parity":
                assert same == 1.0            # the reference's column stage does not depend on H
                continue
            # The compensated loop amplifies last-bit differences of H: one code that rounds the
            # other way shifts the rest of its row.  Both results must be equally good GPTQ
            # solutions: most codes equal, and the objective tr(dW H dW^T) within 2 %.
            H = T.gptq_hessian(lists[name], l0.weight.shape[1], "cuda").double()
            def objective(q):
                d = (q.weight - l0.weight).double()
                return torch.einsum("ik,kl,il->", d, H, d).item()
            assert same >= 0.95, same
            assert abs(objective(la) - objective(lb)) <= 0.02 * objective(la)


def test_awq_accepts_streams():
    import awq_quantizer as aq
    from b200q.streaming import ActivationStream
    K, N = 512, 256
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(K, N, bias=False)).cuda()
    feats = batches(9, 8, 128, K)
    s = ActivationStream(K, normalize=False)
    for f in feats:
        s.add(f)
    stats = [f.abs().mean(0) for f in feats]
    best_list = aq.awq_search_scale_factor(net, 4, 128, {"0": feats}, n_grid=10)
    best_stream = aq.awq_search_scale_factor(net, 4, 128, {"0": s}, n_grid=10)
    assert best_list == best_stream
    a, b = copy.deepcopy(net), copy.deepcopy(net)
    aq.awq_quantize_model_weight(a, 4, 128, {"0": [st.cuda() for st in stats]}, 0.01, 1.5)
    aq.awq_quantize_model_weight(b, 4, 128, {"0": s}, 0.01, 1.5)
    # the kernel's mean|x| differs from torch's in the last bits (different summation order), which
    # can only matter if two channels tie for the last protected slot
    assert (a[0].weight == b[0].weight).float().mean().item() >= 0.999
