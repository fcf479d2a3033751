"""SmoothQuant alpha sweep: the fused error kernel against the oracle's restatement of the measure
(smooth_quant_quantizer.py:327-371 is a stub in the reference: PARITY UNPINNED)."""
import pytest
import torch
import torch.nn as nn

from oracle import quant_oracle as O

pytestmark = pytest.mark.gpu


def case(N, K, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dtype)
    act = torch.rand(K, generator=g) * 4 + 0.05
    act[torch.randperm(K, generator=g)[: K // 50 + 1]] *= 30.0
    return W, act


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K,b,G", [(64, 512, 8, -1), (200, 384, 4, 128), (33, 1024, 8, 256)])
def test_errors_match_oracle(dtype, N, K, b, G):
    from b200q import ops
    W, act = case(N, K, N + K, dtype)
    alphas = torch.linspace(0.0, 1.0, 9).tolist()
    # scales as values of the weight's dtype (the drop-in promotes W instead when they are wider)
    S = torch.stack([O.smooth_scale(act, W, a).float() for a in alphas]).to(dtype).float()
    got = ops.smooth_alpha_errors(W.cuda(), S.cuda(), act.cuda(), b, G).cpu()
    want = O.smooth_alpha_errors(W, S, act, b, G)
    # fp32 per-element errors summed in fp32 per group, fp64 across groups
    assert ((got - want).abs() / want).max().item() < (1e-4 if dtype == torch.float32 else 2e-3)
    assert int(torch.argmin(got)) == int(torch.argmin(want))
    # accumulation across calls (a model walk adds layer after layer on the device)
    twice = ops.smooth_alpha_errors(W.cuda(), S.cuda(), act.cuda(), b, G, got.cuda().clone()).cpu()
    assert torch.allclose(twice, 2 * got, rtol=1e-12)


def test_search_entry_point_returns_the_argmin(capsys):
    import smooth_quant_quantizer as sq
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(256, 128, bias=False), nn.Linear(128, 256, bias=False)).cuda()
    Ws = [m.weight.data.clone() for m in net]
    acts = {"0": case(1, 256, 1)[1], "1": case(1, 128, 2)[1]}
    best = sq.smoothquant_search_alpha(net, [], acts, w_bit=8, q_group_size=-1, n_grid=11, verbose=False)
    alphas = torch.linspace(0.0, 1.0, 11, dtype=torch.float64).tolist()
    total = torch.zeros(11, dtype=torch.float64)
    for (name, act), W in zip(acts.items(), Ws):
        S = torch.stack([O.smooth_scale(act, W.cpu(), a).float() for a in alphas])
        total += O.smooth_alpha_errors(W.cpu(), S, act.clamp(min=1e-5), 8, -1)
    assert best == alphas[int(torch.argmin(total))]
    for m, W in zip(net, Ws):
        assert torch.equal(m.weight.data, W), "the search must not modify the model"
