"""The assertions of the reference's own smoke suite (test_quantization.py: shape kept, finite
output, weights change, error shrinks with more bits, extreme inputs stay finite, SmoothQuant on a
tiny MLP), restated against the drop-in modules with HOST tensors exactly as that suite passes them.
The arithmetic still runs on the GPU (host tensors are streamed through it)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def mse(a, b):
    return ((a - b) ** 2).mean().item()


def test_tensor_level_quantizers_on_host_tensors():
    from quantization_utils import pseudo_quantize_tensor
    from pot_apot_quantizer import pot_quantize_tensor, apot_quantize_tensor
    torch.manual_seed(0)
    w = torch.randn(32, 64)
    for fn in (pseudo_quantize_tensor, pot_quantize_tensor, apot_quantize_tensor):
        for G in (-1, 32):
            q = fn(w, n_bit=4, q_group_size=G)
            assert q.shape == w.shape and q.dtype == w.dtype and not q.is_cuda
            assert torch.isfinite(q).all() and not torch.equal(q, w)
    # more bits -> smaller error (test_quantization.py:168-186)
    errs = [mse(pseudo_quantize_tensor(w, n_bit=b, q_group_size=32), w) for b in (2, 4, 8)]
    assert errs[0] > errs[1] > errs[2]


@pytest.mark.parametrize("scale", [1000.0, 1e-3])
def test_extreme_magnitudes_stay_finite(scale):
    from quantization_utils import pseudo_quantize_tensor
    from pot_apot_quantizer import pot_quantize_tensor, apot_quantize_tensor
    torch.manual_seed(1)
    w = torch.randn(16, 64) * scale
    for fn in (pseudo_quantize_tensor, pot_quantize_tensor, apot_quantize_tensor):
        assert torch.isfinite(fn(w, n_bit=4, q_group_size=32)).all()
    for const in (1.0, -1.0):
        c = torch.full((8, 64), const)
        for fn in (pseudo_quantize_tensor, pot_quantize_tensor, apot_quantize_tensor):
            assert torch.isfinite(fn(c, n_bit=4, q_group_size=32)).all()


def test_smoothquant_pipeline_on_a_host_mlp():
    import smooth_quant_quantizer as sq
    torch.manual_seed(2)
    net = nn.Sequential(nn.Linear(10, 20), nn.ReLU(), nn.Linear(20, 5))
    samples = [torch.randn(4, 10) for _ in range(3)]
    scales = sq.collect_act_scales(net, samples, verbose=False)
    assert set(scales) == {"0", "2"} and scales["0"].shape == (10,) and (scales["0"] > 0).all()
    norms = []
    for alpha in (0.0, 0.25, 0.5, 0.75, 1.0):
        m = nn.Sequential(nn.Linear(10, 20), nn.ReLU(), nn.Linear(20, 5))
        m.load_state_dict(net.state_dict())
        sq.smooth_weights(m, scales, alpha=alpha, verbose=False)
        assert hasattr(m[0], "smoothing_scale") and m[0].smoothing_scale.shape == (10,)
        norms.append(m[0].weight.data.norm().item())
    assert len({round(n, 5) for n in norms}) == 5          # different alphas, different weights
    m = nn.Sequential(nn.Linear(10, 20), nn.ReLU(), nn.Linear(20, 5))
    m.load_state_dict(net.state_dict())
    x = torch.randn(6, 10)
    y0 = m(x)
    out = sq.smoothquant_quantize_and_calibrate(m, 8, -1, samples, alpha=0.5, verbose=False)
    assert set(out) == {"0", "2"}
    assert torch.isfinite(m[0].weight.data).all() and not m[0].weight.data.is_cuda
    # 8-bit weights + the activation hook keep the function close to the original
    assert (m(x) - y0).abs().max() < 0.05
    assert sq.smoothquant_search_alpha(m, samples, scales, w_bit=8, q_group_size=-1, verbose=False) >= 0.0


def test_host_model_walkers_write_back_to_the_host():
    from awq_quantizer import awq_quantize_model_weight
    from gptq_quantizer import gptq_quantize_model_weight
    from oracle import quant_oracle as O
    torch.manual_seed(3)
    net = nn.Sequential(nn.Linear(256, 64, bias=False), nn.Linear(128, 32, bias=False))
    w0 = [net[0].weight.data.clone(), net[1].weight.data.clone()]
    feats = {"0": [torch.rand(256) for _ in range(4)], "1": [torch.rand(128) for _ in range(4)]}
    awq_quantize_model_weight(net, 4, 128, feats, 0.01, 2.0)
    for i, n in enumerate(("0", "1")):
        assert not net[i].weight.data.is_cuda
        assert torch.equal(net[i].weight.data, O.awq_layer(w0[i], feats[n], 4, 128, 0.01, 2.0)["out"])
    net[0].weight.data, net[1].weight.data = w0[0].clone(), w0[1].clone()
    gptq_quantize_model_weight(net, 4, 128, {"0": feats["0"]}, verbose=False)
    assert torch.equal(net[0].weight.data, O.gptq_parity_quant(w0[0], 4)["out"])
    assert torch.equal(net[1].weight.data, O.symmetric_group_quant(w0[1], 4, 128)["out"])
