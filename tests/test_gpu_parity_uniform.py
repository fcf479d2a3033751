"""CUDA path (through the C ABI) vs golden vectors and vs the oracle: GPTQ parity column stage,
symmetric / asymmetric group fake-quant, AWQ, SmoothQuant, activation statistics.
Bar: bit-exact outputs, integer codes, scales and zero points (fp32, fp16, bf16)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import case_dtype
from oracle import quant_oracle as O

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda()


def same(got: torch.Tensor, want: torch.Tensor, what=""):
    got = got.cpu()
    assert got.dtype == want.dtype and got.shape == want.shape, what
    if not torch.equal(got, want):
        bad = (got.float() != want.float())
        raise AssertionError(f"{what}: {bad.sum().item()} of {bad.numel()} elements differ")


# ---------------------------------------------------------------------------------------------
def test_pseudo_quantize_golden(golden):
    from quantization_utils import pseudo_quantize_tensor
    from b200q import ops
    g = golden("uniform")
    for case in g.cases("uniform"):
        dt = torch.float32 if case.startswith("const") else case_dtype(case)
        b, G = (int(v) for v in g.arr(f"uniform/{case}/meta"))
        w = g.tensor(f"uniform/{case}/w", dt)
        want = g.tensor(f"uniform/{case}/out", dt)
        same(pseudo_quantize_tensor(dev(w), n_bit=b, q_group_size=G), want, case)
        same(pseudo_quantize_tensor(w, n_bit=b, q_group_size=G), want, case + " (host tensor in)")
        if g.has(f"uniform/{case}/codes"):
            out, codes, scales, zeros = ops.group_fakequant(dev(w), b, G, return_codes=True)
            assert np.array_equal(codes.cpu().numpy().astype(np.int16), g.arr(f"uniform/{case}/codes"))
            assert np.array_equal(scales.cpu().numpy(), g.arr(f"uniform/{case}/scales"))
            assert np.array_equal(zeros.cpu().numpy(), g.arr(f"uniform/{case}/zeros"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,b,G", [((512, 1024), 4, 128), ((300, 384), 3, 128), ((64, 768), 8, 64),
                                       ((33, 200), 4, -1), ((16, 4096), 4, -1), ((7, 96), 2, 32)])
def test_pseudo_quantize_vs_oracle(dtype, shape, b, G):
    from b200q import ops
    g = torch.Generator().manual_seed(hash((shape, b, G)) % 2**31)
    w = (torch.randn(*shape, generator=g) * 0.02).to(dtype)
    w[0, :5] = 0
    r = O.uniform_group_quant(w, b, G)
    out, codes, scales, zeros = ops.group_fakequant(dev(w), b, G, return_codes=True)
    same(out, r["out"])
    assert torch.equal(codes.cpu().to(torch.int32), r["codes"])
    assert torch.equal(scales.cpu(), r["scales"]) and torch.equal(zeros.cpu(), r["zeros"])


def test_simple_quantize_layer_golden(golden):
    from gptq_quantizer import _simple_quantize_layer
    g = golden("simple")
    for case in g.cases("simple"):
        dt = case_dtype(case)
        b, G = (int(v) for v in g.arr(f"simple/{case}/meta"))
        w = g.tensor(f"simple/{case}/w", dt)
        lin = nn.Linear(w.shape[1], w.shape[0], bias=False)
        lin.weight.data = dev(w)
        _simple_quantize_layer(lin, b, G)
        assert lin.weight.data.is_cuda
        same(lin.weight.data, g.tensor(f"simple/{case}/out", dt), case)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_symmetric_vs_oracle(dtype):
    from b200q import ops
    g = torch.Generator().manual_seed(11)
    for shape, b, G in (((256, 512), 4, 128), ((40, 320), 3, 64), ((9, 100), 5, -1)):
        w = (torch.randn(*shape, generator=g) * 0.02).to(dtype)
        r = O.symmetric_group_quant(w, b, G)
        out, codes, scales, _ = ops.group_fakequant(dev(w), b, G, symmetric=True, return_codes=True)
        same(out, r["out"])
        assert torch.equal(codes.cpu().to(torch.int32), r["codes"])
        assert torch.equal(scales.cpu(), r["scales"])


def test_gptq_layer_golden(golden):
    import gptq_quantizer as gq
    from b200q import ops
    g = golden("gptq")
    for case in g.cases("gptq"):
        dt = case_dtype(case)
        b, ns, act = (int(v) for v in g.arr(f"gptq/{case}/meta"))
        w = g.tensor(f"gptq/{case}/w", dt)
        feats = [f.to(dt) for f in g.tensor(f"gptq/{case}/feats")]
        lin = nn.Linear(w.shape[1], w.shape[0], bias=False)
        lin.weight.data = dev(w)
        gq._gptq_quantize_layer(lin, b, 128, feats, perp_damp=0.01, blocksize=32, nsamples=ns,
                                actorder=bool(act), verbose=False)
        same(lin.weight.data, g.tensor(f"gptq/{case}/out", dt), case)
        if b <= 7:
            out, codes, scales = ops.gptq_parity_quant(dev(w), b, return_codes=True)
            assert np.array_equal(codes.cpu().numpy().astype(np.int16), g.arr(f"gptq/{case}/codes"))
            assert np.array_equal(scales.cpu().numpy(), g.arr(f"gptq/{case}/scales"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K,b", [(1024, 768, 4), (333, 1000, 3), (77, 130, 4), (5, 8, 2)])
def test_gptq_parity_vs_oracle(dtype, N, K, b):
    from b200q import ops
    g = torch.Generator().manual_seed(N * K)
    w = (torch.randn(N, K, generator=g) * 0.02).to(dtype)
    w[:, 0] = 0     # an all-zero column hits the 1e-5 clamp
    r = O.gptq_parity_quant(w, b)
    out, codes, scales = ops.gptq_parity_quant(dev(w), b, return_codes=True)
    same(out, r["out"])
    assert torch.equal(codes.cpu().to(torch.int32), r["codes"])
    assert torch.equal(scales.cpu(), r["scales"])


def test_col_absmax_strided_and_accumulate():
    from b200q import ops
    g = torch.Generator().manual_seed(3)
    big = torch.randn(300, 1024, generator=g).cuda()
    view = big[:, 128:640]                      # row stride 1024, 512 columns
    want = view.abs().amax(0)
    assert torch.equal(ops.col_absmax(view), want)
    acc = ops.col_absmax(view[:100].contiguous())
    ops.col_absmax(view[100:].contiguous(), out=acc, accumulate=True)
    assert torch.equal(acc, want)


# ---------------------------------------------------------------------------------------------
class TinyNet(nn.Module):
    def __init__(self, g, case, dt):
        super().__init__()
        self.fc1 = nn.Linear(256, 48, bias=False)
        self.fc2 = nn.Linear(128, 32, bias=True)
        self.head = nn.Linear(256, 8, bias=False)
        for n in ("fc1", "fc2", "head"):
            getattr(self, n).weight.data = g.tensor(f"{case}/{n}/w", dt)


def test_awq_walker_golden(golden):
    from awq_quantizer import awq_quantize_model_weight
    g = golden("walkers")
    for case in ("f32_sf2", "f32_sf1p5", "f16_sf2", "f32_sf2_b8"):
        dt = case_dtype(case)
        b, G, sf = g.arr(f"awq/{case}/meta")
        net = TinyNet(g, f"awq/{case}", dt).cuda()
        feats = {n: list(g.tensor(f"awq/{case}/{n}/feats")) for n in ("fc1", "fc2")}
        awq_quantize_model_weight(net, w_bit=int(b), q_group_size=int(G), input_feat=feats,
                                  protect_ratio=0.01, scale_factor=float(sf))
        for n in ("fc1", "fc2", "head"):
            same(getattr(net, n).weight.data, g.tensor(f"awq/{case}/{n}/out", dt), f"{case}/{n}")


def test_awq_layer_vs_oracle_larger():
    from awq_quantizer import awq_quantize_model_weight
    g = torch.Generator().manual_seed(21)
    K, N = 1024, 640
    chan = torch.ones(K)
    chan[torch.randperm(K, generator=g)[:10]] = 20.0
    feats = [(torch.randn(64, K, generator=g) * chan).abs().mean(0) for _ in range(32)]
    for dtype in (torch.float32, torch.float16):
        w = (torch.randn(N, K, generator=g) * 0.02).to(dtype)
        net = nn.Sequential(nn.Linear(K, N, bias=False))
        net[0].weight.data = w.clone().cuda()
        awq_quantize_model_weight(net, 4, 128, {"0": feats}, 0.01, 2.0)
        r = O.awq_layer(w, feats, 4, 128, 0.01, 2.0)
        same(net[0].weight.data, r["out"])


def test_gptq_walker_golden(golden):
    from gptq_quantizer import gptq_quantize_model_weight
    g = golden("walkers")
    for case in ("f32_b4", "f32_b3_act"):
        b, G, act = (int(v) for v in g.arr(f"gptqwalk/{case}/meta"))
        net = TinyNet(g, f"gptqwalk/{case}", torch.float32).cuda()
        feats = {n: list(g.tensor(f"gptqwalk/{case}/{n}/feats")) for n in ("fc1", "fc2")}
        gptq_quantize_model_weight(net, w_bit=b, q_group_size=G, input_feat=feats, nsamples=8,
                                   actorder=bool(act), verbose=False)
        for n in ("fc1", "fc2", "head"):
            same(getattr(net, n).weight.data, g.tensor(f"gptqwalk/{case}/{n}/out"), f"{case}/{n}")


def test_smoothquant_walker_golden(golden):
    import smooth_quant_quantizer as sq
    g = golden("walkers")
    for case in ("f32_a0p5", "f32_a0p85", "f32_a0", "f32_a1"):
        b, G, alpha = g.arr(f"smooth/{case}/meta")
        net = TinyNet(g, f"smooth/{case}", torch.float32).cuda()
        act = {n: g.tensor(f"smooth/{case}/{n}/act") for n in ("fc1", "fc2")}
        w0 = {n: getattr(net, n).weight.data.clone().cpu() for n in ("fc1", "fc2", "head")}
        sq.smoothquant_quantize_model_weight(net, int(b), int(G), act, alpha=float(alpha), verbose=False)
        for n in ("fc1", "fc2", "head"):
            m = getattr(net, n)
            want = g.tensor(f"smooth/{case}/{n}/out")
            if n in act:
                s = m.smoothing_scale.cpu()
                # s goes through pow(): torch's CPU pow(x, 0.5) is MKL VML's sqrt (<= 1 ulp, not
                # correctly rounded) and a general exponent is SLEEF's powf, so s is compared to
                # 2 ulp; alpha in {0, 1} involves no transcendental and must be exact
                if float(alpha) in (0.0, 1.0):
                    assert torch.equal(s, g.tensor(f"smooth/{case}/{n}/s")), f"{case}/{n}/s"
                    same(m.weight.data, want, f"{case}/{n}")
                torch.testing.assert_close(s, g.tensor(f"smooth/{case}/{n}/s"), rtol=3e-7, atol=0)
                # given OUR s, everything downstream must be bit-exact
                r = O.smoothquant_layer(w0[n], None, float(alpha), int(b), int(G), s=s)
                same(m.weight.data, r["out"], f"{case}/{n} (own s)")
                assert m._smooth_pre_hook_handle is not None
            else:
                same(m.weight.data, want, f"{case}/{n}")


def test_smooth_weights_and_reverse(golden):
    import smooth_quant_quantizer as sq
    g = golden("walkers")
    w = g.tensor("smoothw/f32_a0p5/w")
    net = nn.Sequential()
    net.add_module("fc1", nn.Linear(256, 48, bias=False))
    net.fc1.weight.data = w.clone().cuda()
    x = torch.randn(4, 256, device="cuda")
    y0 = net(x)
    sq.smooth_weights(net, {"fc1": g.tensor("smoothw/f32_a0p5/act")}, alpha=0.5, verbose=False)
    s = net.fc1.smoothing_scale.cpu()
    torch.testing.assert_close(s, g.tensor("smoothw/f32_a0p5/s"), rtol=3e-7, atol=0)
    same(net.fc1.weight.data, w / s)            # W / s is an exact IEEE division given s
    torch.testing.assert_close(net(x), y0, rtol=1e-4, atol=1e-5)     # the pre-hook keeps y = W x
    sq.reverse_weight_smoothing(net, verbose=False)
    torch.testing.assert_close(net.fc1.weight.data.cpu(), w, rtol=1e-6, atol=0)
    assert not hasattr(net.fc1, "smoothing_scale")


def test_act_stats(golden):
    from b200q import ops
    g = golden("act")
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        x = g.tensor(f"act/{name}/x", dt).cuda()
        assert torch.equal(ops.act_maxabs(x).cpu(), g.tensor(f"act/{name}/maxabs"))
        # mean|x|: fp32 summation order differs from ATen's cascade -> 1e-6 relative
        torch.testing.assert_close(ops.act_meanabs(x).cpu(), g.tensor(f"act/{name}/meanabs"),
                                   rtol=2e-6 if dt == torch.float32 else 1e-3, atol=0)
    x = torch.randn(3, 1000, 512, device="cuda")
    run = ops.act_maxabs(x[0])
    ops.act_maxabs(x[1], out=run)
    ops.act_maxabs(x[2], out=run)
    assert torch.equal(run, x.reshape(-1, 512).abs().amax(0))


def test_get_calib_feat_and_collect_act_scales_hooks():
    import quantization_utils as qu
    import smooth_quant_quantizer as sq
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 32)).cuda()
    samples = [torch.randn(2, 10, 64) for _ in range(3)]
    feats = qu.get_calib_feat(net, None, samples, verbose=False)
    scales = sq.collect_act_scales(net, samples, verbose=False)
    assert sorted(feats) == ["0", "2"] and len(feats["0"]) == 3 and not feats["0"][0].is_cuda
    x0 = torch.cat([s.reshape(-1, 64) for s in samples])
    torch.testing.assert_close(feats["0"][1], samples[1].reshape(-1, 64).abs().mean(0), rtol=1e-5, atol=1e-7)
    assert torch.equal(scales["0"], x0.abs().amax(0))
    assert scales["2"].shape == (128,)
    # on-device capture: same numbers, no host copies, and the AWQ walker takes
    # the [n_batches, C] matrices where it takes the reference's lists
    import awq_quantizer as aq
    qu.CALIB_ON_DEVICE = True
    try:
        dev_feats = qu.get_calib_feat(net, None, samples, verbose=False)
    finally:
        qu.CALIB_ON_DEVICE = False
    assert dev_feats["0"].is_cuda and dev_feats["0"].shape == (3, 64)
    assert torch.equal(dev_feats["0"].cpu(), torch.stack(feats["0"]))
    a = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 32)).cuda()
    b = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 32)).cuda()
    b.load_state_dict(a.state_dict())
    aq.awq_quantize_model_weight(a, 4, 32, feats, 0.05, 2.0)
    aq.awq_quantize_model_weight(b, 4, 32, dev_feats, 0.05, 2.0)
    assert torch.equal(a[0].weight.data, b[0].weight.data) and torch.equal(a[2].weight.data, b[2].weight.data)


def test_importance_sum_is_pythons_left_to_right():
    from awq_quantizer import _importance
    g = torch.Generator().manual_seed(5)
    feats = [torch.rand(777, generator=g) for _ in range(128)]
    assert torch.equal(_importance(feats, "cuda").cpu(), sum(feats).float())
    feats16 = [f.half() for f in feats]
    assert torch.equal(_importance(feats16, "cuda").cpu(), sum(feats16).float())


@pytest.mark.parametrize("K,k", [(4096, 40), (11008, 110), (768, 7), (130, 1), (28672, 286), (64, 64)])
def test_topk_mask_matches_torch_topk(K, k):
    """The radix-select top-k picks the same channel SET as torch.topk (awq_quantizer.py:61)."""
    from b200q import ops
    g = torch.Generator().manual_seed(K + k)
    w = torch.randn(256, K, generator=g) * 0.02
    chan = torch.ones(K)
    chan[torch.randperm(K, generator=g)[: max(1, K // 100)]] = 20.0
    feats = torch.stack([(torch.randn(16, K, generator=g) * chan).abs().mean(0) for _ in range(8)])
    out, mask = ops.awq_layer(w.cuda(), feats.cuda(), 4, -1 if K % 128 else 128, k, 2.0,
                              return_mask=True)
    want = torch.zeros(K, dtype=torch.uint8)
    want[torch.topk(sum(feats).float(), k)[1]] = 1
    assert torch.equal(mask.cpu(), want)
    assert int(mask.sum()) == k


def test_topk_ties_take_lowest_indices():
    from b200q import ops
    feats = torch.ones(1, 256)
    feats[0, 100] = 5.0
    w = torch.randn(8, 256)
    _, mask = ops.awq_layer(w.cuda(), feats.cuda(), 4, 128, 4, 2.0, return_mask=True)
    assert mask.cpu().nonzero().flatten().tolist() == [0, 1, 2, 100]


def test_reused_divisor_division_is_ieee_exact():
    """reciprocal + two Markstein corrections == __fdiv_rn on 2^31 random / adversarial pairs."""
    import ctypes
    from b200q import _lib
    bad = ctypes.c_int64(-1)
    rc = _lib.load().b200q_selftest_div(1 << 31, 12345, ctypes.byref(bad),
                                        torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert bad.value == 0, _lib.load().b200q_last_error().decode()


def test_profile_records_entry_points():
    from b200q import _lib, ops
    w = torch.randn(256, 1024, device="cuda")
    _lib.profile_enable(True)
    for _ in range(3):
        ops.group_fakequant(w, 4, 128)
    ops.col_absmax(w)
    _lib.profile_enable(False)
    q = _lib.profile_query("group_fakequant")
    assert q["launches"] == 3 and q["ms"] > 0 and q["bytes"] == 3 * 2 * w.numel() * 4
    assert _lib.profile_query("col_absmax")["launches"] == 1
    assert _lib.profile_query(None)["launches"] == 4


def test_shape_errors_are_assertions():
    from quantization_utils import pseudo_quantize_tensor
    with pytest.raises(AssertionError):
        pseudo_quantize_tensor(torch.randn(4, 100), n_bit=4, q_group_size=32)
    with pytest.raises(AssertionError):
        pseudo_quantize_tensor(torch.randn(2, 4, 8), n_bit=4, q_group_size=-1)


def test_gptq_boundary_cases_the_reference_accepts():
    """in_features not a multiple of 8 (H cannot be built by the TMA-fed kernel: parity output does
    not need it), and a symmetric-fallback group size that divides the element count but not
    in_features (the reference's reshape(-1, G), gptq_quantizer.py:88-91)."""
    import warnings
    import torch.nn as nn
    import gptq_quantizer as gq
    from oracle import quant_oracle as O
    g = torch.Generator().manual_seed(11)
    w = torch.randn(48, 100, generator=g) * 0.02
    lin = nn.Linear(100, 48, bias=False)
    lin.weight.data = w.clone().cuda()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        gq._gptq_quantize_layer(lin, 4, 128, [torch.randn(16, 100, generator=g) for _ in range(3)],
                                verbose=False)
    assert any("not a multiple of 8" in str(x.message) for x in rec)
    assert torch.equal(lin.weight.data.cpu(), O.gptq_parity_quant(w, 4)["out"])
    # 48 x 100 = 4800 elements in groups of 64: groups straddle rows
    lin.weight.data = w.clone().cuda()
    gq._simple_quantize_layer(lin, 4, 64)
    assert torch.equal(lin.weight.data.cpu(), O.symmetric_group_quant(w, 4, 64)["out"])


def test_pseudo_quantize_rejects_nan_like_the_reference():
    from quantization_utils import pseudo_quantize_tensor
    w = torch.randn(8, 128)
    w[3, 5] = float("nan")
    with pytest.raises(AssertionError):
        pseudo_quantize_tensor(w.cuda(), 4, 128)
