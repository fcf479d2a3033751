"""CUDA POT / APOT (through the C ABI) vs golden vectors and vs the oracle.
Bar (BASELINE.json north_star): level selection bit-exact — exponent codes, level indices, the
chosen grid point and scale per group, and the dequantized values."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import case_dtype
from oracle import quant_oracle as O

pytestmark = pytest.mark.gpu


def same(got, want, what=""):
    got = got.cpu()
    assert got.dtype == want.dtype and got.shape == want.shape, what
    if not torch.equal(got, want):
        bad = got != want
        raise AssertionError(f"{what}: {bad.sum().item()} of {bad.numel()} elements differ")


def test_pot_golden(golden):
    from pot_apot_quantizer import pot_quantize_tensor
    from b200q import ops
    g = golden("pot")
    grid = torch.arange(0.01, 2.01, 0.01)
    assert np.array_equal(grid.numpy(), g.arr("pot/grid"))
    for case in g.cases("pot"):
        b, G = (int(v) for v in g.arr(f"pot/{case}/meta"))
        w = g.tensor(f"pot/{case}/w", case_dtype(case))
        want = g.tensor(f"pot/{case}/out", case_dtype(case))
        same(pot_quantize_tensor(w.cuda(), n_bit=b, q_group_size=G), want, case)
        same(pot_quantize_tensor(w, n_bit=b, q_group_size=G), want, case + " (host tensor in)")
        groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
        out, exps, scale, idx = ops.pot_quant(groups, b, grid, return_codes=True)
        assert np.array_equal(exps.cpu().numpy().reshape(w.shape), g.arr(f"pot/{case}/exps")), case
        assert np.array_equal(idx.cpu().numpy(), g.arr(f"pot/{case}/best_idx")), case
        assert np.array_equal(scale.cpu().numpy(), g.arr(f"pot/{case}/scale")), case


@pytest.mark.parametrize("shape,b,G,mul", [
    ((256, 1024), 4, 128, 1.0), ((64, 512), 3, 128, 30.0), ((32, 256), 8, 128, 1e-3),
    ((16, 768), 4, -1, 1.0), ((24, 320), 4, 64, 1.0), ((10, 56), 2, 7, 1.0), ((6, 4096), 4, -1, 1.0),
])
def test_pot_vs_oracle(shape, b, G, mul):
    from b200q import ops
    g = torch.Generator().manual_seed(abs(hash((shape, b, G))) % 2**31)
    w = torch.randn(*shape, generator=g) * 0.02 * mul
    r = O.pot_quant(w, b, G)
    groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
    out, exps, scale, idx = ops.pot_quant(groups, b, O.pot_grid(), return_codes=True)
    assert torch.equal(idx.cpu(), r["best_idx"]), "chosen grid point differs"
    assert torch.equal(exps.cpu().reshape(shape).to(torch.int32), r["exps"])
    assert torch.equal(scale.cpu(), r["scale"])
    same(out.reshape(shape), r["out"])


def test_pot_power_of_two_neighbourhoods():
    """Group maxima a few ulps around powers of two, and ratios around sqrt(2)*2^n: the places
    where log2f's last bit decides floor() / round()."""
    from b200q import ops
    rows = []
    for e in (-9, -6, -5, -1, 0, 3):
        for d in range(-4, 3):
            top = np.array([2.0 ** e], np.float32).view(np.int32) + d
            g = torch.Generator().manual_seed(1000 + e * 10 + d)
            row = torch.rand(128, generator=g) * float(top.view(np.float32)[0]) * 0.999
            row[17] = float(top.view(np.float32)[0])
            rows.append(row * torch.where(torch.rand(128, generator=g) < 0.5, -1.0, 1.0))
    w = torch.stack(rows)
    r = O.pot_quant(w, 4, 128)
    out, exps, scale, idx = ops.pot_quant(w.cuda(), 4, O.pot_grid(), return_codes=True)
    assert torch.equal(scale.cpu(), r["scale"])
    assert torch.equal(exps.cpu().to(torch.int32), r["exps"])
    same(out, r["out"])


def test_apot_golden(golden):
    from pot_apot_quantizer import apot_quantize_tensor, generate_apot_levels, _apot_signed_levels
    from b200q import ops
    g = golden("apot")
    for key in [k for k in g.z.files if k.startswith("apot_levels/")]:
        n, k = (int(v[1:]) for v in key.split("/")[1].split("_"))
        assert torch.equal(generate_apot_levels(n, k), g.tensor(key))
    for case in g.cases("apot"):
        if case == "big":
            continue
        b, G, k = (int(v) for v in g.arr(f"apot/{case}/meta"))
        w = g.tensor(f"apot/{case}/w", case_dtype(case))
        want = g.tensor(f"apot/{case}/out", case_dtype(case))
        assert torch.equal(_apot_signed_levels(b, k), g.tensor(f"apot/{case}/levels"))
        same(apot_quantize_tensor(w.cuda(), n_bit=b, q_group_size=G, k=k), want, case)
        groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
        out, lidx, scale, idx = ops.apot_quant(groups, _apot_signed_levels(b, k),
                                               O.apot_grid(w.numel()), return_codes=True)
        assert np.array_equal(lidx.cpu().numpy().reshape(w.shape), g.arr(f"apot/{case}/level_idx")), case
        assert np.array_equal(idx.cpu().numpy(), g.arr(f"apot/{case}/best_idx")), case
        assert np.array_equal(scale.cpu().numpy(), g.arr(f"apot/{case}/scale")), case


def test_apot_golden_coarse_grid_branch(golden):
    """numel > 500000 switches the reference to the 20-point grid (pot_apot_quantizer.py:258)."""
    from pot_apot_quantizer import apot_quantize_tensor
    g = golden("apot")
    seed, n, k = (int(v) for v in g.arr("apot/big/seed_shape"))
    gen = torch.Generator().manual_seed(seed)
    w = torch.randn(n, k, generator=gen) * 0.02
    assert hashlib.sha256(w.numpy().tobytes()).digest() == g.arr("apot/big/w_sha256").tobytes()
    out = apot_quantize_tensor(w.cuda(), n_bit=4, q_group_size=128, k=2).cpu()
    assert torch.equal(out[:16], g.tensor("apot/big/out_head"))
    assert hashlib.sha256(out.numpy().tobytes()).digest() == g.arr("apot/big/out_sha256").tobytes()


@pytest.mark.parametrize("shape,b,G,k,mul", [
    ((128, 1024), 4, 128, 2, 1.0), ((32, 512), 8, 128, 2, 10.0), ((16, 256), 6, 128, 3, 1.0),
    ((20, 200), 4, 100, 2, 1.0), ((12, 768), 4, -1, 2, 1.0), ((9, 63), 3, 7, 1, 1.0),
])
def test_apot_vs_oracle(shape, b, G, k, mul):
    from pot_apot_quantizer import _apot_signed_levels
    from b200q import ops
    g = torch.Generator().manual_seed(abs(hash((shape, b, G, k))) % 2**31)
    w = torch.randn(*shape, generator=g) * 0.02 * mul
    r = O.apot_quant(w, b, G, k)
    groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
    out, lidx, scale, idx = ops.apot_quant(groups, _apot_signed_levels(b, k), O.apot_grid(w.numel()),
                                           return_codes=True)
    assert torch.equal(idx.cpu(), r["best_idx"]), "chosen grid point differs"
    assert torch.equal(lidx.cpu().reshape(shape).to(torch.int32), r["level_idx"])
    assert torch.equal(scale.cpu(), r["scale"])
    same(out.reshape(shape), r["out"])


def test_apot_exhaustive_variant_agrees():
    """A level set with sub-ulp spacing forces the exhaustive argmin kernel; it must agree with
    the oracle's literal argmin too."""
    from b200q import ops
    lv = torch.tensor([-1.0, -0.5, -0.5 + 2e-5, 0.0, 0.25, 0.25 + 3e-5, 1.0])
    g = torch.Generator().manual_seed(77)
    w = torch.randn(64, 128, generator=g) * 0.05
    grid = O.apot_grid(w.numel())
    out, lidx, scale, idx = ops.apot_quant(w.cuda(), lv, grid, return_codes=True)
    s0 = w.abs().amax(1, keepdim=True).clamp(min=1e-5)
    best_err = torch.full((64, 1), float("inf")); best_s = s0.clone()
    for b in grid:
        s = s0 * b
        q = lv[torch.argmin((w / s).unsqueeze(-1).sub(lv.view(1, 1, -1)).abs(), dim=-1)]
        err = ((w - s * q) ** 2).sum(1, keepdim=True)
        m = err < best_err
        best_err = torch.where(m, err, best_err); best_s = torch.where(m, s, best_s)
    want_idx = torch.argmin((w / best_s).unsqueeze(-1).sub(lv.view(1, 1, -1)).abs(), dim=-1)
    assert torch.equal(lidx.cpu().long(), want_idx)
    same(out, best_s * lv[want_idx])


def test_pot_apot_model_walkers():
    import torch.nn as nn
    from pot_apot_quantizer import pot_quantize_model_weight, apot_quantize_model_weight
    torch.manual_seed(3)
    for fn, orc in ((pot_quantize_model_weight, lambda w: O.pot_quant(w, 4, 128)["out"]),
                    (lambda m, b, G: apot_quantize_model_weight(m, b, G, k=2),
                     lambda w: O.apot_quant(w, 4, 128, 2)["out"])):
        net = nn.Sequential(nn.Linear(256, 64), nn.ReLU(), nn.Linear(128, 16, bias=False)).cuda()
        w0 = [net[0].weight.data.clone().cpu(), net[2].weight.data.clone().cpu()]
        fn(net, 4, 128)
        same(net[0].weight.data, orc(w0[0]))
        same(net[2].weight.data, orc(w0[1]))
        assert net[0].weight.data.is_cuda


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,b,G,mul", [((128, 1024), 4, 128, 1.0), ((32, 256), 3, 128, 40.0),
                                           ((16, 768), 4, -1, 1.0), ((12, 200), 4, 40, 1.0),
                                           ((9, 60), 4, 12, 1.0), ((16, 256), 4, 128, 2e-3)])
def test_pot_16bit_vs_oracle(dtype, shape, b, G, mul):
    from b200q import ops
    g = torch.Generator().manual_seed(abs(hash((shape, b, G, str(dtype)))) % 2**31)
    w = (torch.randn(*shape, generator=g) * 0.02 * mul).to(dtype)
    w[0, :3] = 0
    r = O.pot_quant(w, b, G)
    groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
    out, exps, scale, idx = ops.pot_quant(groups, b, O.pot_grid(), return_codes=True)
    assert torch.equal(idx.cpu(), r["best_idx"]), "chosen grid point differs"
    assert torch.equal(exps.cpu().reshape(shape).to(torch.int32), r["exps"])
    assert torch.equal(scale.cpu(), r["scale"])
    same(out.reshape(shape), r["out"])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,b,G,k", [((128, 1024), 4, 128, 2), ((32, 256), 8, 128, 2),
                                         ((16, 768), 4, -1, 2), ((12, 200), 4, 40, 2)])
def test_apot_16bit_vs_oracle(dtype, shape, b, G, k):
    from pot_apot_quantizer import _apot_signed_levels
    from b200q import ops
    g = torch.Generator().manual_seed(abs(hash((shape, b, G, k, str(dtype)))) % 2**31)
    w = (torch.randn(*shape, generator=g) * 0.02).to(dtype)
    r = O.apot_quant(w, b, G, k)
    groups = w.cuda().reshape(-1, G) if G > 0 else w.cuda()
    out, lidx, scale, idx = ops.apot_quant(groups, _apot_signed_levels(b, k), O.apot_grid(w.numel()),
                                           return_codes=True)
    assert torch.equal(idx.cpu(), r["best_idx"]), "chosen grid point differs"
    assert torch.equal(lidx.cpu().reshape(shape).to(torch.int32), r["level_idx"])
    assert torch.equal(scale.cpu(), r["scale"])
    same(out.reshape(shape), r["out"])
