"""The oracle against the committed golden vectors (which gen_golden.py produced by running the
unmodified reference).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import Golden, case_dtype
from oracle import quant_oracle as O


def _same(a: torch.Tensor, b: torch.Tensor):
    assert a.dtype == b.dtype and a.shape == b.shape
    assert torch.equal(a, b), f"{(a.float() != b.float()).sum().item()} differing elements"


def test_uniform_group_quant_matches_reference(golden):
    g = golden("uniform")
    for case in g.cases("uniform"):
        dt = case_dtype(case) if not case.startswith("const") else torch.float32
        b, G = (int(v) for v in g.arr(f"uniform/{case}/meta"))
        w = g.tensor(f"uniform/{case}/w", dt)
        _same(O.uniform_group_quant(w, b, G)["out"], g.tensor(f"uniform/{case}/out", dt))


def test_symmetric_group_quant_matches_reference(golden):
    g = golden("simple")
    for case in g.cases("simple"):
        dt = case_dtype(case)
        b, G = (int(v) for v in g.arr(f"simple/{case}/meta"))
        _same(O.symmetric_group_quant(g.tensor(f"simple/{case}/w", dt), b, G)["out"],
              g.tensor(f"simple/{case}/out", dt))


def test_gptq_layer_matches_reference(golden):
    g = golden("gptq")
    for case in g.cases("gptq"):
        dt = case_dtype(case)
        b, ns, act = (int(v) for v in g.arr(f"gptq/{case}/meta"))
        W = g.tensor(f"gptq/{case}/w", dt)
        r = O.gptq_parity_quant(W, b)
        _same(r["out"], g.tensor(f"gptq/{case}/out", dt))
        assert r["codes"].abs().max() <= 2 ** b            # codes in [-2^b, 2^b-1]
        feats = [f.to(dt) for f in g.tensor(f"gptq/{case}/feats")]
        H = O.gptq_hessian(feats, W.shape[1], dt, ns, 0.01)
        H_reg = H + 1e-6 * torch.eye(W.shape[1], dtype=H.dtype)
        assert torch.equal(H_reg.float(), g.tensor(f"gptq/{case}/H_reg"))
        assert torch.equal(O.gptq_perm(H, bool(act)), g.tensor(f"gptq/{case}/perm"))
        # LAPACK's inverse is not bit-reproducible across hosts: compare numerically
        torch.testing.assert_close(O.gptq_hinv(H).float(), g.tensor(f"gptq/{case}/H_inv"),
                                   rtol=2e-3, atol=2e-3)


def test_awq_layer_matches_reference(golden):
    g = golden("walkers")
    for case in g.cases("awq"):
        if case == "meta":
            continue
    for case in ("f32_sf2", "f32_sf1p5", "f16_sf2", "f32_sf2_b8"):
        dt = case_dtype(case)
        b, G, sf = g.arr(f"awq/{case}/meta")
        for layer in ("fc1", "fc2", "head"):
            w = g.tensor(f"awq/{case}/{layer}/w", dt)
            want = g.tensor(f"awq/{case}/{layer}/out", dt)
            if g.has(f"awq/{case}/{layer}/feats"):
                feats = list(g.tensor(f"awq/{case}/{layer}/feats"))
                r = O.awq_layer(w, feats, int(b), int(G), 0.01, float(sf))
                _same(r["out"], want)
                assert np.array_equal(np.sort(r["salient"].numpy()),
                                      g.arr(f"awq/{case}/{layer}/salient"))
            else:
                _same(w, want)          # uncalibrated layers are left untouched


def test_smoothquant_layer_matches_reference(golden):
    g = golden("walkers")
    for case in ("f32_a0p5", "f32_a0p85", "f32_a0", "f32_a1"):
        b, G, alpha = g.arr(f"smooth/{case}/meta")
        for layer in ("fc1", "fc2", "head"):
            w = g.tensor(f"smooth/{case}/{layer}/w")
            act = g.tensor(f"smooth/{case}/{layer}/act") if g.has(f"smooth/{case}/{layer}/act") else None
            # the stored s pins the pow() results of the generating host; feed it back so this
            # check is independent of the local libm
            s = g.tensor(f"smooth/{case}/{layer}/s") if act is not None else None
            r = O.smoothquant_layer(w, act, float(alpha), int(b), int(G), s=s)
            _same(r["out"], g.tensor(f"smooth/{case}/{layer}/out"))
            if act is not None:
                torch.testing.assert_close(O.smooth_scale(act, w, float(alpha)), s, rtol=1e-6, atol=0)
    w = g.tensor("smoothw/f32_a0p5/w")
    r = O.smooth_layer(w, g.tensor("smoothw/f32_a0p5/act"), 0.5)
    _same(r["out"], g.tensor("smoothw/f32_a0p5/out"))      # alpha = 0.5 is sqrt: exact everywhere
    _same(r["s"], g.tensor("smoothw/f32_a0p5/s"))


def test_pot_matches_reference(golden):
    g = golden("pot")
    assert np.array_equal(O.pot_grid().numpy(), g.arr("pot/grid"))
    for case in g.cases("pot"):
        b, G = (int(v) for v in g.arr(f"pot/{case}/meta"))
        r = O.pot_quant(g.tensor(f"pot/{case}/w", case_dtype(case)), b, G)
        _same(r["out"], g.tensor(f"pot/{case}/out", case_dtype(case)))
        assert np.array_equal(r["exps"].numpy().astype(np.uint8), g.arr(f"pot/{case}/exps"))
        assert np.array_equal(r["best_idx"].numpy(), g.arr(f"pot/{case}/best_idx"))


def test_apot_matches_reference(golden):
    g = golden("apot")
    for case in g.cases("apot"):
        if case == "big":
            continue
        b, G, k = (int(v) for v in g.arr(f"apot/{case}/meta"))
        r = O.apot_quant(g.tensor(f"apot/{case}/w", case_dtype(case)), b, G, k)
        _same(r["out"], g.tensor(f"apot/{case}/out", case_dtype(case)))
        assert np.array_equal(r["level_idx"].numpy().astype(np.uint8), g.arr(f"apot/{case}/level_idx"))
    for key in [k for k in g.z.files if k.startswith("apot_levels/")]:
        n, k = (int(v[1:]) for v in key.split("/")[1].split("_"))
        assert torch.equal(O.apot_levels(n, k), g.tensor(key))


def test_act_stats_match_reference(golden):
    g = golden("act")
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        x = g.tensor(f"act/{name}/x", dt)
        assert torch.equal(O.act_meanabs(x).float(), g.tensor(f"act/{name}/meanabs"))
        assert torch.equal(O.act_maxabs(x).float(), g.tensor(f"act/{name}/maxabs"))
