"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
and pin oracle/quant_oracle.py against it.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the
GPU box):

    python oracle/gen_golden.py            # writes tests/golden/, exits non-zero on any mismatch

Every case is produced from a fixed torch CPU seed.  For each case the script asserts
torch.equal(oracle_out, reference_out) BEFORE writing, so a committed fixture certifies both the
reference's output and the oracle's agreement with it.  H / H_inv, which the reference computes and
drops, are captured by wrapping torch.linalg.inv while the reference runs.
"""
from __future__ import annotations

import hashlib
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("LLMQ_REFERENCE_DIR", "/root/reference"))
GOLD = REPO / "tests" / "golden"

if not (REF / "gptq_quantizer.py").exists():
    sys.exit(f"reference tree not found at {REF}")
# the reference modules import each other by bare name: put ONLY the reference on the path
sys.path.insert(0, str(REF))
sys.path.insert(1, str(REPO))

import awq_quantizer as ref_awq  # noqa: E402
import gptq_quantizer as ref_gptq  # noqa: E402
import pot_apot_quantizer as ref_pot  # noqa: E402
import quantization_utils as ref_utils  # noqa: E402
import smooth_quant_quantizer as ref_smooth  # noqa: E402

from oracle import quant_oracle as O  # noqa: E402

assert Path(ref_gptq.__file__).parent == REF, "picked up the wrong gptq_quantizer"


def npy(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().copy()
    return t.detach().numpy().copy()


def gen(seed: int, *shape, std: float = 0.02, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(dtype)


def must_equal(a: torch.Tensor, b: torch.Tensor, what: str) -> None:
    if a.dtype != b.dtype or a.shape != b.shape or not torch.equal(a, b):
        bad = (a.float() != b.float()).sum().item() if a.shape == b.shape else -1
        raise SystemExit(f"ORACLE != REFERENCE for {what}: {bad} differing elements")


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(npy(t).tobytes()).hexdigest()


DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


# --------------------------------------------------------------------------------------------------
def golden_uniform(store: dict) -> None:
    cases = [
        ("f32_b4_g128", "f32", (16, 256), 4, 128, 1.0),
        ("f32_b8_g128", "f32", (16, 256), 8, 128, 1.0),
        ("f32_b3_g32", "f32", (8, 128), 3, 32, 1.0),
        ("f32_b4_row", "f32", (32, 64), 4, -1, 50.0),
        ("f32_b2_g64_big", "f32", (4, 192), 2, 64, 50000.0),
        ("f32_b4_g128_tiny", "f32", (4, 256), 4, 128, 1e-5),
        ("f16_b4_g128", "f16", (16, 256), 4, 128, 1.0),
        ("bf16_b4_g128", "bf16", (16, 256), 4, 128, 1.0),
        ("f16_b8_row", "f16", (8, 96), 8, -1, 1.0),
    ]
    for i, (name, dt, shape, b, G, mul) in enumerate(cases):
        w = (gen(100 + i, *shape) * mul).to(DT[dt])
        ref = ref_utils.pseudo_quantize_tensor(w.clone(), n_bit=b, q_group_size=G)
        orc = O.uniform_group_quant(w.clone(), b, G)
        must_equal(orc["out"], ref, f"uniform/{name}")
        store[f"uniform/{name}/w"] = npy(w)
        store[f"uniform/{name}/out"] = npy(ref)
        store[f"uniform/{name}/codes"] = orc["codes"].numpy().astype(np.int16)
        store[f"uniform/{name}/scales"] = orc["scales"].numpy()
        store[f"uniform/{name}/zeros"] = orc["zeros"].numpy()
        store[f"uniform/{name}/meta"] = np.array([b, G])
    # constant inputs (test_quantization.py:193-199 edge cases)
    for name, val in (("ones", 1.0), ("minus_ones", -1.0), ("zeros", 0.0)):
        w = torch.full((4, 128), val)
        ref = ref_utils.pseudo_quantize_tensor(w.clone(), n_bit=4, q_group_size=128)
        must_equal(O.uniform_group_quant(w, 4, 128)["out"], ref, f"uniform/{name}")
        store[f"uniform/const_{name}/w"] = npy(w)
        store[f"uniform/const_{name}/out"] = npy(ref)
        store[f"uniform/const_{name}/meta"] = np.array([4, 128])


def golden_simple(store: dict) -> None:
    for i, (name, dt, N, K, b, G) in enumerate([
        ("f32_b4_g128", "f32", 16, 256, 4, 128),
        ("f32_b3_g64", "f32", 8, 128, 3, 64),
        ("f32_b4_row", "f32", 8, 96, 4, -1),
        ("f16_b4_g128", "f16", 16, 256, 4, 128),
        ("bf16_b8_g128", "bf16", 16, 256, 8, 128),
    ]):
        lin = nn.Linear(K, N, bias=False)
        lin.weight.data = gen(200 + i, N, K).to(DT[dt])
        w = lin.weight.data.clone()
        ref_gptq._simple_quantize_layer(lin, b, G)
        orc = O.symmetric_group_quant(w, b, G)
        must_equal(orc["out"], lin.weight.data, f"simple/{name}")
        store[f"simple/{name}/w"] = npy(w)
        store[f"simple/{name}/out"] = npy(lin.weight.data)
        store[f"simple/{name}/codes"] = orc["codes"].numpy().astype(np.int16)
        store[f"simple/{name}/scales"] = orc["scales"].numpy()
        store[f"simple/{name}/meta"] = np.array([b, G])


class _InvSpy:
    """Wraps torch.linalg.inv to capture the damped Hessian the reference builds and its inverse."""

    def __enter__(self):
        self.calls = []
        self._orig = torch.linalg.inv

        def spy(a, *args, **kw):
            r = self._orig(a, *args, **kw)
            self.calls.append((a.clone(), r.clone()))
            return r

        torch.linalg.inv = spy
        return self

    def __exit__(self, *exc):
        torch.linalg.inv = self._orig


def golden_gptq(store: dict) -> None:
    cases = [
        # name, dtype, N, K, bits, feat kind, n feats, nsamples, actorder
        ("f32_b4_2d", "f32", 24, 128, 4, "2d", 6, 128, False),
        ("f32_b4_2d_act", "f32", 24, 128, 4, "2d", 6, 4, True),
        ("f32_b3_1d_act", "f32", 16, 256, 3, "1d", 16, 128, True),
        ("f32_b8_1d", "f32", 16, 256, 8, "1d", 16, 8, False),
        ("f16_b4_1d", "f16", 16, 128, 4, "1d", 8, 128, False),
        ("bf16_b4_1d", "bf16", 16, 128, 4, "1d", 8, 128, True),
    ]
    for i, (name, dt, N, K, b, kind, nf, ns, act) in enumerate(cases):
        W = gen(300 + i, N, K).to(DT[dt])
        g = torch.Generator().manual_seed(350 + i)
        chan = torch.ones(K)
        chan[torch.randperm(K, generator=g)[: max(1, K // 50)]] = 20.0   # outlier channels
        if kind == "2d":
            feats = [(torch.randn(40, K, generator=g) * chan) for _ in range(nf)]
        else:
            feats = [(torch.randn(48, K, generator=g) * chan).abs().mean(0) for _ in range(nf)]
        # feats must share W's dtype for the reference's in-place H += (gptq_quantizer.py:144)
        feats = [f.to(DT[dt]) for f in feats]
        lin = nn.Linear(K, N, bias=False)
        lin.weight.data = W.clone()
        with _InvSpy() as spy:
            ref_gptq._gptq_quantize_layer(lin, b, 128, feats, perp_damp=0.01, blocksize=32,
                                          nsamples=ns, actorder=act, verbose=False)
        orc = O.gptq_parity_quant(W, b)
        must_equal(orc["out"], lin.weight.data, f"gptq/{name}")
        (H_reg, H_inv), = spy.calls
        H = O.gptq_hessian(feats, K, DT[dt], ns, 0.01)
        H_reg_o = H + 1e-6 * torch.eye(K, dtype=H.dtype)
        must_equal(H_reg_o, H_reg, f"gptq/{name}/H")
        must_equal(O.gptq_hinv(H), H_inv, f"gptq/{name}/Hinv")
        store[f"gptq/{name}/w"] = npy(W)
        store[f"gptq/{name}/out"] = npy(lin.weight.data)
        store[f"gptq/{name}/codes"] = orc["codes"].numpy().astype(np.int16)
        store[f"gptq/{name}/scales"] = orc["scales"].numpy()
        store[f"gptq/{name}/feats"] = np.stack([f.float().numpy() for f in feats])
        store[f"gptq/{name}/H_reg"] = H_reg.float().numpy()
        store[f"gptq/{name}/H_inv"] = H_inv.float().numpy()
        store[f"gptq/{name}/perm"] = O.gptq_perm(H, act).numpy()
        store[f"gptq/{name}/meta"] = np.array([b, ns, int(act)])


class _TinyNet(nn.Module):
    """Two calibrated Linears and one the calibration never saw."""

    def __init__(self, seed: int, dtype):
        super().__init__()
        self.fc1 = nn.Linear(256, 48, bias=False)
        self.fc2 = nn.Linear(128, 32, bias=True)
        self.head = nn.Linear(256, 8, bias=False)
        for j, m in enumerate((self.fc1, self.fc2, self.head)):
            m.weight.data = gen(seed + j, *m.weight.shape).to(dtype)


def _feat_lists(seed: int, n: int = 12):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, K in (("fc1", 256), ("fc2", 128)):
        chan = torch.ones(K)
        chan[torch.randperm(K, generator=g)[:3]] = 25.0
        out[name] = [(torch.randn(64, K, generator=g) * chan).abs().mean(0) for _ in range(n)]
    return out


def golden_walkers(store: dict) -> None:
    # ---- AWQ walker (awq_quantizer.py:22-84): scale factors 2.0 (runner default) and 1.5
    for name, dt, sf, b in (("f32_sf2", "f32", 2.0, 4), ("f32_sf1p5", "f32", 1.5, 4),
                            ("f16_sf2", "f16", 2.0, 4), ("f32_sf2_b8", "f32", 2.0, 8)):
        net = _TinyNet(400, DT[dt])
        feats = _feat_lists(410)
        w0 = {n: m.weight.data.clone() for n, m in net.named_modules() if isinstance(m, nn.Linear)}
        ref_awq.awq_quantize_model_weight(net, w_bit=b, q_group_size=128, input_feat=feats,
                                          protect_ratio=0.01, scale_factor=sf)
        for n, m in net.named_modules():
            if not isinstance(m, nn.Linear):
                continue
            if n in feats:
                orc = O.awq_layer(w0[n], feats[n], b, 128, 0.01, sf)
                must_equal(orc["out"], m.weight.data, f"awq/{name}/{n}")
                store[f"awq/{name}/{n}/salient"] = np.sort(orc["salient"].numpy())
                store[f"awq/{name}/{n}/feats"] = np.stack([f.numpy() for f in feats[n]])
            else:
                must_equal(w0[n], m.weight.data, f"awq/{name}/{n} (skipped layer)")
            store[f"awq/{name}/{n}/w"] = npy(w0[n])
            store[f"awq/{name}/{n}/out"] = npy(m.weight.data)
        store[f"awq/{name}/meta"] = np.array([b, 128, sf])

    # ---- GPTQ walker (gptq_quantizer.py:22-75): `head` has no features -> symmetric fallback
    for name, dt, b, act in (("f32_b4", "f32", 4, False), ("f32_b3_act", "f32", 3, True)):
        net = _TinyNet(420, DT[dt])
        feats = _feat_lists(430)
        w0 = {n: m.weight.data.clone() for n, m in net.named_modules() if isinstance(m, nn.Linear)}
        ref_gptq.gptq_quantize_model_weight(net, w_bit=b, q_group_size=128, input_feat=feats,
                                            nsamples=8, actorder=act, verbose=False)
        for n, m in net.named_modules():
            if not isinstance(m, nn.Linear):
                continue
            orc = O.gptq_parity_quant(w0[n], b) if n in feats else O.symmetric_group_quant(w0[n], b, 128)
            must_equal(orc["out"], m.weight.data, f"gptqwalk/{name}/{n}")
            store[f"gptqwalk/{name}/{n}/w"] = npy(w0[n])
            store[f"gptqwalk/{name}/{n}/out"] = npy(m.weight.data)
            if n in feats:
                store[f"gptqwalk/{name}/{n}/feats"] = np.stack([f.numpy() for f in feats[n]])
        store[f"gptqwalk/{name}/meta"] = np.array([b, 128, int(act)])

    # ---- SmoothQuant walker (smooth_quant_quantizer.py:268-323); alpha 0.5 is exact (sqrt)
    for name, dt, alpha, b in (("f32_a0p5", "f32", 0.5, 8), ("f32_a0p85", "f32", 0.85, 8),
                               ("f32_a0", "f32", 0.0, 4), ("f32_a1", "f32", 1.0, 4)):
        net = _TinyNet(440, DT[dt])
        g = torch.Generator().manual_seed(450)
        act = {"fc1": torch.rand(256, generator=g) * 8 + 0.01, "fc2": torch.rand(128, generator=g) * 3}
        act["fc2"][5] = 0.0   # exercises the 1e-5 clamp
        w0 = {n: m.weight.data.clone() for n, m in net.named_modules() if isinstance(m, nn.Linear)}
        ref_smooth.smoothquant_quantize_model_weight(net, w_bit=b, q_group_size=128, act_scales=act,
                                                     alpha=alpha, verbose=False)
        for n, m in net.named_modules():
            if not isinstance(m, nn.Linear):
                continue
            orc = O.smoothquant_layer(w0[n], act.get(n), alpha, b, 128)
            must_equal(orc["out"], m.weight.data, f"smooth/{name}/{n}")
            store[f"smooth/{name}/{n}/w"] = npy(w0[n])
            store[f"smooth/{name}/{n}/out"] = npy(m.weight.data)
            if n in act:
                must_equal(orc["s"], m.smoothing_scale, f"smooth/{name}/{n}/s")
                store[f"smooth/{name}/{n}/act"] = act[n].numpy()
                store[f"smooth/{name}/{n}/s"] = m.smoothing_scale.numpy()
        store[f"smooth/{name}/meta"] = np.array([b, 128, alpha])

    # smooth_weights alone (no quantisation)      smooth_quant_quantizer.py:112-199
    net = _TinyNet(460, torch.float32)
    g = torch.Generator().manual_seed(461)
    act = {"fc1": torch.rand(256, generator=g) * 8 + 0.01}
    w0 = net.fc1.weight.data.clone()
    ref_smooth.smooth_weights(net, act, alpha=0.5, verbose=False)
    must_equal(O.smooth_layer(w0, act["fc1"], 0.5)["out"], net.fc1.weight.data, "smooth_weights")
    store["smoothw/f32_a0p5/w"] = npy(w0)
    store["smoothw/f32_a0p5/act"] = act["fc1"].numpy()
    store["smoothw/f32_a0p5/out"] = npy(net.fc1.weight.data)
    store["smoothw/f32_a0p5/s"] = net.fc1.smoothing_scale.numpy()


def golden_pot(store: dict) -> None:
    cases = [
        ("f32_b4_g128", (8, 256), 4, 128, 1.0),
        ("f32_b3_g128", (4, 256), 3, 128, 1.0),
        ("f32_b8_g128", (4, 128), 8, 128, 1.0),
        ("f32_b4_row64", (32, 64), 4, -1, 50.0),      # test_quantization.py:55-58 shape
        ("f32_b4_g100", (4, 200), 4, 100, 1.0),       # vector tail + scalar tail in the row sum
        ("f32_b4_g32", (8, 128), 4, 32, 1000.0),
        ("f32_b4_g128_tiny", (4, 128), 4, 128, 1e-3),
        ("f32_b4_row1024", (3, 1024), 4, -1, 1.0),    # cascade levels of the row sum
    ]
    for i, (name, shape, b, G, mul) in enumerate(cases):
        w = gen(500 + i, *shape) * mul
        if name == "f32_b4_g128":
            w[0, :128] = 0.0                       # all-zero group
            w[1, 5] = 0.0                          # exact zero inside a group
            w[2, :128] = 0.03125                   # constant power of two
        ref = ref_pot.pot_quantize_tensor(w.clone(), n_bit=b, q_group_size=G)
        orc = O.pot_quant(w.clone(), b, G)
        must_equal(orc["out"], ref, f"pot/{name}")
        store[f"pot/{name}/w"] = npy(w)
        store[f"pot/{name}/out"] = npy(ref)
        store[f"pot/{name}/exps"] = orc["exps"].numpy().astype(np.uint8)
        store[f"pot/{name}/scale"] = orc["scale"].numpy()
        store[f"pot/{name}/best_idx"] = orc["best_idx"].numpy()
        store[f"pot/{name}/meta"] = np.array([b, G])
    # fp16 / bf16 weights: torch evaluates every op in the tensor dtype (config.json:61 is float16)
    for i, (name, dt, shape, b, G, mul) in enumerate([
        ("f16_b4_g128", "f16", (8, 256), 4, 128, 1.0),
        ("bf16_b4_g128", "bf16", (8, 256), 4, 128, 1.0),
        ("f16_b3_g64", "f16", (8, 128), 3, 64, 20.0),
        ("bf16_b4_row96", "bf16", (6, 96), 4, -1, 1.0),
        ("f16_b4_g128_small", "f16", (4, 128), 4, 128, 1e-3),   # ratios underflow in fp16
        ("f16_b4_g20", "f16", (4, 40), 4, 20, 1.0),             # one ATen vector + scalar tail
    ]):
        w = (gen(550 + i, *shape) * mul).to(DT[dt])
        if name == "f16_b4_g128":
            w[0, :128] = 0.0
            w[1, 3] = 0.0
        ref = ref_pot.pot_quantize_tensor(w.clone(), n_bit=b, q_group_size=G)
        orc = O.pot_quant(w.clone(), b, G)
        must_equal(orc["out"], ref, f"pot/{name}")
        store[f"pot/{name}/w"] = npy(w)
        store[f"pot/{name}/out"] = npy(ref)
        store[f"pot/{name}/exps"] = orc["exps"].numpy().astype(np.uint8)
        store[f"pot/{name}/scale"] = orc["scale"].numpy()
        store[f"pot/{name}/best_idx"] = orc["best_idx"].numpy()
        store[f"pot/{name}/meta"] = np.array([b, G])
    store["pot/grid"] = O.pot_grid().numpy()


def golden_apot(store: dict) -> None:
    cases = [
        ("f32_b4k2_g128", (8, 256), 4, 128, 2, 1.0),
        ("f32_b8k2_g128", (4, 256), 8, 128, 2, 1.0),   # 511 levels -> 32-level cap
        ("f32_b2k1_g128", (4, 128), 2, 128, 1, 1.0),
        ("f32_b4k2_row64", (32, 64), 4, -1, 2, 50.0),  # test_quantization.py:96 shape
        ("f32_b6k3_g64", (8, 128), 6, 64, 3, 1.0),
        ("f32_b4k2_g100", (4, 200), 4, 100, 2, 1000.0),
    ]
    for i, (name, shape, b, G, k, mul) in enumerate(cases):
        w = gen(600 + i, *shape) * mul
        if name == "f32_b4k2_g128":
            w[0, :128] = 0.0
            w[1, 7] = 0.0
        ref = ref_pot.apot_quantize_tensor(w.clone(), n_bit=b, q_group_size=G, k=k)
        orc = O.apot_quant(w.clone(), b, G, k)
        must_equal(orc["out"], ref, f"apot/{name}")
        store[f"apot/{name}/w"] = npy(w)
        store[f"apot/{name}/out"] = npy(ref)
        store[f"apot/{name}/level_idx"] = orc["level_idx"].numpy().astype(np.uint8)
        store[f"apot/{name}/scale"] = orc["scale"].numpy()
        store[f"apot/{name}/best_idx"] = orc["best_idx"].numpy()
        store[f"apot/{name}/levels"] = orc["levels"].numpy()
        store[f"apot/{name}/meta"] = np.array([b, G, k])
    for i, (name, dt, shape, b, G, k, mul) in enumerate([
        ("f16_b4k2_g128", "f16", (8, 256), 4, 128, 2, 1.0),
        ("bf16_b4k2_g128", "bf16", (8, 256), 4, 128, 2, 1.0),
        ("f16_b8k2_g64", "f16", (8, 128), 8, 64, 2, 30.0),
        ("bf16_b4k2_row96", "bf16", (6, 96), 4, -1, 2, 1.0),
    ]):
        w = (gen(660 + i, *shape) * mul).to(DT[dt])
        if name == "f16_b4k2_g128":
            w[0, :128] = 0.0
        ref = ref_pot.apot_quantize_tensor(w.clone(), n_bit=b, q_group_size=G, k=k)
        orc = O.apot_quant(w.clone(), b, G, k)
        must_equal(orc["out"], ref, f"apot/{name}")
        store[f"apot/{name}/w"] = npy(w)
        store[f"apot/{name}/out"] = npy(ref)
        store[f"apot/{name}/level_idx"] = orc["level_idx"].numpy().astype(np.uint8)
        store[f"apot/{name}/scale"] = orc["scale"].numpy()
        store[f"apot/{name}/best_idx"] = orc["best_idx"].numpy()
        store[f"apot/{name}/levels"] = orc["levels"].numpy()
        store[f"apot/{name}/meta"] = np.array([b, G, k])
    # the coarse-grid branch (numel > 500000, pot_apot_quantizer.py:258): input from the seed,
    # output kept as a digest plus the first rows
    w = gen(650, 4100, 128)
    ref = ref_pot.apot_quantize_tensor(w.clone(), n_bit=4, q_group_size=128, k=2)
    orc = O.apot_quant(w.clone(), 4, 128, 2)
    must_equal(orc["out"], ref, "apot/big")
    store["apot/big/seed_shape"] = np.array([650, 4100, 128])
    store["apot/big/w_sha256"] = np.frombuffer(bytes.fromhex(sha(w)), dtype=np.uint8)
    store["apot/big/out_sha256"] = np.frombuffer(bytes.fromhex(sha(ref)), dtype=np.uint8)
    store["apot/big/out_head"] = npy(ref[:16])
    store["apot/big/best_idx_head"] = orc["best_idx"].numpy()[:64]
    for (n, k) in ((2, 2), (4, 2), (2, 1), (1, 2), (2, 3)):
        must_equal(O.apot_levels(n, k), ref_pot.generate_apot_levels(n, k), f"apot_levels/{n},{k}")
        store[f"apot_levels/n{n}_k{k}"] = ref_pot.generate_apot_levels(n, k).numpy()


def golden_actstats(store: dict) -> None:
    # the two hook reductions, isolated from the model forward
    g = torch.Generator().manual_seed(700)
    for name, dt in (("f32", "f32"), ("f16", "f16")):
        x = (torch.randn(2, 96, 256, generator=g)).to(DT[dt])
        mean = x.view(-1, x.shape[-1]).abs().mean(dim=0)          # quantization_utils.py:231
        mx = x.reshape(-1, x.shape[-1]).abs().max(dim=0)[0]       # smooth_quant_quantizer.py:62,68
        must_equal(O.act_meanabs(x), mean, f"act/{name}/mean")
        must_equal(O.act_maxabs(x), mx, f"act/{name}/max")
        store[f"act/{name}/x"] = npy(x)
        store[f"act/{name}/meanabs"] = mean.float().numpy()
        store[f"act/{name}/maxabs"] = mx.float().numpy()


def main() -> None:
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    GOLD.mkdir(parents=True, exist_ok=True)
    families = {
        "uniform": golden_uniform, "simple": golden_simple, "gptq": golden_gptq,
        "walkers": golden_walkers, "pot": golden_pot, "apot": golden_apot, "act": golden_actstats,
    }
    for fam, fn in families.items():
        store: dict = {}
        fn(store)
        store["__provenance__"] = np.array(
            [f"reference={REF}", f"torch={torch.__version__}",
             f"cpu_capability={torch.backends.cpu.get_cpu_capability()}"])
        np.savez_compressed(GOLD / f"{fam}.npz", **store)
        size = (GOLD / f"{fam}.npz").stat().st_size
        print(f"{fam:8s}: {len(store):4d} arrays, {size/1024:.0f} KiB — oracle == reference on all cases")


if __name__ == "__main__":
    main()
