#!/usr/bin/env python
"""bench.py — quantize-only throughput of the weight-quantization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--method default|awq|awq_fixed|gptq|gptq_fast|pot|apot|smoothquant|smoothquant_search]
                    [--model llama2-7b|llama3-8b|opt-125m|tiny|matrix-NxK] [--dtype f32|f16|bf16]

One "step" = one pass of the chosen method over EVERY nn.Linear of the named model shape
(random-init weights, synthetic calibration activations), driven through the reference-compatible
entry points of llm-quantization_b200/ (`awq_search_scale_factor`, `awq_quantize_model_weight`,
`gptq_quantize_model_weight`, ...).

Default (`--method default`) = BASELINE.json's metric "Llama-7B GPTQ/AWQ w4g128 quantize seconds":
the top-level line is configs[1] -- Llama-2-7B shapes, AWQ w4 g128 WITH the 20-point scale grid
search (per Linear: per-batch mean|x| statistics of the 128 x 2048-token calibration activations,
the Gram matrix X^T X (tcgen05 GEMM), the 20 candidate reconstruction losses tr(dW H dW^T)
(tcgen05 GEMM), the fused scale/quantize/unscale pass with the winning factor) -- and the sibling
block `"gptq"` is the same model through `gptq_quantize_model_weight` (Hessian + damped inverse +
column stage), timed in the same process in its own region, with its own `roofline`, `e2e`,
`cpu_baseline` and `stages`.
At N > 1 every Linear's output rows are sharded over the ranks and the calibration samples are
dealt to them (Gram / Hessian partials all-reduced over NCCL, H^-1 broadcast, candidate losses
all-reduced): strong scaling.

JSON line: `value` = rows/s with weights and activations resident in HBM; `seconds` = s per model;
`e2e` = the same entry points on a model whose weights sit in pinned HOST memory (H2D / D2H of every
weight inside the timed region; activations stay on the device, where the out-of-scope forward pass
leaves them -- gptq_quantizer.py:243-246); `roofline` = the dominant kernel, timed with CUDA events
inside the timed steps by the library itself: `achieved` counts the MMAs the kernel EXECUTES (the
SYRK runs the tiles that touch the upper triangle), `algorithmic_tflops` the full 2*T*K^2 of
SURVEY.md 8(d); `cpu_baseline` = the reference's own functions (imported unmodified from
baseline/_ref) on this box's host cores on a bounded sample; `clocks`; `gpu_launches`.

`--impl reference` times the reference's CPU implementation only: the unmodified reference modules
staged in baseline/_ref (by __graft_entry__.build()), driven through their own entry points on
fixed row / token samples and extrapolated; the oracle port is the fallback when the staged files
are absent, and the stand-in for the AWQ search, whose body is a stub in the reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for _p in (str(REPO / "llm-quantization_b200"), str(REPO)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# the GPTQ walker runs up to 16 Hessian inverses side by side on their own streams: give them their
# own hardware work queues (the default of 8 makes streams share queues and serialise)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

# (name, out_features N, in_features K, count)
MODELS = {
    "llama2-7b": [("attn.qkvo", 4096, 4096, 128), ("mlp.gate_up", 11008, 4096, 64),
                  ("mlp.down", 4096, 11008, 32), ("lm_head", 32000, 4096, 1)],
    "llama3-8b": [("attn.qo", 4096, 4096, 64), ("attn.kv", 1024, 4096, 64),
                  ("mlp.gate_up", 14336, 4096, 64), ("mlp.down", 4096, 14336, 32),
                  ("lm_head", 128256, 4096, 1)],
    "opt-125m": [("attn.qkvo", 768, 768, 48), ("fc1", 3072, 768, 12), ("fc2", 768, 3072, 12),
                 ("lm_head", 50272, 768, 1)],
    # one transformer block of Llama-2-7B (the model is 32 of these plus lm_head): the unit the ncu
    # launch list is taken on -- the factorisation graphs alone are ~1100 kernels per Linear
    "llama2-7b-block": [("attn.qkvo", 4096, 4096, 4), ("mlp.gate_up", 11008, 4096, 2),
                        ("mlp.down", 4096, 11008, 1)],
    "tiny": [("a", 512, 1024, 4), ("b", 1024, 512, 2)],
}
W_BIT, GROUP = 4, 128
N_CALIB, CALIB_TOKENS = 128, 2048        # calibration batches x tokens per batch
N_GRID = 20
DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
METHODS = ("default", "awq", "awq_fixed", "gptq", "gptq_fast", "pot", "apot", "smoothquant",
           "smoothquant_search")
NEEDS_ACTS = ("awq", "gptq")
# dominant C-ABI entry point per method: (name, bound)
DOMINANT = {"awq": ("hessian_gemm", "tensor"), "gptq": ("hessian_gemm", "tensor"),
            "awq_fixed": ("group_fakequant", "hbm"), "gptq_fast": ("gptq_parity_quant", "hbm"),
            "smoothquant": ("group_fakequant", "hbm"), "pot": ("pot_quant", "hbm"),
            "apot": ("apot_quant", "hbm"), "smoothquant_search": ("smooth_alpha_errors", "hbm")}
STAGE_NAMES = ("hessian_gemm", "hessian_prescale", "hessian_reduce", "awq_search_gemm", "awq_search_delta",
               "awq_search_fold", "act_meanabs", "group_fakequant", "gptq_parity_quant", "col_absmax",
               "spd_inverse", "pot_quant", "apot_quant", "seq_sum_rows", "smooth_alpha_errors",
               "smooth_scale")


def layer_list(model: str):
    return [(f"{name}.{i}", N, K) for name, N, K, count in MODELS[model] for i in range(count)]


def shard(n: int, world: int, rank: int):
    base, extra = divmod(n, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


class ShapeModel(nn.Module):
    """One nn.Linear per Linear of the named architecture (weights only; forward is never run)."""

    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleDict()


def synth_acts(K: int, device, seed: int, n: int, tokens: int) -> torch.Tensor:
    """[n, tokens, K] bf16 calibration activations: N(0,1) with 1 % of the channels scaled x20."""
    g = torch.Generator(device=device).manual_seed(seed)
    chan = torch.ones(K, device=device)
    chan[torch.randperm(K, device=device, generator=g)[: max(1, K // 100)]] = 20.0
    out = torch.empty((n, tokens, K), dtype=torch.bfloat16, device=device)
    for i in range(n):
        out[i] = (torch.randn(tokens, K, device=device, generator=g) * chan).to(torch.bfloat16)
    return out


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [v.strip() for v in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# one step of each method through the public entry points
# ------------------------------------------------------------------------------------------------
def make_step(method: str, acts_by_K, stats_by_K, act_scale_by_K, host: bool = False):
    import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer
    from b200q import pipeline

    def by_layer(model, table):
        return {n: table[m.in_features] for n, m in model.named_modules() if isinstance(m, nn.Linear)}

    if method == "awq":
        def step(model):
            # (a10) per-batch mean|x| statistics, as the calibration hooks would collect them
            # (a data-parallel calibration run produces them per rank: batches dealt, rows gathered);
            # the host-weights run gets them on the host, where the reference's hooks leave them
            stats = stats_by_K if host else \
                {K: awq_quantizer._stat_rows(x, x.device) for K, x in acts_by_K.items()}
            # (a9) 20-point grid search on the raw activations, (a8) quantize with the winner.
            # Host-resident weights: the device copy made for the search serves the quantize pass
            # too (pipeline.keep_resident), so every weight crosses PCIe once in each direction.
            with pipeline.keep_resident():
                best = awq_quantizer.awq_search_scale_factor(model, W_BIT, GROUP, by_layer(model, acts_by_K),
                                                             protect_ratio=0.01, n_grid=N_GRID)
                awq_quantizer.awq_quantize_model_weight(model, W_BIT, GROUP, by_layer(model, stats),
                                                        protect_ratio=0.01, scale_factor=best)
            return best
        return step
    if method == "awq_fixed":
        return lambda model: awq_quantizer.awq_quantize_model_weight(
            model, W_BIT, GROUP, by_layer(model, stats_by_K), protect_ratio=0.01, scale_factor=2.0)
    if method in ("gptq", "gptq_fast"):
        def step(model):
            gptq_quantizer.BUILD_HESSIAN = method == "gptq"
            feats = by_layer(model, acts_by_K if method == "gptq" else stats_by_K)
            gptq_quantizer.gptq_quantize_model_weight(model, W_BIT, GROUP, feats, actorder=True,
                                                      verbose=False)
        return step
    if method == "pot":
        return lambda model: pot_apot_quantizer.pot_quantize_model_weight(model, W_BIT, GROUP)
    if method == "apot":
        return lambda model: pot_apot_quantizer.apot_quantize_model_weight(model, W_BIT, GROUP, k=2)
    if method == "smoothquant":
        return lambda model: smooth_quant_quantizer.smoothquant_quantize_model_weight(
            model, 8, GROUP, by_layer(model, act_scale_by_K), alpha=0.5, verbose=False)
    if method == "smoothquant_search":
        def step(model):
            # (a18) 20-point alpha sweep, (a17) smooth + quantize with the winner
            scales = by_layer(model, act_scale_by_K)
            with pipeline.keep_resident():
                alpha = smooth_quant_quantizer.smoothquant_search_alpha(model, [], scales, 8, GROUP,
                                                                        n_grid=N_GRID, verbose=False)
                smooth_quant_quantizer.smoothquant_quantize_model_weight(model, 8, GROUP, scales,
                                                                         alpha=alpha, verbose=False)
            return alpha
        return step
    raise SystemExit(f"unknown method {method}")


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own functions on the host cores, FIXED samples per distinct shape
# ------------------------------------------------------------------------------------------------
CPU_TOKENS = 2048        # calibration tokens fed to the activation-side stages


def cpu_rows(method: str, N: int, K: int) -> int:
    """Weight rows per distinct Linear shape the CPU arm works on.  A fixed function of the shape
    (NOT of a time budget): both arms and every step use the same slices."""
    per = {"pot": 1 << 19, "apot": 1 << 20, "smoothquant_search": 1 << 19, "awq": 1 << 21}.get(method, 1 << 23)
    return min(N, max(64, per // K // 64 * 64))
_REF_CACHE = {}


def _reference_modules():
    """The UNMODIFIED reference modules from baseline/_ref (staged by __graft_entry__.build()),
    imported under private names so they cannot shadow the drop-in modules; None when absent."""
    if "mods" in _REF_CACHE:
        return _REF_CACHE["mods"]
    mods = None
    ref = REPO / "baseline" / "_ref"
    if (ref / "gptq_quantizer.py").exists():
        import importlib.util
        mods = {}
        saved = {k: sys.modules.get(k) for k in ("quantization_utils",)}
        try:
            # the reference's quantizer modules do `from quantization_utils import ...` at import
            # time: give them THEIR quantization_utils for the duration of the import
            for name in ("quantization_utils", "gptq_quantizer", "awq_quantizer", "pot_apot_quantizer",
                         "smooth_quant_quantizer"):
                spec = importlib.util.spec_from_file_location(f"_llmq_ref_{name}", ref / f"{name}.py")
                mod = importlib.util.module_from_spec(spec)
                if name == "quantization_utils":
                    sys.modules["quantization_utils"] = mod
                spec.loader.exec_module(mod)
                mods[name] = mod
        except Exception as exc:  # missing optional dependency of the reference, ...
            print(f"[bench] reference import failed ({exc!r}); falling back to the oracle port",
                  file=sys.stderr)
            mods = None
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    _REF_CACHE["mods"] = mods
    return mods


def _best_of(fn, reps: int = 2):
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(method: str, model: str, dtype, tokens_total: int):
    """Seconds per model of the reference's CPU path, from fixed samples: per distinct Linear shape
    cpu_rows() weight rows and CPU_TOKENS calibration tokens, scaled by rows / tokens / layer
    count.  kind = "reference" when the staged reference modules ran.  `sample_seconds` is the wall
    time this call actually spent (one "step" of the reference arm)."""
    from oracle import quant_oracle as O
    ref = _reference_modules()          # (first call imports transformers / datasets: not timed)
    t_call = time.perf_counter()
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    total_s, total_rows, notes = 0.0, 0, []
    legs = set()
    for name, N, K, count in MODELS[model]:
        rows = cpu_rows(method, N, K)
        w = (torch.randn(rows, K, generator=g) * 0.02).to(dtype)
        feats = [torch.rand(K, generator=g) for _ in range(N_CALIB)]
        act = torch.rand(K, generator=g) * 5
        lin = nn.Linear(K, rows, bias=False)
        net = nn.Sequential(lin)

        def fresh():
            lin.weight.data = w.clone()

        row_scaled = 0.0     # seconds that scale with the number of weight rows
        fixed = 0.0          # seconds per layer independent of the rows (Hessian, inverse, stats)
        if method in ("awq", "gptq"):
            tok = min(tokens_total, CPU_TOKENS)
            x = torch.randn(tok, K, generator=g).to(dtype)
        if method == "awq":
            # (a10) the hook's reduction, quantization_utils.py:231, per calibration batch
            t_stat = _best_of(lambda: x.view(-1, K).abs().mean(dim=0))
            fixed += t_stat * (tokens_total / tok)
            # (a9) search: the reference's body is a stub -- time the builder's restatement of its
            # docstring (fp32 Gram matrix on the sample + per-candidate quantize and tr(dW H dW^T))
            t0 = time.perf_counter()
            H = (x.float().T @ x.float()) / tok
            fixed += (time.perf_counter() - t0) * (tokens_total / tok)
            imp = sum(feats)
            sal = torch.topk(imp, max(1, int(K * 0.01)))[1]
            cands = torch.linspace(1, 2, N_GRID, dtype=torch.float64).tolist()
            t0 = time.perf_counter()
            losses = O.awq_search_losses(w.float(), H, sal, W_BIT, GROUP, cands)
            row_scaled += time.perf_counter() - t0
            legs.add("search: builder restatement (reference is a stub)")
            best = cands[int(torch.argmin(losses))]
            # (a8) the reference's own walker
            if ref:
                def run():
                    fresh()
                    ref["awq_quantizer"].awq_quantize_model_weight(net, W_BIT, GROUP, {"0": feats}, 0.01, best)
                row_scaled += _best_of(run)
                legs.add("awq_quantize_model_weight: imported reference")
            else:
                row_scaled += _best_of(lambda: O.awq_layer(w, feats, W_BIT, GROUP, 0.01, best))
        elif method == "gptq":
            if ref:
                # the reference's _gptq_quantize_layer on 1 and on 5 calibration samples of `tok`
                # tokens: the difference is 4x the per-sample Hessian cost (gptq_quantizer.py:137-144),
                # the rest (inverse :160-165, K-iteration column loop :173-197) is per layer
                def run(n):
                    fresh()
                    ref["gptq_quantizer"]._gptq_quantize_layer(lin, W_BIT, GROUP, [x] * n, actorder=True,
                                                               verbose=False)
                t1 = _best_of(lambda: run(1), 1)
                t5 = _best_of(lambda: run(5), 1)
                per_sample = max(t5 - t1, 0.0) / 4
                fixed += per_sample * (tokens_total / tok) + max(t1 - per_sample, 0.0)
                legs.add("_gptq_quantize_layer: imported reference (its column loop is a Python loop "
                         "over K columns whose cost barely depends on the rows: counted per layer)")
            else:
                t0 = time.perf_counter()
                H = O.gptq_hessian([x], K, torch.float32, 128, 0.01)
                fixed += (time.perf_counter() - t0) * (tokens_total / tok)
                fixed += _best_of(lambda: O.gptq_hinv(H), 1)
                row_scaled += _best_of(lambda: O.gptq_parity_quant(w, W_BIT))
        elif method == "awq_fixed":
            if ref:
                def run():
                    fresh()
                    ref["awq_quantizer"].awq_quantize_model_weight(net, W_BIT, GROUP, {"0": feats}, 0.01, 2.0)
                row_scaled += _best_of(run)
            else:
                row_scaled += _best_of(lambda: O.awq_layer(w, feats, W_BIT, GROUP, 0.01, 2.0))
        elif method == "gptq_fast":
            row_scaled += _best_of(lambda: O.gptq_parity_quant(w, W_BIT))
            legs.add("closed form of the reference's column loop (oracle port)")
        elif method == "pot":
            fn = ref["pot_apot_quantizer"].pot_quantize_tensor if ref else \
                (lambda t, n_bit, q_group_size: O.pot_quant(t, n_bit, q_group_size))
            row_scaled += _best_of(lambda: fn(w, n_bit=W_BIT, q_group_size=GROUP), 1)
        elif method == "apot":
            # (the grid follows from the element count of the tensor handed in: the slice is kept
            # above 500000 elements so the reference picks the same 20-point grid as for the layer)
            if ref and rows * K > 500000:
                row_scaled += _best_of(lambda: ref["pot_apot_quantizer"].apot_quantize_tensor(
                    w, n_bit=W_BIT, q_group_size=GROUP, k=2), 1)
            else:
                row_scaled += _best_of(lambda: O.apot_quant(w, W_BIT, GROUP, 2, total_elements=N * K), 1)
        elif method == "smoothquant":
            if ref:
                def run():
                    fresh()
                    if hasattr(lin, "smoothing_scale"):
                        del lin.smoothing_scale
                    ref["smooth_quant_quantizer"].smoothquant_quantize_model_weight(
                        net, 8, GROUP, {"0": act}, alpha=0.5, verbose=False)
                row_scaled += _best_of(run)
            else:
                row_scaled += _best_of(lambda: O.smoothquant_layer(w, act, 0.5, 8, GROUP))
        elif method == "smoothquant_search":
            alphas = torch.linspace(0, 1, N_GRID, dtype=torch.float64).tolist()
            t0 = time.perf_counter()
            S = torch.stack([O.smooth_scale(act.clamp(min=1e-5), w, a).float() for a in alphas])
            errs = O.smooth_alpha_errors(w, S.to(w.dtype).float(), act, 8, GROUP)
            O.smoothquant_layer(w, act, alphas[int(torch.argmin(errs))], 8, GROUP)
            row_scaled += time.perf_counter() - t0
            legs.add("alpha sweep: builder restatement (reference is a stub)")
        dt = row_scaled * (N / rows) + fixed
        total_s += dt * count
        total_rows += N * count
        notes.append(f"{rows}x{K}")
    kind = "reference" if (ref and method in ("awq", "awq_fixed", "gptq", "pot", "apot", "smoothquant")) else "port"
    return {"value": total_rows / total_s, "unit": "rows/s", "seconds_per_model": total_s,
            "sample_seconds": time.perf_counter() - t_call,
            "cores": torch.get_num_threads(), "kind": kind,
            "legs": sorted(legs),
            "sample": ("unmodified reference functions (baseline/_ref) " if kind == "reference" else
                       "oracle port (torch CPU ops) ") +
                      "on all host threads, once per distinct Linear shape on the first rows " +
                      ", ".join(notes) +
                      (f" and {min(tokens_total, CPU_TOKENS)} calibration tokens" if method in NEEDS_ACTS else "") +
                      f" of {model}; whole-model time extrapolated by rows, tokens and layer count"}


def executed_fraction_syrk(K: int) -> float:
    """Share of the 128 x 256 output tiles of X^T X that touch the upper triangle (= run)."""
    tm, tn = -(-K // 128), -(-K // 256)
    return sum(1 for m in range(tm) for n in range(tn) if (n + 1) * 256 > m * 128) / (tm * tn)


def describe(method: str, args, bits: int) -> str:
    return {"awq": f"AWQ w{bits} g{GROUP} with the {N_GRID}-point scale grid search "
                   f"(activation stats + Gram matrix + candidate losses + quantize), "
                   f"{args.calib_batches} x {args.calib_tokens}-token calibration activations (bf16)",
            "awq_fixed": f"AWQ w{bits} g{GROUP}, fixed scale factor 2.0 (benchmark_runner flow)",
            "gptq": f"GPTQ w{bits} act-order: Hessian + damped inverse + reference-parity column stage, "
                    f"{args.calib_batches} x {args.calib_tokens}-token calibration activations (bf16)",
            "gptq_fast": f"GPTQ w{bits}, reference-parity column stage only (H, H^-1 cannot reach the output)",
            "pot": f"POT w{bits} g{GROUP}, 200-point scale search",
            "apot": f"APOT w{bits} g{GROUP} k2, 20-point scale search",
            "smoothquant": f"SmoothQuant w{bits} g{GROUP} alpha 0.5 (weight side; the reference has no "
                           f"activation quantizer, smooth_quant_quantizer.py:363-371)",
            "smoothquant_search": f"SmoothQuant w{bits} g{GROUP} with the {N_GRID}-point alpha sweep "
                                  f"(weight-side reconstruction error per alpha + smooth + quantize; the "
                                  f"reference has no activation (A8) quantizer)"}[method]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--method", default="default", choices=METHODS,
                    help="default = AWQ + search (top-level line) and GPTQ (sibling block 'gptq')")
    ap.add_argument("--model", default="llama2-7b",
                    help="one of " + ", ".join(sorted(MODELS)) + ", or matrix-NxK for a single Linear "
                         "(BASELINE configs[4]: the 4096x4096 ... 28672x8192 sweep)")
    ap.add_argument("--bits", type=int, default=W_BIT, choices=[2, 3, 4, 8],
                    help="weight bits of the GPTQ / AWQ / POT / APOT methods (BASELINE configs[2]: w3, w4)")
    ap.add_argument("--dtype", default="f32", choices=sorted(DTYPES))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--calib-batches", type=int, default=N_CALIB)
    ap.add_argument("--calib-tokens", type=int, default=CALIB_TOKENS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.model.startswith("matrix-"):
        try:
            n_, k_ = (int(v) for v in args.model[len("matrix-"):].lower().split("x"))
        except ValueError:
            ap.error("--model matrix-NxK needs two integers, e.g. matrix-28672x8192")
        if n_ <= 0 or k_ <= 0 or k_ % GROUP:
            ap.error(f"--model matrix-NxK: K must be a positive multiple of {GROUP}")
        MODELS[args.model] = [("matrix", n_, k_, 1)]
    elif args.model not in MODELS:
        ap.error(f"unknown --model {args.model}")
    globals()["W_BIT"] = args.bits
    # the contract is ONE JSON line on stdout: progress text of the entry points (the reference's
    # functions print, e.g. "Searching for optimal scale factor...") goes to stderr
    # (NCCL and other native libraries write to file descriptor 1 directly: keep a private copy of
    # it for the JSON line and point fd 1 at stderr)
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dtype = DTYPES[args.dtype]
    tokens_total = args.calib_batches * args.calib_tokens
    methods = ["awq", "gptq"] if args.method == "default" else [args.method]
    primary = methods[0]
    total_rows = sum(N * c for _, N, _, c in MODELS[args.model])
    total_elems = sum(N * K * c for _, N, K, c in MODELS[args.model])

    def bits_of(m):
        return 8 if m.startswith("smoothquant") else W_BIT

    def metric_of(m):
        return f"{args.model}_{m}_w{bits_of(m)}g{GROUP}_quantize_rows_per_s"

    def workload_of(m):
        return (f"{args.model}-shape {describe(m, args, bits_of(m))}; every nn.Linear incl. lm_head, "
                f"random-init {args.dtype} weights")

    def config_of(m):
        return {"workload": workload_of(m), "linears": len(layer_list(args.model)), "rows": total_rows,
                "weights": total_elems,
                "sharding": f"output rows / {world}, calibration samples / {world}",
                "l2": "inputs larger than L2 (per-step weight and activation bytes >> 126 MB), no flush"}

    if args.impl == "reference":
        # The reference's CPU implementation on this box's host cores.  One STEP = one pass over the
        # bounded sample (fixed row / token slices of every distinct Linear shape); `ms_per_step`
        # is the wall time of such a pass, `value` / `seconds` the whole-model figure extrapolated
        # from the mean over the timed passes.  The sibling method of the default run is sampled
        # once (its pass includes 11008-wide LAPACK inverses: ~40 s).
        if rank != 0:
            return
        blocks = {}
        for i, m in enumerate(methods):
            once = i > 0 or m in ("pot", "apot", "gptq")
            n_warm, n_steps = (0, 1) if once else (args.warmup, args.steps)
            vals = [cpu_baseline(m, args.model, dtype, tokens_total) for _ in range(n_warm + n_steps)][n_warm:]
            secs = sum(v["seconds_per_model"] for v in vals) / len(vals)
            best = dict(vals[-1])
            best.update(value=total_rows / secs, seconds_per_model=secs)
            blocks[m] = {
                "impl": "reference", "metric": metric_of(m), "value": total_rows / secs, "unit": "rows/s",
                "seconds": secs,
                "ms_per_step": sum(v["sample_seconds"] for v in vals) / len(vals) * 1e3,
                "steps_timed": len(vals),
                "config": config_of(m), "cpu_baseline": best,
                "e2e": {"value": total_rows / secs, "unit": "rows/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        line = dict(blocks[primary])
        line.update({"n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                     "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                     "dtype": args.dtype, "data": "synthetic", "gpu_launches": 0})
        for m in methods[1:]:
            line[m] = blocks[m]
        print(json.dumps(line), file=out, flush=True)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path exists)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as td
    if world > 1:
        # NCCL's kernels on a high-priority stream: the block scheduler then places them ahead of the
        # thousands of pending GEMM blocks instead of at the GEMM's tail, so an exchange really runs
        # WHILE the next layer's Gram GEMM does (the quantizers queue exchanges on a side stream)
        opts = None
        try:
            opts = td.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        td.init_process_group("nccl", device_id=device, pg_options=opts)
    from b200q import _lib, ops, dist as bdist

    layers = layer_list(args.model)
    model = ShapeModel()
    originals = {}
    gen = torch.Generator(device=device).manual_seed(1000 + rank)
    for name, N, K in layers:
        r0, r1 = shard(N, world, rank)
        key = name.replace(".", "_")
        lin = nn.Linear(K, 1, bias=False)
        w = (torch.randn(r1 - r0, K, device=device, generator=gen) * 0.02).to(dtype)
        lin.weight = nn.Parameter(w, requires_grad=False)
        lin.out_features = r1 - r0
        model.layers[key] = lin
        originals[key] = w
    Ks = sorted({K for _, _, K in layers})
    acts_by_K = {}
    if any(m in NEEDS_ACTS for m in methods):
        acts_by_K = {K: synth_acts(K, device, 7 + K, args.calib_batches, args.calib_tokens) for K in Ks}
        stats_by_K = {K: ops.act_meanabs_batched(x).to(x.dtype) for K, x in acts_by_K.items()}
    else:
        small = {K: synth_acts(K, device, 7 + K, args.calib_batches, 64) for K in Ks}
        stats_by_K = {K: ops.act_meanabs_batched(x).float() for K, x in small.items()}
        del small
    act_scale_by_K = {K: v.float().amax(0) * 4 for K, v in stats_by_K.items()}

    def reset(m, src):
        for n, lin in m.layers.items():
            lin.weight.data = src[n]

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {}
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    # the host-resident copy of the model (pinned), shared by the e2e runs of all methods
    host_pack = {}

    def host_model():
        if host_pack:
            return host_pack["model"], host_pack["flat"], host_pack["total"]
        names = list(originals)
        total = sum(originals[n].numel() for n in names)
        flat = torch.empty(total, dtype=dtype, pin_memory=True)
        hm, off = ShapeModel(), 0
        for n in names:
            w = originals[n]
            view = flat[off:off + w.numel()].view(w.shape)
            view.copy_(w)
            off += w.numel()
            lin = nn.Linear(w.shape[1], 1, bias=False)
            lin.weight = nn.Parameter(view, requires_grad=False)
            hm.layers[n] = lin
        torch.cuda.synchronize()
        host_pack.update(model=hm, flat=flat, total=total, src={n: hm.layers[n].weight.data for n in names})
        return hm, flat, total

    def run_method(method: str):
        step_fn = make_step(method, acts_by_K, stats_by_K, act_scale_by_K)

        def one_step():
            reset(model, originals)
            if world > 1:
                with bdist.row_sharded():
                    return step_fn(model)
            return step_fn(model)

        for _ in range(args.warmup):
            one_step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = _lib.launch_count()
        _lib.profile_enable(True)
        bdist.WAIT_EVENTS = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        walker_timings = None
        if os.environ.get("B200Q_WALKER_TIMINGS") and method == "gptq":
            import gptq_quantizer as _gq
            walker_timings = _gq.TIMINGS = []
            _gq.HOST_LAPS = {}
            _gq.TRACE = []
        barrier()
        e0.record()
        result = None
        for _ in range(args.steps):
            result = one_step()
        e1.record()
        barrier()
        if walker_timings is not None:
            torch.cuda.synchronize()
            phases = {}
            for ph, a, b in walker_timings:
                phases[ph] = phases.get(ph, 0.0) + a.elapsed_time(b) / args.steps
            print(f"[rank {rank}] walker phases (ms/step): " +
                  ", ".join(f"{k} {v:.1f}" for k, v in phases.items()), file=sys.stderr, flush=True)
            print(f"[rank {rank}] walker host seconds per step: " +
                  ", ".join(f"{k} {v / args.steps:.3f}" for k, v in _gq.HOST_LAPS.items()),
                  file=sys.stderr, flush=True)
            last = None
            for g0, b, e, K in _gq.TRACE[:16 * 3] + _gq.TRACE[-16 * 4:-16 * 2]:
                if last is not g0:
                    print(f"[trace] group of K={K}:", file=sys.stderr)
                    last = g0
                print(f"[trace]   begin +{g0.elapsed_time(b):8.2f} ms  end +{g0.elapsed_time(e):8.2f} ms",
                      file=sys.stderr)
            _gq.TIMINGS = _gq.HOST_LAPS = _gq.TRACE = None
        _lib.profile_enable(False)
        ms_total = e0.elapsed_time(e1)
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        # time the main stream spent waiting for collectives inside the timed steps (b200q.dist
        # brackets every wait with two events when WAIT_EVENTS is a list) and the bytes they moved
        nccl_wait_ms = sum(a.elapsed_time(b) for a, b, _n in bdist.WAIT_EVENTS) / args.steps
        nccl_bytes = sum(n for _a, _b, n in bdist.WAIT_EVENTS) / args.steps
        bdist.WAIT_EVENTS = None
        dom_name, dom_bound = DOMINANT[method]
        kq = {n: _lib.profile_query(n) for n in STAGE_NAMES}
        kall = _lib.profile_query(None)

        t = torch.tensor([ms_total, nccl_wait_ms], dtype=torch.float64, device=device)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_step = float(t[0].item()) / args.steps
        nccl_wait_ms = float(t[1].item())

        # ------------------------------------------------------------ end to end, host-resident weights
        e2e = None
        if not args.no_e2e:
            reset(model, originals)          # drop the device-resident quantized copies first
            hm, flat, total = host_model()
            # per-batch statistics arrive on the host, as the reference's hooks produce them (.cpu())
            stats_host = {K: v.cpu() for K, v in stats_by_K.items()}
            scale_host = {K: v.cpu() for K, v in act_scale_by_K.items()}
            host_step = make_step(method, acts_by_K, stats_host, scale_host, host=True)

            def run_host():
                reset(hm, host_pack["src"])
                if world > 1:
                    with bdist.row_sharded():
                        host_step(hm)
                else:
                    host_step(hm)

            # the quantized weights overwrite the pinned host tensors in place: restore them
            # (untimed) so that every timed pass quantizes the original weights
            def restore():
                off = 0
                for n in originals:
                    w = originals[n]
                    flat[off:off + w.numel()].view(w.shape).copy_(w)
                    off += w.numel()
                torch.cuda.synchronize()

            run_host()
            n_e2e = max(1, min(args.steps, 2))
            e2e_ms = 0.0
            for _ in range(n_e2e):
                restore()
                barrier()
                t0 = time.perf_counter()
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                run_host()
                h1.record()
                torch.cuda.synchronize()
                e2e_ms += max((time.perf_counter() - t0) * 1e3, h0.elapsed_time(h1))
            e2e_ms /= n_e2e
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
            if world > 1:
                td.all_reduce(t, op=td.ReduceOp.MAX)
            e2e_ms = float(t.item())
            nbytes = total * flat.element_size()
            stat_bytes = sum(stats_host[l.in_features].numel() * stats_host[l.in_features].element_size()
                             for l in hm.layers.values()) if method.startswith(("awq", "gptq")) else 0
            e2e = {"value": total_rows / (e2e_ms * 1e-3), "unit": "rows/s", "seconds": e2e_ms * 1e-3,
                   "h2d_bytes_per_step": nbytes + stat_bytes, "d2h_bytes_per_step": nbytes,
                   "how": "same entry points on a model whose weights live in pinned host memory: per Linear "
                          "H2D prefetch / kernels / D2H overlap on three streams (b200q.pipeline); a "
                          "search pass keeps its device copies for the quantize pass "
                          "(pipeline.keep_resident), so each weight crosses PCIe once per direction; "
                          "calibration activations stay on the device; wall clock vs CUDA events, the larger"}

        roofline = None
        q = kq[dom_name]
        if q["launches"] > 0 and q["ms"] > 0:
            avg_ms = q["ms"] / q["launches"]
            if dom_bound == "tensor":
                # timed inside a seconds-long step under the power cap -> sustained bf16 peak
                peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
                algo = q["flops"] / (q["ms"] * 1e-3) / 1e12
                # the Gram/Hessian kernel is a SYRK: it runs only the 128x256 tiles that touch the
                # upper triangle and mirrors them, i.e. it EXECUTES about half of the algorithmic
                # 2*T*K^2 flops.  The roofline fraction counts the MMAs actually issued.
                wsum = sum(K * K for _, _, K in layers)
                exe = sum(K * K * executed_fraction_syrk(K) for _, _, K in layers) / wsum
                # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel
                # variant at the same token count (profiles/traffic_r2.json), averaged over layers
                traffic = None
                tj = REPO / "profiles" / "traffic_r2.json"
                if tj.exists() and world == 1:
                    tr = json.loads(tj.read_text())
                    byK = tr.get(method, tr).get("dram_bytes_per_launch_by_K", {})
                    if tr.get("tokens") == tokens_total and all(str(K) in byK for _, _, K in layers):
                        traffic = sum(byK[str(K)] for _, _, K in layers) / len(layers)
                roofline = {"bound": "tensor", "kernel": dom_name, "achieved": algo * exe, "peak": peak,
                            "unit": "TFLOP/s", "frac": algo * exe / peak,
                            "peak_source": src + ", sustained bf16 (kernel timed inside a long step)",
                            "traffic": traffic,
                            "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write, "
                                            "profiles/traffic_r2.json)",
                            "launches_timed": q["launches"], "avg_launch_ms": avg_ms,
                            "algorithmic_flops_per_launch": q["flops"] / q["launches"],
                            "algorithmic_tflops": algo, "executed_mma_fraction": exe,
                            "note": "achieved = MMAs the kernel issues (SYRK: the 128x256 tiles touching "
                                    "the upper triangle, mirrored in the epilogue) / CUDA-event time; "
                                    "algorithmic_tflops = SURVEY 8(d)'s full 2*T*K^2 count over the same time"}
            else:
                peak = float(peaks.get("hbm_gbs", 6650.0))
                ach = q["bytes"] / (q["ms"] * 1e-3) / 1e9
                roofline = {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "peak_source": src, "traffic": None,
                            "launches_timed": q["launches"], "avg_launch_ms": avg_ms,
                            "algorithmic_bytes_per_launch": q["bytes"] / q["launches"]}
        stages = {n: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                      **({"tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12} if v["flops"] > 0 and v["ms"] > 0 else {}),
                      **({"gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9} if v["bytes"] > 0 and v["ms"] > 0 else {})}
                  for n, v in kq.items() if v["launches"] > 0}
        if "awq_search_gemm" in stages and "tflops" in stages["awq_search_gemm"]:
            # the search GEMM multiplies by H folded onto its lower triangle and stops each output
            # tile's k-loop at the tile's right edge: it executes about half of the algorithmic flops
            def tri_fraction(K):
                tn = -(-K // 256)
                return sum(min(K, (j + 1) * 256) for j in range(tn)) / (tn * K)
            wsum = sum(N * K * K for _, N, K in layers)
            fr = sum(N * K * K * tri_fraction(K) for _, N, K in layers) / wsum
            stages["awq_search_gemm"]["executed_mma_fraction"] = fr
            stages["awq_search_gemm"]["executed_tflops"] = stages["awq_search_gemm"]["tflops"] * fr
        if "spd_inverse" in stages and method == "gptq" and world == 1:
            stages["spd_inverse"]["note"] = ("inverses of up to 8 layers run concurrently on side streams: "
                                             "ms_per_step is the SUM of per-stream durations, not wall time")
        if world > 1:
            stages["nccl_exposed"] = {"ms_per_step": nccl_wait_ms, "bytes_per_step_per_rank": nccl_bytes,
                                      "note": "time the compute stream waited on NCCL work inside the timed "
                                              "steps (max over ranks) and the payload bytes of this rank's "
                                              "collectives per step"}
        cpu = None if (args.no_cpu_baseline or world > 1 or rank != 0) else \
            cpu_baseline(method, args.model, dtype, tokens_total)
        return {
            "metric": metric_of(method), "value": total_rows / (ms_step * 1e-3), "unit": "rows/s",
            "seconds": ms_step * 1e-3, "ms_per_step": ms_step,
            "config": config_of(method),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "clocks": clocks, "stages": stages, "kernel_ms_per_step": kall["ms"] / args.steps,
            "result": result if isinstance(result, float) else None,
        }

    blocks = {m: run_method(m) for m in methods}
    if rank == 0:
        line = dict(blocks[primary])
        line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                     "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                     "dtype": args.dtype, "data": "synthetic"})
        for m in methods[1:]:
            line[m] = blocks[m]
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
