#!/usr/bin/env python
"""bench.py — quantize-only throughput of the weight-quantization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--method awq|awq_fixed|gptq|gptq_fast|pot|apot|smoothquant]
                    [--model llama2-7b|llama3-8b|opt-125m|tiny] [--dtype f32|f16|bf16]

One "step" = one pass of the chosen method over EVERY nn.Linear of the named model shape
(random-init weights, synthetic calibration activations), driven through the reference-compatible
entry points of llm-quantization_b200/ (`awq_search_scale_factor`, `awq_quantize_model_weight`,
`gptq_quantize_model_weight`, ...).

Default workload = BASELINE.json configs[1]: Llama-2-7B shapes, AWQ w4 g128 WITH the 20-point scale
grid search.  Per Linear that is: per-batch mean|x| statistics of the 128 x 2048-token calibration
activations, the Gram matrix X^T X (tcgen05 GEMM), the 20 candidate reconstruction losses
tr(dW H dW^T) (tcgen05 GEMM), and the fused scale/quantize/unscale pass with the winning factor.
At N > 1 every Linear's output rows are sharded over the ranks and the calibration samples are
dealt to them (Gram partials all-reduced over NCCL, candidate losses all-reduced): strong scaling.

JSON line: `value` = rows/s with weights and activations resident in HBM; `seconds` = s per model;
`e2e` = the same entry points on a model whose weights sit in pinned HOST memory (H2D / D2H of every
weight inside the timed region; activations stay on the device, where the out-of-scope forward pass
leaves them — gptq_quantizer.py:243-246); `roofline` = the dominant kernel, timed with CUDA events
inside the timed steps by the library itself; `cpu_baseline` = the oracle port on this box's host
cores on a bounded sample; `clocks`; `gpu_launches`.

`--impl reference` times the CPU oracle port only: the reference is pure Python/torch, its tree
is absent on the GPU box, and the oracle is pinned bit-for-bit against it by tests/golden.
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
    "tiny": [("a", 512, 1024, 4), ("b", 1024, 512, 2)],
}
W_BIT, GROUP = 4, 128
N_CALIB, CALIB_TOKENS = 128, 2048        # calibration batches x tokens per batch
N_GRID = 20
DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
METHODS = ("awq", "awq_fixed", "gptq", "gptq_fast", "pot", "apot", "smoothquant", "smoothquant_search")
NEEDS_ACTS = ("awq", "gptq")
# dominant C-ABI entry point per method: (name, bound)
DOMINANT = {"awq": ("hessian_gemm", "tensor"), "gptq": ("hessian_gemm", "tensor"),
            "awq_fixed": ("group_fakequant", "hbm"), "gptq_fast": ("gptq_parity_quant", "hbm"),
            "smoothquant": ("group_fakequant", "hbm"), "pot": ("pot_quant", "hbm"),
            "apot": ("apot_quant", "hbm"), "smoothquant_search": ("smooth_alpha_errors", "hbm")}


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
def make_step(method: str, acts_by_K, stats_by_K, act_scale_by_K):
    import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer
    from b200q import ops

    def by_layer(model, table):
        return {n: table[m.in_features] for n, m in model.named_modules() if isinstance(m, nn.Linear)}

    if method == "awq":
        def step(model):
            # (a10) per-batch mean|x| statistics, as the calibration hooks would collect them
            # (a data-parallel calibration run produces them per rank: batches dealt, rows gathered)
            stats = {K: awq_quantizer._stat_rows(x, x.device) for K, x in acts_by_K.items()}
            # (a9) 20-point grid search on the raw activations, (a8) quantize with the winner
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
            alpha = smooth_quant_quantizer.smoothquant_search_alpha(model, [], scales, 8, GROUP,
                                                                    n_grid=N_GRID, verbose=False)
            smooth_quant_quantizer.smoothquant_quantize_model_weight(model, 8, GROUP, scales,
                                                                     alpha=alpha, verbose=False)
            return alpha
        return step
    raise SystemExit(f"unknown method {method}")


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores, bounded sample per distinct shape, extrapolated
# ------------------------------------------------------------------------------------------------
def cpu_baseline(method: str, model: str, dtype, tokens_total: int, budget_s: float = 25.0):
    from oracle import quant_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    shapes = MODELS[model]
    per = budget_s / len(shapes)
    total_s, total_rows, notes = 0.0, 0, []
    for name, N, K, count in shapes:
        elem_cost = {"pot": 2e-6, "apot": 1.4e-6, "smoothquant_search": 5e-7}.get(
            method, 2e-8 if method != "awq" else 6e-7)
        rows = N if elem_cost * N * K <= per else max(64, int(per / (elem_cost * K)) // 64 * 64)
        rows = min(rows, N)
        w = (torch.randn(rows, K, generator=g) * 0.02).to(dtype)
        feats = [torch.rand(K, generator=g) for _ in range(N_CALIB)]
        act = torch.rand(K, generator=g) * 5
        extra = 0.0
        t0 = time.perf_counter()
        if method in ("awq", "gptq"):
            # activation-side work on a token sample, scaled to the full calibration set
            tok = min(tokens_total, 2048)
            x = torch.randn(tok, K, generator=g)
            t1 = time.perf_counter()
            if method == "awq":
                O.act_meanabs(x)
                H = (x.T @ x) / tok
            else:
                H = O.gptq_hessian([x], K, torch.float32, 128, 0.01)
                O.gptq_hinv(H)
            extra = (time.perf_counter() - t1)
            t_inv = 0.0
            if method == "gptq":
                t2 = time.perf_counter(); O.gptq_hinv(H); t_inv = time.perf_counter() - t2
            # the X-dependent part scales with tokens, the inverse does not
            extra = (extra - t_inv) * (tokens_total / tok) + t_inv
            t0 = time.perf_counter()
            if method == "awq":
                imp = sum(feats)
                sal = torch.topk(imp, max(1, int(K * 0.01)))[1]
                cands = torch.linspace(1, 2, N_GRID, dtype=torch.float64).tolist()
                losses = O.awq_search_losses(w.float(), H, sal, W_BIT, GROUP, cands)
                O.awq_layer(w, feats, W_BIT, GROUP, 0.01, cands[int(torch.argmin(losses))])
            else:
                O.gptq_parity_quant(w, W_BIT)
        elif method == "awq_fixed":
            O.awq_layer(w, feats, W_BIT, GROUP, 0.01, 2.0)
        elif method == "gptq_fast":
            O.gptq_parity_quant(w, W_BIT)
        elif method == "pot":
            O.pot_quant(w, W_BIT, GROUP)
        elif method == "apot":
            O.apot_quant(w, W_BIT, GROUP, 2, total_elements=N * K)
        elif method == "smoothquant":
            O.smoothquant_layer(w, act, 0.5, 8, GROUP)
        elif method == "smoothquant_search":
            alphas = torch.linspace(0, 1, N_GRID, dtype=torch.float64).tolist()
            S = torch.stack([O.smooth_scale(act.clamp(min=1e-5), w, a).float() for a in alphas])
            errs = O.smooth_alpha_errors(w, S.to(w.dtype).float(), act, 8, GROUP)
            O.smoothquant_layer(w, act, alphas[int(torch.argmin(errs))], 8, GROUP)
        dt = (time.perf_counter() - t0) * (N / rows) + extra
        total_s += dt * count
        total_rows += N * count
        notes.append(f"{rows}x{K}")
    return {"value": total_rows / total_s, "unit": "rows/s", "seconds_per_model": total_s,
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle port (torch CPU ops, all host threads) timed once per distinct Linear shape "
                      "on row slices " + ", ".join(notes) +
                      (f" and a {min(tokens_total, 2048)}-token activation sample" if method in NEEDS_ACTS else "") +
                      f" of {model}; whole-model time extrapolated by rows, tokens and layer count"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--method", default="awq", choices=METHODS)
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
    out = sys.stdout
    sys.stdout = sys.stderr

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dtype = DTYPES[args.dtype]
    tokens_total = args.calib_batches * args.calib_tokens
    bits = 8 if args.method.startswith("smoothquant") else W_BIT
    what = {"awq": f"AWQ w{bits} g{GROUP} with the {N_GRID}-point scale grid search "
                   f"(activation stats + Gram matrix + candidate losses + quantize), "
                   f"{args.calib_batches} x {args.calib_tokens}-token calibration activations (bf16)",
            "awq_fixed": f"AWQ w{bits} g{GROUP}, fixed scale factor 2.0 (benchmark_runner flow)",
            "gptq": f"GPTQ w{bits} act-order: Hessian + damped inverse + reference-parity column stage, "
                    f"{args.calib_batches} x {args.calib_tokens}-token calibration activations (bf16)",
            "gptq_fast": f"GPTQ w{bits}, reference-parity column stage only (H, H^-1 cannot reach the output)",
            "pot": f"POT w{bits} g{GROUP}, 200-point scale search",
            "apot": f"APOT w{bits} g{GROUP} k2, 20-point scale search",
            "smoothquant": f"SmoothQuant w{bits} g{GROUP} alpha 0.5",
            "smoothquant_search": f"SmoothQuant w{bits} g{GROUP} with the {N_GRID}-point alpha sweep "
                                  f"(reconstruction error per alpha + smooth + quantize)"}[args.method]
    workload = f"{args.model}-shape {what}; every nn.Linear incl. lm_head, random-init {args.dtype} weights"
    total_rows = sum(N * c for _, N, _, c in MODELS[args.model])
    total_elems = sum(N * K * c for _, N, K, c in MODELS[args.model])
    metric = f"{args.model}_{args.method}_w{bits}g{GROUP}_quantize_rows_per_s"

    if args.impl == "reference":
        if rank != 0:
            return
        vals = [cpu_baseline(args.method, args.model, dtype, tokens_total, budget_s=20.0)
                for _ in range(args.warmup + args.steps)][args.warmup:]
        best = max(vals, key=lambda r: r["value"])
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": best["value"], "unit": "rows/s",
            "seconds": best["seconds_per_model"], "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": best["seconds_per_model"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": {"workload": workload}, "cpu_baseline": best,
            "e2e": {"value": best["value"], "unit": "rows/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), file=out, flush=True)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path exists)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    import torch.distributed as td
    if world > 1:
        td.init_process_group("nccl", device_id=device)
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
    if args.method in NEEDS_ACTS:
        acts_by_K = {K: synth_acts(K, device, 7 + K, args.calib_batches, args.calib_tokens) for K in Ks}
        stats_by_K = {K: ops.act_meanabs_batched(x).to(x.dtype) for K, x in acts_by_K.items()}
    else:
        small = {K: synth_acts(K, device, 7 + K, args.calib_batches, 64) for K in Ks}
        stats_by_K = {K: ops.act_meanabs_batched(x).float() for K, x in small.items()}
        del small
    act_scale_by_K = {K: v.float().amax(0) * 4 for K, v in stats_by_K.items()}
    step_fn = make_step(args.method, acts_by_K, stats_by_K, act_scale_by_K)
    local_bytes = sum(w.numel() * w.element_size() for w in originals.values())

    def reset(m, src):
        for n, lin in m.layers.items():
            lin.weight.data = src[n]

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def one_step(m=model):
        if m is model:
            reset(model, originals)
        if world > 1:
            with bdist.row_sharded():
                return step_fn(m)
        return step_fn(m)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    walker_timings = None
    if os.environ.get("B200Q_WALKER_TIMINGS") and args.method == "gptq":
        import gptq_quantizer as _gq
        walker_timings = _gq.TIMINGS = []
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
        _gq.TIMINGS = None
    _lib.profile_enable(False)
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dom_name, dom_bound = DOMINANT[args.method]
    kq = {n: _lib.profile_query(n) for n in
          ("hessian_gemm", "hessian_prescale", "hessian_reduce", "awq_search_gemm", "awq_search_delta", "awq_search_fold",
           "act_meanabs", "group_fakequant", "gptq_parity_quant", "col_absmax", "spd_inverse",
           "pot_quant", "apot_quant", "seq_sum_rows", "smooth_alpha_errors", "smooth_scale")}
    kall = _lib.profile_query(None)

    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # ---------------------------------------------------------------- end to end, host-resident weights
    e2e = None
    if not args.no_e2e:
        names = list(originals)
        total = sum(originals[n].numel() for n in names)
        flat = torch.empty(total, dtype=dtype, pin_memory=True)
        host_model, host_src, off = ShapeModel(), {}, 0
        for n in names:
            w = originals[n]
            view = flat[off:off + w.numel()].view(w.shape)
            view.copy_(w)
            off += w.numel()
            lin = nn.Linear(w.shape[1], 1, bias=False)
            lin.weight = nn.Parameter(view, requires_grad=False)
            host_model.layers[n] = lin
        torch.cuda.synchronize()
        # per-batch statistics arrive on the host, as the reference's hooks produce them (.cpu())
        stats_host = {K: v.cpu() for K, v in stats_by_K.items()}
        scale_host = {K: v.cpu() for K, v in act_scale_by_K.items()}
        host_step = make_step(args.method, acts_by_K, stats_host, scale_host)
        if args.method == "awq":
            # the search reads device activations; only the quantize call consumes host statistics
            import awq_quantizer

            def host_step(m, _acts=acts_by_K):   # noqa: F811
                tbl = lambda t: {n: t[l.in_features] for n, l in m.named_modules() if isinstance(l, nn.Linear)}  # noqa: E731
                best = awq_quantizer.awq_search_scale_factor(m, W_BIT, GROUP, tbl(_acts), 0.01, n_grid=N_GRID)
                awq_quantizer.awq_quantize_model_weight(m, W_BIT, GROUP, tbl(stats_host), 0.01, best)

        def run_host():
            if world > 1:
                with bdist.row_sharded():
                    host_step(host_model)
            else:
                host_step(host_model)

        run_host()
        barrier()
        n_e2e = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(n_e2e):
            run_host()
        h1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        e2e_ms = max(wall_ms, h0.elapsed_time(h1) / n_e2e)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_ms = float(t.item())
        passes = 2 if args.method == "awq" else 1      # search + quantize each stream the weights in
        nbytes = total * flat.element_size()
        stat_bytes = sum(stats_host[l.in_features].numel() * stats_host[l.in_features].element_size()
                         for l in host_model.layers.values()) if args.method.startswith(("awq", "gptq")) else 0
        e2e = {"value": total_rows / (e2e_ms * 1e-3), "unit": "rows/s", "seconds": e2e_ms * 1e-3,
               "h2d_bytes_per_step": passes * nbytes + stat_bytes, "d2h_bytes_per_step": nbytes,
               "how": "same entry points on a model whose weights live in pinned host memory: per Linear "
                      "H2D prefetch / kernels / D2H overlap on three streams (b200q.pipeline); calibration "
                      "activations stay on the device; wall clock vs CUDA events, the larger"}
        del host_model, flat

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {}
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    q = kq[dom_name]
    roofline = None
    if q["launches"] > 0 and q["ms"] > 0:
        avg_ms = q["ms"] / q["launches"]
        if dom_bound == "tensor":
            # timed inside a seconds-long step under the power cap -> sustained bf16 peak
            peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
            ach = q["flops"] / (q["ms"] * 1e-3) / 1e12
            # the Gram/Hessian kernel is a SYRK: it runs only the 128x256 tiles that touch the upper
            # triangle and mirrors them, so it EXECUTES about half of the algorithmic 2*T*K^2 flops
            def executed_fraction(K):
                tm, tn = -(-K // 128), -(-K // 256)
                return sum(1 for m in range(tm) for n in range(tn) if (n + 1) * 256 > m * 128) / (tm * tn)
            wsum = sum(K * K for _, _, K in layers)
            exe = sum(K * K * executed_fraction(K) for _, _, K in layers) / wsum
            # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel at
            # the same token count (profiles/traffic_r1.json), averaged over this model's layers
            traffic = None
            tj = REPO / "profiles" / "traffic_r1.json"
            if tj.exists() and world == 1:
                tr = json.loads(tj.read_text())
                byK = tr.get("dram_bytes_per_launch_by_K", {})
                if tr.get("tokens") == tokens_total and all(str(K) in byK for _, _, K in layers):
                    traffic = sum(byK[str(K)] for _, _, K in layers) / len(layers)
            roofline = {"bound": "tensor", "kernel": dom_name, "achieved": ach, "peak": peak,
                        "unit": "TFLOP/s", "frac": ach / peak,
                        "peak_source": src + ", sustained bf16 (kernel timed inside a long step)",
                        "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read+write)",
                        "launches_timed": q["launches"], "avg_launch_ms": avg_ms,
                        "algorithmic_flops_per_launch": q["flops"] / q["launches"],
                        "executed_mma_fraction": exe, "executed_tflops": ach * exe,
                        "executed_frac_of_peak": ach * exe / peak,
                        "note": "achieved = algorithmic 2*T*K^2 flops (SURVEY 8d, full count) / CUDA-event "
                                "time; X^T X is symmetric and the kernel runs only the tiles touching the "
                                "upper triangle (mirror-written in the epilogue), so achieved can exceed "
                                "the GEMM peak; executed_* count the MMAs actually issued"}
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
    # the CPU baseline is a rank-0, single-GPU-run figure (the host cores are shared by all ranks)
    cpu = None if (args.no_cpu_baseline or world > 1 or rank != 0) else \
        cpu_baseline(args.method, args.model, dtype, tokens_total)
    line = {
        "metric": metric, "value": total_rows / (ms_step * 1e-3), "unit": "rows/s",
        "seconds": ms_step * 1e-3, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload, "linears": len(layers), "rows": total_rows,
                   "weights": total_elems, "sharding": f"output rows / {world}, calibration samples / {world}",
                   "l2": "inputs larger than L2 (per-step weight and activation bytes >> 126 MB), no flush"},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "stages": stages, "kernel_ms_per_step": kall["ms"] / args.steps,
        "result": result if isinstance(result, float) else None,
    }
    print(json.dumps(line), file=out, flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
