#!/usr/bin/env python
"""bench.py — quantize-only throughput of the weight-quantization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--method awq|gptq|pot|apot|smoothquant]
                    [--model llama2-7b|llama3-8b|opt-125m|tiny] [--impl reference]

One "step" = one pass of the chosen quantizer over EVERY nn.Linear of the named model shape
(random-init weights, synthetic calibration statistics), through the reference-compatible model
walker (`awq_quantize_model_weight`, ...).  Default workload: BASELINE.json configs[1], Llama-2-7B
shapes, AWQ w4 g128, fp32 weights.  At N > 1 every Linear's output rows are sharded over the ranks
(strong scaling: the model is fixed), launched by torchrun with one rank per GPU.

The JSON line carries: `value` = rows/s with weights resident in HBM, `seconds` = s per model,
`e2e` = the same through host (pinned) buffers incl. H2D/D2H, `roofline` for the dominant kernel
(CUDA-event timed inside the timed steps), `cpu_baseline` = the oracle port on this box's host
cores on a bounded per-shape sample, `clocks`, `gpu_launches`.

`--impl reference` times the CPU oracle port only (the reference is pure Python/torch and its tree
is not present on the GPU box; the oracle is pinned bit-for-bit against it by tests/golden).
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

# ------------------------------------------------------------------------------------------------
# model shapes: (name, out_features N, in_features K, count)
# ------------------------------------------------------------------------------------------------
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
N_CALIB = 128          # calibration batches -> one mean|x| vector each
DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def layer_list(model: str):
    out = []
    for name, N, K, count in MODELS[model]:
        for i in range(count):
            out.append((f"{name}.{i}", N, K))
    return out


def shard(n: int, world: int, rank: int):
    base, extra = divmod(n, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


class ShapeModel(nn.Module):
    """Holds one nn.Linear per Linear of the named architecture (weights only; never run forward)."""

    def __init__(self):
        super().__init__()
        self.layers = nn.ModuleDict()


def synth_feats(K: int, device, seed: int) -> torch.Tensor:
    """[N_CALIB, K] per-batch mean|x| statistics of N(0,1) activations with 1% outlier channels x20,
    produced by the library's own act_meanabs kernel (quantization_utils.py:231 semantics)."""
    from b200q import ops
    g = torch.Generator(device=device).manual_seed(seed)
    chan = torch.ones(K, device=device)
    chan[torch.randperm(K, device=device, generator=g)[: max(1, K // 100)]] = 20.0
    rows = []
    for _ in range(N_CALIB):
        x = torch.randn(256, K, device=device, generator=g) * chan
        rows.append(ops.act_meanabs(x))
    return torch.stack(rows)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
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
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [v.strip() for v in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the quantizers behind one switch
# ------------------------------------------------------------------------------------------------
def make_runner(method: str, feats_by_K, act_by_K):
    import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer

    def feat_dict(model):
        return {n: feats_by_K[m.in_features] for n, m in model.named_modules()
                if isinstance(m, nn.Linear)}

    if method == "awq":
        return lambda model: awq_quantizer.awq_quantize_model_weight(
            model, W_BIT, GROUP, feat_dict(model), protect_ratio=0.01, scale_factor=2.0)
    if method == "gptq":
        return lambda model: gptq_quantizer.gptq_quantize_model_weight(
            model, W_BIT, GROUP, feat_dict(model), verbose=False)
    if method == "pot":
        return lambda model: pot_apot_quantizer.pot_quantize_model_weight(model, W_BIT, GROUP)
    if method == "apot":
        return lambda model: pot_apot_quantizer.apot_quantize_model_weight(model, W_BIT, GROUP, k=2)
    if method == "smoothquant":
        return lambda model: smooth_quant_quantizer.smoothquant_quantize_model_weight(
            model, 8, GROUP, {n: act_by_K[m.in_features] for n, m in model.named_modules()
                              if isinstance(m, nn.Linear)}, alpha=0.5, verbose=False)
    raise SystemExit(f"unknown method {method}")


DOMINANT = {"awq": "group_fakequant", "gptq": "gptq_parity_quant", "smoothquant": "group_fakequant",
            "pot": "pot_quant", "apot": "apot_quant"}


def kernel_summary(_lib, name):
    """CUDA-event time of every call of C-ABI entry point `name` recorded during the timed steps
    (the library brackets its launches with events on the launching stream, see b200quant.h)."""
    q = _lib.profile_query(name)
    if q["launches"] == 0:
        return None
    return {"launches": q["launches"], "avg_ms": q["ms"] / q["launches"],
            "bytes_per_launch": q["bytes"] / q["launches"],
            "gbs": q["bytes"] / (q["ms"] * 1e-3) / 1e9 if q["ms"] > 0 else 0.0}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores, one Linear per distinct shape, extrapolated
# ------------------------------------------------------------------------------------------------
def cpu_baseline(method: str, model: str, dtype, budget_s: float = 25.0):
    from oracle import quant_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    total_s, total_rows, sample_rows, notes = 0.0, 0, 0, []
    shapes = MODELS[model]
    per_shape_budget = budget_s / len(shapes)
    for name, N, K, count in shapes:
        # POT/APOT cost ~2 us/element on a few cores: time a row slice and scale by rows
        est = {"pot": 2e-6, "apot": 1.4e-6}.get(method, 1.2e-8) * N * K
        rows = N if est <= per_shape_budget else max(64, int(N * per_shape_budget / est) // 64 * 64)
        rows = min(rows, N)
        w = (torch.randn(rows, K, generator=g) * 0.02).to(dtype)
        feats = [torch.rand(K, generator=g) for _ in range(N_CALIB)]
        act = torch.rand(K, generator=g) * 5
        t0 = time.perf_counter()
        if method == "awq":
            O.awq_layer(w, feats, W_BIT, GROUP, 0.01, 2.0)
        elif method == "gptq":
            O.gptq_parity_quant(w, W_BIT)
        elif method == "pot":
            O.pot_quant(w, W_BIT, GROUP)
        elif method == "apot":
            O.apot_quant(w, W_BIT, GROUP, 2, total_elements=N * K)
        elif method == "smoothquant":
            O.smoothquant_layer(w, act, 0.5, 8, GROUP)
        dt = time.perf_counter() - t0
        total_s += dt * (N / rows) * count
        total_rows += N * count
        sample_rows += rows
        notes.append(f"{rows}x{K}")
    return {"value": total_rows / total_s, "unit": "rows/s", "seconds_per_model": total_s,
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle port (torch CPU ops) timed once on " + ", ".join(notes) +
                      f" of {model}; whole-model time extrapolated by rows x layer count"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--method", default="awq", choices=sorted(DOMINANT))
    ap.add_argument("--model", default="llama2-7b", choices=sorted(MODELS))
    ap.add_argument("--dtype", default="f32", choices=sorted(DTYPES))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dtype = DTYPES[args.dtype]
    workload = (f"{args.model}-shape {args.method.upper()} w{8 if args.method == 'smoothquant' else W_BIT} "
                f"g{GROUP}, every nn.Linear incl. lm_head, random-init {args.dtype} weights")
    total_rows = sum(N * c for _, N, _, c in MODELS[args.model])
    total_elems = sum(N * K * c for _, N, K, c in MODELS[args.model])

    # -------------------------------------------------------------- reference arm: CPU only
    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for i in range(args.warmup + args.steps):
            r = cpu_baseline(args.method, args.model, dtype, budget_s=20.0)
            if i >= args.warmup:
                vals.append(r)
        best = max(vals, key=lambda r: r["value"])
        line = {"impl": "reference", "metric": f"{args.model}_{args.method}_w4g128_quantize_rows_per_s",
                "value": best["value"], "unit": "rows/s", "seconds": best["seconds_per_model"],
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": best["seconds_per_model"] * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload}, "cpu_baseline": best,
                "e2e": {"value": best["value"], "unit": "rows/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------- B200 arm
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
        lin = nn.Linear(K, 1, bias=False)           # placeholder weight, replaced below
        w = (torch.randn(r1 - r0, K, device=device, generator=gen) * 0.02).to(dtype)
        lin.weight = nn.Parameter(w, requires_grad=False)
        lin.out_features = r1 - r0
        model.layers[name.replace(".", "_")] = lin
        originals[name.replace(".", "_")] = w
    Ks = sorted({K for _, _, K in layers})
    feats_by_K = {K: synth_feats(K, device, 7 + K) for K in Ks}
    act_by_K = {K: feats_by_K[K].amax(0) * 4 for K in Ks}
    run = make_runner(args.method, feats_by_K, act_by_K)
    local_bytes = sum(w.numel() * w.element_size() for w in originals.values())

    def reset():
        for n, lin in model.layers.items():
            lin.weight.data = originals[n]

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def one_step():
        reset()
        if world > 1:
            with bdist.row_sharded():
                run(model)
        else:
            run(model)

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    _lib.profile_enable(False)
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ksum = kernel_summary(_lib, DOMINANT[args.method])
    kall = _lib.profile_query(None)

    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # -------------------------------------------------------------- end to end through host buffers
    e2e = None
    if not args.no_e2e:
        from b200q import pipeline
        e2e = pipeline.bench_host_roundtrip(args.method, model, originals, feats_by_K, act_by_K,
                                            W_BIT, GROUP, steps=max(1, min(args.steps, 2)),
                                            world=world, barrier=barrier)
        t = torch.tensor([e2e["ms_per_step"]], dtype=torch.float64, device=device)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": total_rows / (e2e_ms * 1e-3), "unit": "rows/s", "seconds": e2e_ms * 1e-3,
               "h2d_bytes_per_step": e2e["h2d_bytes"], "d2h_bytes_per_step": e2e["d2h_bytes"],
               "how": e2e["how"]}

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = None
    if ksum is not None:
        roofline = {"bound": "hbm", "kernel": DOMINANT[args.method], "achieved": ksum["gbs"],
                    "peak": hbm_peak, "unit": "GB/s", "frac": ksum["gbs"] / hbm_peak,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650",
                    "traffic": None, "launches_timed": ksum["launches"],
                    "avg_launch_ms": ksum["avg_ms"], "algorithmic_bytes_per_launch": ksum["bytes_per_launch"]}
    cpu = None if args.no_cpu_baseline else cpu_baseline(args.method, args.model, dtype)
    line = {
        "metric": f"{args.model}_{args.method}_w4g128_quantize_rows_per_s",
        "value": total_rows / (ms_step * 1e-3), "unit": "rows/s", "seconds": ms_step * 1e-3,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload, "linears": len(layers), "rows": total_rows,
                   "weights": total_elems, "sharding": f"output rows / {world}",
                   "l2": "inputs (per-step weight bytes >> 126 MB L2) larger than L2, no flush"},
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "clocks": clocks,
        "hbm_gbs_whole_step": 2 * local_bytes / (ms_step * 1e-3) / 1e9,
        "kernel_ms_per_step": kall["ms"] / args.steps,
    }
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
