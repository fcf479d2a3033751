"""Per-stage timing of the tensor-core stages at BASELINE shapes (CUDA events via the library's
profiler).   python tools/bench_stages.py [hessian|inverse|search] ..."""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import _lib, tensor_ops as T

which = sys.argv[1:] or ["hessian"]
peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {}
PEAK = peaks.get("bf16_tflops", 1590.0)


def timed(names, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    return {n: _lib.profile_query(n) for n in names}


if "hessian" in which:
    for K, tokens, rows in ((4096, 262144, 2048), (11008, 262144, 2048), (768, 262144, 2048),
                            (4096, 128, 1)):
        n = tokens // rows
        X = (torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16))
        q = timed(["hessian_prescale", "hessian_gemm", "hessian_reduce"], lambda: T.hessian_accum(X, rows))
        g = q["hessian_gemm"]
        ms = g["ms"] / g["launches"]
        tf = g["flops"] / g["launches"] / (ms * 1e-3) / 1e12
        pre = q["hessian_prescale"]["ms"] / q["hessian_prescale"]["launches"]
        red = q["hessian_reduce"]["ms"] / q["hessian_reduce"]["launches"]
        print(json.dumps({"stage": "hessian", "K": K, "T": tokens, "gemm_ms": round(ms, 3),
                          "tflops": round(tf, 1), "frac_of_bf16_peak": round(tf / PEAK, 3),
                          "prescale_ms": round(pre, 3), "reduce_ms": round(red, 3)}))
        del X
        T.release_workspace()
        torch.cuda.empty_cache()
