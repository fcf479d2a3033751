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

if "inverse" in which:
    for K in (4096, 11008):
        X = torch.randn(8192, K, device="cuda", dtype=torch.bfloat16)
        H = T.hessian_finalize(T.hessian_accum(X, 2048), 1.0 / 4, 0.01)
        del X
        for name, kw in (("inverse", dict(want_inverse=True, want_upper=False)),
                         ("upper_factor", dict(want_inverse=False, want_upper=True))):
            qq = timed(["spd_inverse", "inv_potrf", "inv_trtri"], lambda: T.spd_inverse(H, ridge=1e-6, **kw), reps=2)
            q = qq["spd_inverse"]
            ms = q["ms"] / q["launches"]
            print(json.dumps({"stage": name, "K": K, "ms": round(ms, 2),
                              "potrf_ms": round(qq["inv_potrf"]["ms"] / qq["inv_potrf"]["launches"], 2),
                              "trtri_ms": round(qq["inv_trtri"]["ms"] / qq["inv_trtri"]["launches"], 2),
                              "tflops_fp32": round(q["flops"] / q["launches"] / (ms * 1e-3) / 1e12, 1)}))
        T.release_workspace(); torch.cuda.empty_cache()

if "search" in which:
    for N, K in ((4096, 4096), (11008, 4096), (4096, 11008)):
        W = torch.randn(N, K, device="cuda") * 0.02
        X = torch.randn(4096, K, device="cuda", dtype=torch.bfloat16)
        H = T.hessian_finalize(T.hessian_accum(X, 2048, normalize=False), 1.0 / 4096, 0.0)
        mask = torch.zeros(K, dtype=torch.uint8, device="cuda"); mask[:: 100] = 1
        cands = torch.linspace(1, 2, 20).tolist()
        q = timed(["awq_search_delta", "awq_search_gemm"],
                  lambda: T.awq_search_losses(W, H, mask, 4, 128, cands), reps=2)
        g = q["awq_search_gemm"]; d = q["awq_search_delta"]
        ms = g["ms"] / g["launches"]
        print(json.dumps({"stage": "awq_search", "N": N, "K": K, "gemm_ms": round(ms, 3),
                          "tflops": round(g["flops"] / g["launches"] / (ms * 1e-3) / 1e12, 1),
                          "frac_of_bf16_peak": round(g["flops"] / g["launches"] / (ms * 1e-3) / 1e12 / PEAK, 3),
                          "delta_ms": round(d["ms"] / d["launches"], 3),
                          "delta_gbs": round(d["bytes"] / d["ms"] / 1e6, 0)}))
        T.release_workspace(); torch.cuda.empty_cache()

if "compensated" in which:
    for N, K in ((4096, 4096),):
        W = torch.randn(N, K, device="cuda") * 0.02
        X = torch.randn(8192, K, device="cuda", dtype=torch.bfloat16)
        H = T.hessian_finalize(T.hessian_accum(X, 2048), 1.0 / 4, 0.01)
        q = timed(["gptq_compensated", "spd_inverse"], lambda: T.gptq_compensated(W, H, 4, 128), reps=2)
        g = q["gptq_compensated"]
        print(json.dumps({"stage": "gptq_compensated", "N": N, "K": K, "ms": round(g["ms"] / g["launches"], 2),
                          "upper_factor_ms": round(q["spd_inverse"]["ms"] / q["spd_inverse"]["launches"], 2)}))

if "levels" in which:
    from b200q import ops
    sys.path.insert(0, str(REPO / "llm-quantization_b200"))
    from pot_apot_quantizer import _apot_signed_levels
    W = torch.randn(4096, 4096, device="cuda") * 0.02
    grp = W.view(-1, 128)
    for name, fn, evals in (
            ("pot_quant", lambda: ops.pot_quant(grp, 4, torch.arange(0.01, 2.01, 0.01)), 200),
            ("apot_quant", lambda: ops.apot_quant(grp, _apot_signed_levels(4, 2), torch.arange(0.01, 2.01, 0.1)), 20)):
        q = timed([name], fn, reps=3)[name]
        ms = q["ms"] / q["launches"]
        print(json.dumps({"stage": name, "shape": "4096x4096 f32 g128", "ms": round(ms, 3),
                          "gbs_algorithmic": round(q["bytes"] / q["launches"] / ms / 1e6, 1),
                          "T_candidate_evals_per_s": round(W.numel() * evals / (ms * 1e-3) / 1e12, 3)}))
