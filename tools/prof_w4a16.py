"""Timing of the dequant-fused W4A16 GEMM against a cuBLAS fp16 GEMM on the dequantised weight."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import export as E, qlinear as Q


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for M, N, K in ((2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008), (8192, 4096, 4096), (16, 4096, 4096)):
    W = (torch.randn(N, K, device="cuda") * 0.02).half()
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    rec = E.export_uniform(W, 4, 128)
    Wd = E.dequantize(rec)
    t_q = timed(lambda: Q.w4a16_linear(x, rec))
    t_d = timed(lambda: x @ Wd.T)
    fl = 2.0 * M * N * K
    print(f"M={M:5d} N={N:5d} K={K:5d}: w4a16 {t_q * 1e3:8.1f} us = {fl / t_q / 1e9:7.1f} TF/s | "
          f"cuBLAS fp16 on the dequantised copy {t_d * 1e3:8.1f} us = {fl / t_d / 1e9:7.1f} TF/s", flush=True)
