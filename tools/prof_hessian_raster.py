"""Interleaved sweep of the pair kernel's rasterisation band height (B200Q_HESSIAN_RASTER is read
per call): rounds of [each value: 4 launches] so that clock / power drift hits all values alike."""
import ctypes
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
libc = ctypes.CDLL(None)
vals = [int(v) for v in sys.argv[1:]] or [2, 3, 4, 6, 8, 12]
for K in (11008, 4096):
    X = torch.randn(262144, K, device="cuda", dtype=torch.bfloat16)
    tot = {v: 0.0 for v in vals}
    for rnd in range(5):
        for v in vals:
            libc.setenv(b"B200Q_HESSIAN_RASTER", str(v).encode(), 1)
            T.hessian_accum(X, 2048, normalize=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                T.hessian_accum(X, 2048, normalize=False)
            e1.record()
            torch.cuda.synchronize()
            if rnd > 0:
                tot[v] += e0.elapsed_time(e1) / 4
    print(f"K={K}: " + ", ".join(f"raster {v}: {tot[v] / 4:.3f} ms" for v in vals), flush=True)
    del X
    T.release_workspace()
    torch.cuda.empty_cache()
