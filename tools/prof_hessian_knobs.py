"""Time the Gram / Hessian GEMM at K = 11008 (and 4096) under the current env knobs
(B200Q_HESSIAN_PAIR / _RASTER / _SPLITS): 6 back-to-back launches after 2 warm-ups."""
import os
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B200Q_HESSIAN"))
for K in (11008, 4096):
    X = torch.randn(262144, K, device="cuda", dtype=torch.bfloat16)
    for normalize in (False, True):
        for _ in range(2):
            T.hessian_accum(X, 2048, normalize=normalize)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            T.hessian_accum(X, 2048, normalize=normalize)
        e1.record()
        torch.cuda.synchronize()
        print(f"[{tag or 'defaults'}] K={K} normalize={normalize}: {e0.elapsed_time(e1) / 6:.3f} ms", flush=True)
    del X
    T.release_workspace()
    torch.cuda.empty_cache()
