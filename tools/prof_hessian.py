"""Hessian/Gram GEMM at the two in_features of Llama-2-7B with the full 128 x 2048-token calibration
set (target of the ncu capture that provides bench.py's roofline.traffic)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for K in (4096, 11008):
    X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        T.hessian_accum(X, 2048, normalize=False)
    torch.cuda.synchronize()
    del X
    T.release_workspace()
    torch.cuda.empty_cache()
print("ok")
