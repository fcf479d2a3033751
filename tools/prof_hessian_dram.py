"""One Gram GEMM launch per rasterisation band height at K = 11008 (ncu target: DRAM bytes)."""
import ctypes
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
libc = ctypes.CDLL(None)
K = 11008
X = torch.randn(262144, K, device="cuda", dtype=torch.bfloat16)
for v in [int(a) for a in sys.argv[1:]] or [2, 4, 8, 16]:
    libc.setenv(b"B200Q_HESSIAN_RASTER", str(v).encode(), 1)
    T.hessian_accum(X, 2048, normalize=False)
    torch.cuda.synchronize()
print("ok")
