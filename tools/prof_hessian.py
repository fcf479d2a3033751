import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
K = int(sys.argv[1]) if len(sys.argv) > 1 else 11008
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    T.hessian_accum(X, 2048)
torch.cuda.synchronize()
print("ok")
