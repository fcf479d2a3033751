"""One launch of each tensor-core kernel at a Llama-2-7B layer shape (target of ncu --set full)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
K, N, tokens = 4096, 4096, 65536
X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
W = torch.randn(N, K, device="cuda") * 0.02
mask = torch.zeros(K, dtype=torch.uint8, device="cuda"); mask[::100] = 1
for _ in range(2):
    H = T.hessian_finalize(T.hessian_accum(X, 2048, normalize=False), 1.0 / tokens, 0.0)
    T.awq_search_losses(W, H, mask, 4, 128, torch.linspace(1, 2, 20).tolist())
torch.cuda.synchronize()
print("ok")
