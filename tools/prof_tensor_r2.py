"""Round-2 ncu target: the tensor-core kernels exactly as the headline bench launches them
(262144 tokens of bf16 activations; K = 4096 and 11008):
  * hessian_gemm2_kernel<bf16, chunked>   plain Gram matrix of the AWQ search      (normalize=False)
  * hessian_gemm2_kernel<bf16, per-sample> GPTQ Hessian, sample weights folded in  (normalize=True)
  * awq_loss_gemm_kernel                   20 candidates x 4096 rows
  * w4a16_gemm_kernel                      2048 tokens x 4096 x 4096 on a packed record
One launch of each per K, after one untimed warm-up launch."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T, export as E, qlinear as Q

tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for K in (4096, 11008):
    X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        G = T.hessian_accum(X, 2048, normalize=False)
    for _ in range(2):
        H = T.hessian_accum(X, 2048, normalize=True)
    del X, H
    W = torch.randn(4096, K, device="cuda") * 0.02
    mask = torch.zeros(K, dtype=torch.uint8, device="cuda")
    mask[:: 100] = 1
    cands = torch.linspace(1.0, 2.0, 20, dtype=torch.float64).tolist()
    for _ in range(2):
        T.awq_search_losses(W, G, mask, 4, 128, cands)
    del G
    rec = E.export_uniform(W.half(), 4, 128)
    x = torch.randn(2048, K, device="cuda", dtype=torch.float16)
    for _ in range(2):
        Q.w4a16_linear(x, rec)
    torch.cuda.synchronize()
    del W, rec, x
    T.release_workspace()
    torch.cuda.empty_cache()
print("ok")
