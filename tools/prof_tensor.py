"""One or two launches of every tensor-core kernel at the Llama-2-7B shapes of the headline bench
(262144 calibration tokens; K = 4096 and 11008) -- the target of the `ncu --set full` capture that
also provides bench.py's roofline.traffic (profiles/traffic_r1.json)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cands = torch.linspace(1, 2, 20).tolist()
for K in (4096, 11008, -4096):
    gptq = K < 0
    K = abs(K)
    X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
    W = torch.randn(4096, K, device="cuda") * 0.02
    mask = torch.zeros(K, dtype=torch.uint8, device="cuda"); mask[::100] = 1
    H = None
    if not gptq:
        H = T.hessian_accum(X, 2048, normalize=False)        # Gram matrix, activations read in place
        T.awq_search_losses(W, H, mask, 4, 128, cands)       # delta + fold + persistent loss GEMM
    else:
        Hn = T.hessian_accum(X, 2048)                        # GPTQ Hessian, per-sample accumulation
        T.hessian_finalize(Hn, 1.0 / (tokens // 2048), 0.01)
        T.spd_inverse(Hn)                                    # recursive factor-and-invert (split GEMMs)
    torch.cuda.synchronize()
    del X, W, H
    T.release_workspace()
    torch.cuda.empty_cache()
print("ok")
