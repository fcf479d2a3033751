"""Per-stage timings of the AWQ search at one layer shape (rows N of one rank's shard)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T, _lib
for N, K, tokens in [(4096, 4096, 32768), (512, 4096, 32768), (1376, 4096, 32768), (512, 11008, 32768), (4096, 11008, 32768)]:
    X = torch.randn(tokens, K, device="cuda", dtype=torch.bfloat16)
    W = torch.randn(N, K, device="cuda") * 0.02
    mask = torch.zeros(K, dtype=torch.uint8, device="cuda"); mask[::100] = 1
    cands = torch.linspace(1, 2, 20).tolist()
    for it in range(3):
        if it == 1:
            _lib.profile_enable(True)
        H = T.hessian_accum(X, 2048, normalize=False)
        T.awq_search_losses(W, H, mask, 4, 128, cands)
    torch.cuda.synchronize()
    print(f"N={N} K={K} T={tokens}")
    for name in ("hessian_gemm", "hessian_reduce", "awq_search_delta", "awq_search_fold", "awq_search_gemm"):
        q = _lib.profile_query(name)
        if q["launches"]:
            ms = q["ms"] / q["launches"]
            extra = f" {q['flops']/q['launches']/ms/1e9:.0f} TF/s" if q["flops"] else f" {q['bytes']/q['launches']/ms/1e6:.0f} GB/s"
            print(f"  {name:18s} {ms*1e3:9.1f} us{extra}")
    _lib.profile_enable(False)
    del X, W, H
