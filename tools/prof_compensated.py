"""Timing of the error-compensated GPTQ column loop at Llama-2-7B layer shapes."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T, _lib
for N, K in [(4096, 4096), (11008, 4096), (4096, 11008)]:
    X = torch.randn(2 * K, K, device="cuda")
    H = (X.T @ X) / (2 * K) + 0.01 * torch.eye(K, device="cuda")
    U = T.spd_inverse(H, want_inverse=False, want_upper=True)
    W = torch.randn(N, K, device="cuda") * 0.02
    for it in range(3):
        if it == 1:
            _lib.profile_enable(True)
        Q = T.gptq_compensated(W.clone(), U, 4, 128)
    torch.cuda.synchronize()
    q = _lib.profile_query("gptq_compensated")
    print(f"N={N} K={K}: {q['ms'] / q['launches']:.3f} ms per layer")
    _lib.profile_enable(False)
    del X, H, U, W, Q
