"""Timing of the SmoothQuant alpha-sweep kernel (20 alphas) at Llama-2-7B layer shapes."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import ops, _lib
for N, K, G in [(4096, 4096, 128), (4096, 4096, -1), (4096, 11008, 128)]:
    W = torch.randn(N, K, device="cuda") * 0.02
    act = torch.rand(K, device="cuda") * 4 + 0.1
    S = torch.rand(20, K, device="cuda") + 0.5
    for it in range(3):
        if it == 1:
            _lib.profile_enable(True)
        ops.smooth_alpha_errors(W, S, act, 8, G)
    torch.cuda.synchronize()
    q = _lib.profile_query("smooth_alpha_errors")
    print(f"N={N} K={K} G={G}: {q['ms'] / q['launches']:.3f} ms")
    _lib.profile_enable(False)
