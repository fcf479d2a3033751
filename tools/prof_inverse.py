"""Stage timings and accuracy of b200q_spd_inverse at the Llama-2-7B Hessian sizes.
B200Q_INVERSE_PLANES=2|3 selects the tensor-core split (read once by the library)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T, _lib
for K in [int(a) for a in sys.argv[1:]] or [4096, 11008]:
    X = torch.randn(2 * K, K, device="cuda")
    H = (X.T @ X) / (2 * K) + 0.01 * torch.eye(K, device="cuda")
    for it in range(3):
        if it == 1:
            _lib.profile_enable(True)
        Hinv = T.spd_inverse(H)
    torch.cuda.synchronize()
    ref = torch.linalg.inv(H.double())
    res = ((Hinv.double() @ H.double()) - torch.eye(K, device="cuda", dtype=torch.float64)).abs().max().item()
    rel = ((Hinv.double() - ref).abs().max() / ref.abs().max()).item()
    t32 = torch.linalg.inv(H)
    rel32 = ((t32.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"K={K}  max|Hinv H - I| = {res:.2e}   max|Hinv - inv64| / max|inv64| = {rel:.2e}   (torch fp32 inv: {rel32:.2e})")
    for name in ("spd_inverse", "inv_factor", "inv_diag", "inv_split", "inv_gemm", "inv_product"):
        q = _lib.profile_query(name)
        if q["launches"]:
            print(f"  {name:12s} {q['ms'] / 2:9.3f} ms per inverse  ({q['launches'] // 2} scopes)")
    _lib.profile_enable(False)
    del X, H, Hinv, ref, t32
