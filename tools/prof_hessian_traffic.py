"""ncu target for bench.py's roofline.traffic (profiles/traffic_r2.json): one launch of the Gram
GEMM (AWQ search: plain sum) and one of the GPTQ Hessian GEMM (per-sample weights) at each
in_features of Llama-2-7B, 262144 tokens of bf16 activations, library defaults.
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ... -k regex:hessian_gemm
Launch order in the capture: K=4096 gram, K=4096 hessian, K=11008 gram, K=11008 hessian."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import tensor_ops as T
for K in (4096, 11008):
    X = torch.randn(262144, K, device="cuda", dtype=torch.bfloat16)
    T.hessian_accum(X, 2048, normalize=False)
    T.hessian_accum(X, 2048, normalize=True)
    torch.cuda.synchronize()
    del X
    T.release_workspace()
    torch.cuda.empty_cache()
print("ok")
