"""Launch each hot kernel a few times on one large matrix — the target of `ncu` captures.

    python tools/prof_kernels.py [kernel ...]     kernels: awq gptq smooth uniform pot apot absmax
"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
import torch
from b200q import ops
from pot_apot_quantizer import _apot_signed_levels

which = sys.argv[1:] or ["awq", "gptq", "smooth", "uniform", "pot", "apot", "absmax", "meanabs", "pack"]
torch.manual_seed(0)
N, K = 8192, 8192                      # 268 MB fp32 in + 268 MB out: beyond the 126 MB L2
w = torch.randn(N, K, device="cuda") * 0.02
feats = torch.rand(128, K, device="cuda")
act = torch.rand(K, device="cuda") * 5 + 0.1
small = w[:1024].contiguous()
acts = torch.randn(16, 2048, K, device="cuda", dtype=torch.bfloat16)      # 537 MB of activations
codes = torch.randint(0, 16, (N, K), device="cuda", dtype=torch.uint8)
from b200q import export
for it in range(3):
    if "awq" in which:
        ops.awq_layer(w, feats, 4, 128, K // 100, 2.0)
    if "gptq" in which:
        ops.gptq_parity_layer(w, 4)
    if "smooth" in which:
        ops.smoothquant_layer(w, act, 0.5, 8, 128)
    if "uniform" in which:
        ops.group_fakequant(w, 4, 128)
    if "absmax" in which:
        ops.col_absmax(w)
    if "pot" in which:
        ops.pot_quant(small.view(-1, 128), 4, torch.arange(0.01, 2.01, 0.01))
    if "apot" in which:
        ops.apot_quant(small.view(-1, 128), _apot_signed_levels(4, 2), torch.arange(0.01, 2.01, 0.1))
    if "meanabs" in which:
        ops.act_meanabs_batched(acts)
    if "pack" in which:
        export.pack_codes(codes, 4)
torch.cuda.synchronize()
print("ok")
