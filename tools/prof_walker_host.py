"""Where does the HOST spend its time in the GPTQ model walker?  cProfile of one step over 32
Linears of 4096 x 4096 (16 x 2048-token bf16 calibration activations)."""
import cProfile
import os
import pstats
import sys
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO / "llm-quantization_b200"), str(REPO)):
    sys.path.insert(0, p)
import torch
import torch.nn as nn
import gptq_quantizer as gq

dev = torch.device("cuda", 0)
K = N = 4096
L = 32
g = torch.Generator(device=dev).manual_seed(0)
acts = (torch.randn(16, 2048, K, device=dev, generator=g)).to(torch.bfloat16)
net = nn.Sequential(*[nn.Linear(K, 1, bias=False) for _ in range(L)])
Ws = [torch.randn(N, K, device=dev, generator=g) * 0.02 for _ in range(L)]


def step():
    for lin, w in zip(net, Ws):
        lin.weight = nn.Parameter(w, requires_grad=False)
    gq.gptq_quantize_model_weight(net, 4, 128, {str(i): acts for i in range(L)}, actorder=True, verbose=False)


step(); step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
st.sort_stats("cumulative").print_stats(30)
