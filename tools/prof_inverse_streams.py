"""How much do concurrent Hessian inverses on separate CUDA streams overlap?  Wall time of n
factorisations on n streams vs one after another, at the Llama-2-7B sizes.
    python tools/prof_inverse_streams.py [K ...]"""
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO / "llm-quantization_b200"), str(REPO)):
    sys.path.insert(0, p)
if os.environ.get("CONN"):
    os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = os.environ["CONN"]
import torch
from b200q import tensor_ops as T

Ks = [int(a) for a in sys.argv[1:]] or [4096, 11008]
dev = torch.device("cuda", 0)
for K in Ks:
    g = torch.Generator(device=dev).manual_seed(K)
    X = torch.randn(K + 512, K, device=dev, generator=g)
    H0 = (X.T @ X) / X.shape[0]
    H0 += 0.01 * torch.diag(H0).mean() * torch.eye(K, device=dev)
    del X
    for n in (1, 2, 4, 8, 16):
        streams = [torch.cuda.Stream(dev) for _ in range(n)]
        Hs = [H0.clone() for _ in range(n)]
        outs = [None] * n

        def run():
            main = torch.cuda.current_stream(dev)
            ev = torch.cuda.Event()
            ev.record(main)
            for i, s in enumerate(streams):
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    outs[i] = T.spd_inverse(Hs[i], ridge=1e-6, check=False)
            for s in streams:
                main.wait_stream(s)

        run(); run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            run()
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / reps
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"K={K:6d} streams={n:2d}: {ms:8.2f} ms per round = {ms / n:7.2f} ms per inverse "
              f"(host launch time {host_ms:7.2f} ms per round)", flush=True)
        del Hs, outs
