"""A/B of the APOT candidate loop on a B200: nearest level through the cell table
(csrc/apot_cells.h, default) against the bisecting kernel (B200Q_APOT_CELLS=0), same process.
Prints ms per call and checks that outputs, level indices, scales and chosen grid points are
bit-identical between the two on full-size matrices; then runs the POT / APOT parity tests.

    python tools/apot_ab.py [log file]
"""
import os
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "llm-quantization_b200"))
LOG = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout


def say(*a):
    print(*a, file=LOG, flush=True)
    if LOG is not sys.stdout:
        print(*a, flush=True)


def main():
    import torch
    from b200q import ops
    from pot_apot_quantizer import _apot_signed_levels
    dev = torch.device("cuda:0")
    ok = True
    for dtype, n, k, bits, kk in ((torch.float32, 8192, 8192, 4, 2), (torch.float16, 4096, 4096, 4, 2),
                                  (torch.bfloat16, 4096, 4096, 4, 2), (torch.float32, 4096, 4096, 8, 2),
                                  (torch.float32, 2048, 4096, 3, 1)):
        g = torch.Generator(device="cpu").manual_seed(n + k + bits)
        w = (torch.randn(n, k, generator=g) * 0.02).to(dtype).to(dev)
        w[0, :5] = 0
        groups = w.reshape(-1, 128)
        levels = _apot_signed_levels(bits, kk)
        grid = torch.arange(0.01, 2.01, 0.1 if w.numel() > 500000 else 0.05)
        res, ms = {}, {}
        for mode in ("0", "1"):
            os.environ["B200Q_APOT_CELLS"] = mode
            for _ in range(2):
                r = ops.apot_quant(groups, levels, grid, return_codes=True)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(5):
                r = ops.apot_quant(groups, levels, grid, return_codes=True)
            t1.record()
            torch.cuda.synchronize()
            ms[mode] = t0.elapsed_time(t1) / 5
            res[mode] = r
        same = all(torch.equal(a, b) for a, b in zip(res["0"], res["1"]))
        ok &= same
        evals = w.numel() * grid.numel()
        say(f"{str(dtype):16s} {n}x{k} w{bits} k{kk}: bisect {ms['0']:.3f} ms, cells {ms['1']:.3f} ms "
            f"({ms['0'] / ms['1']:.2f}x; {evals / ms['1'] / 1e6:.0f} G evals/s), identical: {same}")
    os.environ.pop("B200Q_APOT_CELLS", None)
    say("A/B identical on every case" if ok else "A/B MISMATCH")
    import pytest
    t = time.time()
    rc = pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                      str(REPO / "tests" / "test_gpu_parity_levels.py"),
                      str(REPO / "tests" / "test_packing.py"),
                      str(REPO / "tests" / "test_gpu_baseline_shapes.py") + "::test_pot_apot_row_slices_with_global_numel"])
    say(f"pytest rc={int(rc)} in {time.time() - t:.1f} s")
    return 0 if ok and int(rc) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
