"""Row-sharded == unsharded on REAL GPUs: tools/dist_check.py under torchrun,
one rank per GPU over NCCL, for every world size in {2, 4, 8} the box offers.  Skipped on a
single-GPU box (the driver's round-end tier); run with `gpurun --gpus N` and keep the logs under
profiles/ (dist_check_<N>gpu.log)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_unsharded_over_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           str(REPO / "tools" / "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=str(REPO))
    out = REPO / "gpurun_out"
    if out.is_dir():
        (out / f"dist_check_{world}gpu.log").write_text(res.stdout + "\n--- stderr tail ---\n" + res.stderr[-4000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "ALL PASS" in res.stdout and "FAIL " not in res.stdout, res.stdout[-3000:]
