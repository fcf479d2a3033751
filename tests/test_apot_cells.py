"""The cell table of the APOT nearest-level search (csrc/apot_cells.h) against the literal
argmin of pot_apot_quantizer.py:294-297, on the CPU: the header is plain C++, so the very
functions the kernel evaluates (apot_cell_of / apot_cell_entry / apot_lookup) are compiled with
g++ into tests/native/apot_cells_check.cpp and swept over threshold / level / cell-boundary
neighbourhoods and 5 M random points per level set."""
import shutil
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "llm-quantization_b200"
if str(PKG) not in sys.path:
    sys.path.insert(0, str(PKG))

# (n_bit, k) -> must the table be usable?  (6,3) and (4,4) have levels closer than one cell:
# the host keeps the bisecting kernel for them.
LEVEL_SETS = [((4, 2), True), ((3, 1), True), ((8, 2), True), ((4, 1), True), ((2, 1), True),
              ((5, 2), True), ((6, 3), False), ((4, 4), False)]


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cell_table_lookup_is_the_literal_argmin(tmp_path):
    from pot_apot_quantizer import _apot_signed_levels
    blob = tmp_path / "levels.bin"
    with open(blob, "wb") as f:
        f.write(struct.pack("i", len(LEVEL_SETS) + 1))
        for (b, k), _ in LEVEL_SETS:
            lv = _apot_signed_levels(b, k).numpy().astype(np.float32)
            f.write(struct.pack("i", lv.size))
            f.write(lv.tobytes())
        odd = np.array([-3.0, -0.4, 0.1, 0.11, 2.5], np.float32)      # asymmetric, not in [-1, 1]
        f.write(struct.pack("i", odd.size))
        f.write(odd.tobytes())
    exe = tmp_path / "apot_cells_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(PKG / "csrc"),
                    str(REPO / "tests" / "native" / "apot_cells_check.cpp"), "-o", str(exe)], check=True)
    res = subprocess.run([str(exe), str(blob)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [ln.split() for ln in res.stdout.strip().splitlines()]
    assert len(lines) == len(LEVEL_SETS) + 1
    for ((b, k), usable), ln in zip(LEVEL_SETS, lines):
        assert int(ln[5]) == int(usable), (b, k, ln)
        assert int(ln[9]) == 0, (b, k, ln)
        if usable:
            assert int(ln[7]) > 5_000_000
    assert int(lines[-1][5]) == 1 and int(lines[-1][9]) == 0
