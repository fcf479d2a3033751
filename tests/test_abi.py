"""The C-ABI library builds for sm_100a, loads without a GPU, and exports exactly what
include/b200quant.h declares.  CPU only (no compute calls)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from b200q import _lib, build

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "b200quant.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert path.exists()
    lib = _lib.load()
    assert lib.b200q_version() >= 100
    assert lib.b200q_launch_count() >= 0


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(str(build.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in b200quant.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in b200q/_lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_library_is_sm100a_and_has_no_other_arch():
    out = subprocess.run(["cuobjdump", "--list-elf", str(build.LIB_PATH)], capture_output=True,
                         text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_bad_arguments_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    # null pointers / bad shapes return B200Q_EINVAL before any launch
    assert lib.b200q_col_absmax(None, 4, 4, 4, 0, None, 0, None) == -1
    assert b"null" in lib.b200q_last_error()
    assert lib.b200q_group_fakequant(1, 1, None, None, None, 4, 100, 32, 4, 0, 0, None, 0, None) == -1
    assert b"divisible" in lib.b200q_last_error()
    with pytest.raises(AssertionError):
        _lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(-3, "x")
