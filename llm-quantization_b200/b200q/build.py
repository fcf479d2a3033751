"""Build libb200quant.so in-tree with nvcc for sm_100a.

The library is plain CUDA C++ with a C ABI (include/b200quant.h); it links only libcudart, so it
cross-compiles on a machine without a GPU and travels to the B200 box as a prebuilt file.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent          # llm-quantization_b200/
CSRC = PKG_DIR / "csrc"
REPO = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libb200quant.so"
OBJ_DIR = CSRC / "build"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]

# per-file extra flags.  The elementwise / level-search kernels must round every op separately to
# stay bit-identical with torch (no FMA contraction); the GEMM-shaped kernels want FMAs.
SOURCES = {
    "core.cu": [],
    "elementwise.cu": ["-fmad=false"],
    "levels.cu": ["-fmad=false"],
    "tensorcore.cu": [],
    "linalg.cu": [],
    "splitgemm.cu": [],
    "packing.cu": [],
    "qgemm.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200quant.so cannot be built (no CPU fallback exists)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libb200quant.so. Returns its path."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [REPO / "include" / "b200quant.h"]
    objs = []
    log_lines = []
    jobs = []
    for name, extra in SOURCES.items():
        src = CSRC / name
        if not src.exists():
            continue
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            jobs.append((name, [nvcc, *ARCH, *COMMON, *extra, "-I", str(REPO / "include"), "-c",
                                str(src), "-o", str(obj)]))
    # the translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda j: subprocess.run(j[1], capture_output=True, text=True), jobs))
    for (name, cmd), res in zip(jobs, results):
        log_lines.append("$ " + " ".join(cmd))
        log_lines.append(res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {name}")
        if verbose:
            print(res.stderr)
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libb200quant.so failed")
    if log_lines:
        (OBJ_DIR / "ptxas.log").write_text("\n".join(log_lines))
    return LIB_PATH


REFERENCE_SRC = Path("/root/reference")
REFERENCE_STAGE = REPO / "baseline" / "_ref"


def stage_reference() -> Path | None:
    """Copy the UNMODIFIED reference modules (flat *.py + config.json) from /root/reference into the
    git-ignored baseline/_ref/, so that they travel to the GPU box with the snapshot.  Used there
    only (a) to run the reference's own test_quantization.py against the drop-in modules and (b) as
    the timed CPU implementation of `bench.py --impl reference`.  Nothing in the product imports it
    except the documented out-of-scope pass-throughs of quantization_utils (HF load / perplexity).
    Returns the staged directory, or None when no reference checkout is present (GPU box: the
    files staged earlier are used as they are)."""
    if not REFERENCE_SRC.is_dir():
        return REFERENCE_STAGE if (REFERENCE_STAGE / "test_quantization.py").exists() else None
    REFERENCE_STAGE.mkdir(parents=True, exist_ok=True)
    for src in sorted(REFERENCE_SRC.glob("*.py")) + [REFERENCE_SRC / "config.json"]:
        if src.exists():
            dst = REFERENCE_STAGE / src.name
            if not dst.exists() or dst.read_bytes() != src.read_bytes():
                shutil.copyfile(src, dst)
    return REFERENCE_STAGE


def reference_dir() -> Path | None:
    """Where an unmodified reference checkout can be imported from: $LLMQ_REFERENCE_DIR when set
    (authoritative), else /root/reference, else the staged baseline/_ref."""
    env = os.environ.get("LLMQ_REFERENCE_DIR")
    for cand in ((env,) if env else (REFERENCE_SRC, REFERENCE_STAGE)):
        if cand and (Path(cand) / "quantization_utils.py").exists():
            return Path(cand)
    return None


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
