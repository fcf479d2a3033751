"""The reference's OWN test file (test_quantization.py:458-503, 13 tests) executed UNCHANGED
against the drop-in modules on the GPU.

`__graft_entry__.build()` stages the unmodified reference files into the git-ignored baseline/_ref/
(they travel to the GPU box with the snapshot).  The file is run in a subprocess through runpy with
llm-quantization_b200/ FIRST on sys.path, so `import awq_quantizer` & co. resolve to the B200
modules, while `benchmark_runner` (out of scope, imported by test_framework_imports) and
`config.json` (read from the CWD by test_config_loading) come from the staged reference."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "llm-quantization_b200"
REF = REPO / "baseline" / "_ref"

RUNNER = r"""
import runpy, sys
pkg, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [pkg, ref]
import awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer, quantization_utils
for m in (awq_quantizer, gptq_quantizer, pot_apot_quantizer, smooth_quant_quantizer, quantization_utils):
    assert m.__file__.startswith(pkg), (m.__name__, m.__file__)
from b200q import _lib
n0 = _lib.launch_count()
try:
    runpy.run_path(ref + "/test_quantization.py", run_name="__main__")
except SystemExit as e:
    code = e.code
else:
    code = 0
print(f"B200Q_LAUNCHES {_lib.launch_count() - n0}")
import benchmark_runner
assert benchmark_runner.__file__.startswith(ref)
assert benchmark_runner.awq_quantize_model_weight is awq_quantizer.awq_quantize_model_weight
assert benchmark_runner.gptq_quantize_model_weight is gptq_quantizer.gptq_quantize_model_weight
sys.exit(code or 0)
"""


def test_reference_test_quantization_runs_unchanged_on_the_drop_in():
    if not (REF / "test_quantization.py").exists():
        pytest.skip("baseline/_ref/test_quantization.py missing: run __graft_entry__.build() where "
                    "/root/reference exists (the staged files ship with the gpurun snapshot); the last "
                    "recorded run is profiles/r2_reference_test_quantization.log")
    res = subprocess.run([sys.executable, "-c", RUNNER, str(PKG), str(REF)], cwd=str(REF),
                         capture_output=True, text=True, timeout=900)
    tail = (res.stdout[-3000:] + "\n" + res.stderr[-3000:])
    assert res.returncode == 0, tail
    assert "TEST RESULTS: 13 passed, 0 failed" in res.stdout, tail
    launches = [int(l.split()[1]) for l in res.stdout.splitlines() if l.startswith("B200Q_LAUNCHES")]
    assert launches and launches[0] > 20, f"the drop-in's kernels did not run: {launches}"
    out = REPO / "gpurun_out"
    if out.is_dir():
        (out / "reference_test_quantization.log").write_text(res.stdout + "\n--- stderr ---\n" + res.stderr)
