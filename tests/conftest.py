"""pytest configuration: the `gpu` marker, import paths, golden-fixture access."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "llm-quantization_b200"
for p in (str(PKG), str(REPO)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Access to tests/golden/<family>.npz with keys 'family/case/array'."""

    def __init__(self, family: str):
        self.family = family
        self.z = np.load(GOLDEN / f"{family}.npz")

    def cases(self, prefix: str):
        seen = []
        for k in self.z.files:
            parts = k.split("/")
            if parts[0] == prefix and len(parts) >= 3 and parts[1] not in seen:
                seen.append(parts[1])
        return seen

    def has(self, key: str) -> bool:
        return key in self.z.files

    def arr(self, key: str) -> np.ndarray:
        return self.z[key]

    def tensor(self, key: str, dtype: torch.dtype = None) -> torch.Tensor:
        a = self.z[key]
        if a.dtype == np.int16 and dtype == torch.bfloat16:
            return torch.from_numpy(a.copy()).view(torch.bfloat16)
        t = torch.from_numpy(a.copy())
        return t if dtype is None else t.to(dtype)


def case_dtype(name: str) -> torch.dtype:
    return {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[name.split("_")[0]]


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(family: str) -> Golden:
        if family not in cache:
            cache[family] = Golden(family)
        return cache[family]

    return get
