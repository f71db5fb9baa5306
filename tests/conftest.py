"""pytest configuration: markers and import paths.

* `gpu` marks tests that need a real B200 (the driver runs `-m gpu` on a GPU box and
  `-m "not gpu"` in the CPU-only build container).
* The product lives in `pangu-pytorch-demo_b200/` which is put on sys.path exactly like the
  reference scripts put their repo root there (finetune/finetune_fully.py:3-5), so
  `from models.pangu_model import PanguModel` resolves to the B200 implementation.
* `oracle/` is test infrastructure and is only importable from tests / bench / smoke.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pangu-pytorch-demo_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: full-resolution CPU oracle run (about a minute)")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens.npz")
    return np.load(path, allow_pickle=False)
