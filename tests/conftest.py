"""pytest configuration: `gpu` marker (tests that need a B200), repo root on sys.path, and the
oracle built on demand.  `-m "not gpu"` must pass on a CPU-only box; `-m gpu` are the parity
tests proper and call the CUDA path through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size configuration (tens of seconds of oracle time)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import core, vecalign_oracle
    core.build()
    return vecalign_oracle


@pytest.fixture(scope="session")
def ocore():
    from oracle import core
    core.build()
    return core


@pytest.fixture(scope="session")
def svb():
    """The package with its CUDA library loaded.  The product path never builds or falls back by itself (a missing
    libsvx.so raises); the TEST harness builds it once when a clean checkout is tested before build() has run."""
    import speech_vecalign_b200 as pkg
    if not os.path.exists(pkg.capi.LIB_PATH):
        pkg.capi.build()
    pkg.capi.lib()
    return pkg


def same_alignments(a, b):
    return [(list(x), list(y)) for x, y in a] == [(list(x), list(y)) for x, y in b]


def ulp_diff(a, b):
    """max |a-b| in units of fp32 ulp at the magnitude of b."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.abs(b).astype(np.float32))))
