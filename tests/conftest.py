import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FILES = ["mtrand32", "mtrand32_new1", "mtrand32_new", "mtrand64", "matrix"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def inputs():
    """Token streams of the reference's text matrices (tests/golden/make_fixtures.py)."""
    return np.load(os.path.join(GOLDEN, "inputs.npz"))


@pytest.fixture(scope="session")
def golden():
    """Answers produced by the reference's own pivotedA / verifyInv + oracle outputs
    (tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "golden_cpu.npz"))


def template(inputs, name, n, dtype=np.float32):
    """N x N template exactly as the reference's main() reads it: a PREFIX of the token
    stream (templated/luBatchedInplace.cu:31-34)."""
    suf = "_f32" if np.dtype(dtype) == np.float32 else "_f64"
    return np.ascontiguousarray(inputs[name + suf][: n * n].reshape(n, n).astype(dtype))


def synthetic(n, batch, dtype, seed=None, dominant=False):
    """SURVEY.md section 8(d) extra sets: distinct uniform(0,1) matrices, optionally + N*I."""
    rng = np.random.default_rng(1000 * n + (0 if np.dtype(dtype) == np.float32 else 1) if seed is None else seed)
    A = rng.uniform(0.0, 1.0, size=(batch, n, n)).astype(dtype)
    if dominant:
        A += n * np.eye(n, dtype=dtype)
    return A
