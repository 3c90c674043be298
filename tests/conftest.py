"""Shared fixtures.  `-m "not gpu"` covers the oracle, the host logic and the C-ABI exports;
`-m gpu` tests are the parity tests proper and call through the C-ABI on a B200."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def pyoracle():
    from oracle import pyoracle as po
    po.build()
    return po


@pytest.fixture(scope="session")
def model_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("models"))


@pytest.fixture(scope="session")
def model_path(pkg, model_dir):
    """model_path(arch) -> path of the seeded random-init ggml file of that architecture."""
    cache = {}

    def get(arch: str) -> str:
        if arch not in cache:
            p = os.path.join(model_dir, f"ggml-{arch}.bin")
            pkg.ggml_file.make_model(p, arch)
            cache[arch] = p
        return cache[arch]

    return get


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "micro_golden.npz"))


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def mel_close(x, ref, tol=1e-4):
    """mel tolerance of BASELINE.json's north_star ("within 1e-4 relative"), stated as in
    SURVEY.md section 7: |d| <= tol * max(1, |ref|) element-wise AND rel-L2 <= tol (a pure
    element-wise relative test is ill-posed: normalised values pass through 0)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return bool(np.all(np.abs(x - ref) <= tol * np.maximum(1.0, np.abs(ref)))) and rel_l2(x, ref) <= tol
