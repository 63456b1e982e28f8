"""Shared test plumbing.

`-m "not gpu"` : oracle vs the golden fixtures generated from the reference, host logic, C-ABI load.
`-m gpu`       : parity of the CUDA path (through the C ABI) against the oracle and the fixtures.
Only tests may import oracle/; the product package (deep-neural-net_b200/mlpcode) never does.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
PKG_ROOT = REPO / "deep-neural-net_b200"
for p in (str(PKG_ROOT), str(REPO / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): build it once if a fresh checkout lacks it."""
    lib = PKG_ROOT / "lib" / "libbnn_b200.so"
    if not lib.exists():
        import subprocess
        subprocess.check_call(["make", "-C", str(PKG_ROOT / "csrc"), "-j", "8"])


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with np.load(GOLDEN / f"{name}.npz") as f:
            return {k: f[k] for k in f.files}
    return load


@pytest.fixture(scope="session")
def oracle():
    import bnn_oracle
    bnn_oracle.build_c()
    return bnn_oracle


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (b = reference)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def params_from_golden(g, prefix, n_layers):
    return dict(W=[g[f"{prefix}W{i}"].copy() for i in range(n_layers)],
                bn=[{k: g[f"{prefix}{k}{i}"].copy() for k in ("gamma", "beta", "mu", "sigma")}
                    for i in range(n_layers)])
