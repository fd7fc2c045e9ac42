import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("PYTHONHASHSEED", "0")
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    return np.load(GOLDEN / name, allow_pickle=False)


@pytest.fixture(scope="session")
def model_golden():
    return _load("model_golden.npz")


@pytest.fixture(scope="session")
def snip_golden():
    return _load("snip_golden.npz")


@pytest.fixture(scope="session")
def host_golden():
    return _load("host_golden.npz")


@pytest.fixture(scope="session")
def fit_golden():
    return _load("fit_golden.npz")


@pytest.fixture(scope="session")
def override_golden():
    return _load("override_golden.npz")


def golden_params(g, tag):
    return dict(zip([str(k) for k in g[f"{tag}_param_names"]], [float(v) for v in g[f"{tag}_param_vals"]]))
