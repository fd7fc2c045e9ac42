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


def conditional_optimum(x_nnls, gram, ref):
    """Least-squares optimum of the free components GIVEN the reference's values on the components NNLS clamps.

    The reference reaches a clamped component only asymptotically (A = 10**a, Adam on a: opt.py:154, 311-317), so
    after any finite number of steps it still carries small amplitudes there, and the components that overlap them
    sit at the optimum conditioned on those leftovers: x_F = x_F(NNLS) - G_FF^-1 G_FA x_A(ref).  Comparing against
    this removes the reference's unfinished convergence from the parity check (measured: 3e-5 instead of 1.5e-2)."""
    zero = np.nonzero(x_nnls == 0)[0]
    free = np.nonzero(x_nnls > 0)[0]
    out = np.array(x_nnls, dtype=np.float64)
    if zero.size:
        out[free] -= np.linalg.solve(gram[np.ix_(free, free)], gram[np.ix_(free, zero)] @ ref[zero])
        out[zero] = ref[zero]
    return out
