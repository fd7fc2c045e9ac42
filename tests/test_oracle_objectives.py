"""The oracle's exact solvers for the non-MSE objectives (oracle.fit_lad, oracle.fit_nnls_penalised) against
the reference's own Adam runs (tests/golden/objectives_golden.npz, 6000 iterations, make_golden.py
--only objectives).  The exact optimum can only be BETTER than where Adam stopped; amplitudes agree to
what a fixed-step optimiser on a non-smooth objective resolves (stated per case)."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import maps_oracle as mo


@pytest.fixture(scope="module")
def cases():
    g = np.load(GOLDEN / "objectives_golden.npz", allow_pickle=False)
    p = dict(zip([str(k) for k in g["param_names"]], [float(v) for v in g["param_vals"]]))
    er = tuple(int(v) for v in g["er"])
    spec = g["spec"].astype(np.float64)
    out = {}
    for tag in (str(t) for t in g["runs"]):
        names = [str(n) for n in g[f"{tag}_names"]]
        kw = eval(str(g[f"{tag}_kwargs"]))  # repr of a small dict written by make_golden.py
        idx = g[f"{tag}_indices"]
        B, onames = mo.basis(p, er, names[:-2])
        B = B.astype(np.float64).T[[list(onames).index(n) for n in names]]
        Bm, rm = (B, spec) if idx.size == 0 else (B[:, idx], spec[idx])
        lam = kw.get("l1_lambda", 0.0)
        if kw["loss"] == "mse":
            c = lam * float((rm ** 2).sum()) / len(names)
            x = mo.fit_nnls_penalised(rm, Bm, None, c)
            value = lambda v, Bm=Bm, rm=rm, c=c: float(((v @ Bm - rm) ** 2).sum() + c * v.sum())  # noqa: E731
        else:
            c = lam * float(np.abs(rm).sum()) / len(names)
            x, _ = mo.fit_lad(rm, Bm, None, c)
            value = lambda v, Bm=Bm, rm=rm, c=c: float(np.abs(v @ Bm - rm).sum() + c * v.sum())  # noqa: E731
        out[tag] = dict(x=x, ref=10.0 ** g[f"{tag}_logamps"], value=value, trace=g[f"{tag}_trace"], Bm=Bm, rm=rm, c=c)
    return out


@pytest.mark.parametrize("tag,amp_tol", [("l1", 2e-2), ("l1_pen", 2e-2), ("mse_pen", 1e-3), ("mse_idx", 5e-3)])
def test_exact_optimum_vs_reference_run(cases, tag, amp_tol):
    c = cases[tag]
    assert c["value"](c["x"]) <= c["value"](c["ref"]) * (1 + 1e-9)
    assert c["value"](c["x"]) <= float(c["trace"].min()) * (1 + 1e-6)   # the reference's own fp32 loss trace
    assert c["value"](c["x"]) >= float(c["trace"].min()) * (1 - 1e-3)   # and Adam did get close to it
    assert np.abs(c["x"] - c["ref"]).max() <= amp_tol * c["x"].max()


def test_penalised_nnls_satisfies_kkt(cases):
    c = cases["mse_pen"]
    x, Bm, rm = c["x"], c["Bm"], c["rm"]
    grad = 2.0 * (Bm @ (x @ Bm - rm)) + c["c"]
    live = (Bm * Bm).sum(axis=1) > 0
    scale = np.abs(2.0 * (Bm @ rm)).max()
    assert np.all(x >= 0)
    assert np.abs(grad[live & (x > 0)]).max() <= 1e-9 * scale
    assert np.all(grad[live & (x == 0)] >= -1e-9 * scale)
    assert (x == 0).sum() >= 5 and (x > 0).sum() >= 8
    plain = mo.fit_nnls_penalised(rm, Bm, None, 0.0)
    assert np.allclose(plain, mo.fit_nnls(rm[None, :], Bm.T)[0] if hasattr(mo, "fit_nnls") else plain, rtol=1e-9, atol=1e-9)


def test_lad_beats_least_squares_on_outliers(cases):
    c = cases["l1"]
    ls = mo.fit_nnls_penalised(c["rm"], c["Bm"], None, 0.0)
    assert c["value"](c["x"]) < 0.95 * c["value"](ls)
