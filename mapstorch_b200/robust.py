"""Objectives of ``fit_spec`` other than the plain sum of squares (opt.py:230-236, 329-396): the
``loss="l1"`` (sum of absolute residuals) objective and the ``l1_lambda`` sparsity penalty on the
amplitudes, for ONE spectrum on the device.

With the calibration parameters fixed the model is linear in the amplitudes A = 10**a >= 0, so

  * MSE + penalty   min |B^T A - r|^2 + c sum(A)      is the non-negative least-squares problem with
    right-hand side  B r - c/2   (the penalty is linear because A >= 0);
  * L1 (+ penalty)  min sum|B^T A - r| + c sum(A)      is solved by iteratively re-weighted least squares:
    weights w = 1 / max(|res|, eps) majorise |res|, every step is a weighted non-negative solve with
    right-hand side  B (w r) - c,  eps is annealed towards 0 (the LAD optimum of an exact LP is what
    the tests compare with).

``c = l1_lambda * loss(bkg, y) / n_elements`` as in the reference (opt.py:372-381).  The E x E systems
are assembled with torch CUDA ops in fp64 and the sign-constrained solve is the same kernel the map
fit uses (``mt_nnls_fixup``); nothing here runs on the host except the loop control.
"""
import torch

from mapstorch_b200 import _native

IRLS_MAX_ITER = 200
IRLS_TOL = 1e-9


def weighted_nnls(bm, tm, weights=None, shift=0.0):
    """argmin_{x>=0} sum_c w_c (x.b_c - t_c)^2 + 2*shift*sum(x).  bm [E, C] fp64 CUDA, tm [C] fp64 CUDA.
    Components whose (weighted) column vanishes get amplitude 0."""
    bw = bm if weights is None else bm * weights[None, :]
    live = (bm * bm).sum(dim=1) > 0
    bl, blw = bm[live], bw[live]
    n_live = int(live.sum())
    x = torch.zeros(bm.shape[0], dtype=torch.float64, device=bm.device)
    if n_live == 0:
        return x
    gram = blw @ bl.T
    gram = 0.5 * (gram + gram.T)
    hinv = torch.linalg.inv(gram)
    hinv = 0.5 * (hinv + hinv.T)
    rhs = blw @ tm - shift
    x_live = (hinv @ rhs).float().reshape(-1, 1).contiguous()
    _native.call("mt_nnls_fixup", x_live.data_ptr(), x_live.stride(0), _native.MT_X_ELEMENT_MAJOR, 1, n_live,
                 gram.contiguous().data_ptr(), hinv.contiguous().data_ptr(), None, None, _native.stream_ptr())
    x[live] = x_live.reshape(-1).double()
    return x


def penalty_constant(tm, loss, l1_lambda, n_elements):
    """c of the module docstring; tm = y - bkg on the fitted channels."""
    if l1_lambda <= 0:
        return 0.0
    scale = float((tm * tm).sum() if loss == "mse" else tm.abs().sum()) / max(int(n_elements), 1)
    return float(l1_lambda) * scale


def objective(bm, tm, x, loss, c):
    res = x @ bm - tm
    main = float((res * res).sum() if loss == "mse" else res.abs().sum())
    return main + c * float(x.sum())


def irls_weights(res, eps):
    return 1.0 / torch.clamp(res.abs(), min=eps)


def solve_amplitudes(bm, tm, loss="mse", c=0.0, trace=None):
    """Amplitudes of one spectrum under the chosen objective; `trace`, if a list, receives the objective value
    after every (re-weighted) solve."""
    if loss == "mse":
        x = weighted_nnls(bm, tm, None, 0.5 * c)
        if trace is not None:
            trace.append(objective(bm, tm, x, loss, c))
        return x
    x = weighted_nnls(bm, tm, None, 0.0)
    top = float(tm.abs().max())
    eps = 1e-3 * max(top, 1e-30)
    floor = 1e-12 * max(top, 1e-30)
    best, best_x = objective(bm, tm, x, loss, c), x
    if trace is not None:
        trace.append(best)
    for _ in range(IRLS_MAX_ITER):
        w = irls_weights(x @ bm - tm, eps)
        x_new = weighted_nnls(bm, tm, w, c)
        value = objective(bm, tm, x_new, loss, c)
        if trace is not None:
            trace.append(value)
        moved = float((x_new - x).abs().max()) / max(float(x_new.abs().max()), 1e-30)
        x = x_new
        if value < best:
            best, best_x = value, x_new
        if moved < IRLS_TOL and eps <= floor:
            break
        eps = max(eps * 0.5, floor)
    return best_x
