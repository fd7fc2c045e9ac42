"""mt_spec_solve (csrc/mt_spec.cu): the one-launch amplitude solve inside fit_spec(tune_params=True).

Oracle: the exact non-negative least-squares solution of the same problem on the same fp32 basis in float64
(oracle.fit_nnls_penalised = Lawson-Hanson on the Cholesky factor; weighted problems by scaling rows with sqrt(w)),
plus the torch-op route the kernel replaced (robust.weighted_nnls).  Tolerance: 1e-9 of the largest amplitude --
both sides work in float64 on identical inputs; what differs is the summation order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-9


def make_problem(n_comp, n_ch, seed, dead=(), negative=True):
    """Overlapping Gaussian columns + a spectrum some of whose unconstrained amplitudes are negative."""
    rng = np.random.default_rng(seed)
    ch = np.arange(n_ch, dtype=np.float64)
    centres = np.sort(rng.uniform(0.05 * n_ch, 0.95 * n_ch, n_comp))
    widths = rng.uniform(4.0, 25.0, n_comp)
    basis = np.exp(-0.5 * ((ch[None, :] - centres[:, None]) / widths[:, None]) ** 2)
    basis += 0.3 * np.exp(-0.5 * ((ch[None, :] - (centres[:, None] + 40.0)) / widths[:, None]) ** 2)
    for e in dead:
        basis[e] = 0.0
    amps = rng.uniform(0.0, 500.0, n_comp)
    amps[rng.random(n_comp) < 0.3] = 0.0
    y = amps @ basis + 20.0 * np.exp(-ch / (0.3 * n_ch))
    y = rng.poisson(np.clip(y, 0, None)).astype(np.float64)
    if negative:
        y -= 0.5 * basis[n_comp // 2] * (amps[n_comp // 2] + 50.0)  # drives that component below zero
    bkg = 15.0 * np.exp(-ch / (0.25 * n_ch))
    return basis.astype(np.float32), y.astype(np.float32), bkg.astype(np.float32)


def exact(basis, y, bkg, idx, w, l1_lambda, robust, n_elements):
    from oracle import maps_oracle as mo

    t = (y - (bkg if bkg is not None else np.float32(0))).astype(np.float64)  # fp32 difference, like the product
    bm = basis.astype(np.float64)
    if idx is not None:
        bm, t = bm[:, idx], t[idx]
    c = 0.0
    if l1_lambda > 0:
        c = l1_lambda * float(np.abs(t).sum() if robust else (t * t).sum()) / max(n_elements, 1)
    shift = c if robust else 0.5 * c
    sw = np.ones_like(t) if w is None else np.sqrt(w)
    x = mo.fit_nnls_penalised(t * sw, bm * sw[None, :], None, 2.0 * shift)
    res = x @ bm - t
    vec = sw * res
    pk = 2.0 if robust else 1.0
    value = float(vec @ vec + pk * c * x.sum())
    objective = float((np.abs(res).sum() if robust else (res * res).sum()) + c * x.sum())
    return x, res, vec, value, objective, c


def run(basis, y, bkg=None, idx=None, w=None, l1_lambda=0.0, robust=False, n_elements=1):
    from mapstorch_b200 import robust as rb

    dev = lambda a, dt: None if a is None else torch.as_tensor(a, dtype=dt, device="cuda")  # noqa: E731
    x, res, vec, stats = rb.spec_solve(dev(basis, torch.float32), dev(y, torch.float32), dev(bkg, torch.float32),
                                       dev(idx, torch.int32), dev(w, torch.float64), l1_lambda, robust, n_elements)
    torch.cuda.synchronize()
    return x.cpu().numpy(), res.cpu().numpy(), vec.cpu().numpy(), stats.cpu().numpy()


def check(got, want, n_k, penalty):
    x, res, vec, stats = (g[0] for g in got)
    ex, eres, evec, value, objective, c = want
    scale = max(np.abs(ex).max(), 1e-30)
    assert stats[3] == 0.0, "solver flagged a singular block"
    np.testing.assert_allclose(x, ex, rtol=0, atol=REL * scale)
    rscale = max(np.abs(eres).max(), 1e-30)
    np.testing.assert_allclose(res, eres, rtol=0, atol=1e-7 * rscale)  # residual of 1e-9-accurate amplitudes
    np.testing.assert_allclose(vec[:n_k], evec, rtol=0, atol=1e-7 * max(np.abs(evec).max(), 1e-30))
    pk = vec[n_k:]
    if penalty:
        assert np.all(pk >= 0) and pk.max() > 0
    else:
        assert np.all(pk == 0)
    np.testing.assert_allclose(stats[0], value, rtol=1e-9)
    np.testing.assert_allclose(stats[1], objective, rtol=1e-9)
    np.testing.assert_allclose(stats[2], c, rtol=1e-12, atol=0)
    assert (x >= 0).all()


@pytest.mark.parametrize("n_comp,n_ch", [(12, 1401), (1, 257), (5, 64), (33, 2048), (93, 2048), (96, 4096)])
def test_plain_and_background(n_comp, n_ch):
    basis, y, bkg = make_problem(n_comp, n_ch, seed=n_comp)
    check(run(basis, y), exact(basis, y, None, None, None, 0.0, False, 1), n_ch, False)
    check(run(basis, y, bkg), exact(basis, y, bkg, None, None, 0.0, False, 1), n_ch, False)
    # some amplitude is actually clamped in these problems (otherwise the pivoting is not exercised)
    assert (exact(basis, y, bkg, None, None, 0.0, False, 1)[0] == 0).any() or n_comp < 12


def test_dead_components_and_an_all_dead_basis():
    basis, y, bkg = make_problem(12, 800, seed=3, dead=(0, 7))
    got = run(basis, y, bkg)
    check(got, exact(basis, y, bkg, None, None, 0.0, False, 1), 800, False)
    assert got[0][0][0] == 0.0 and got[0][0][7] == 0.0
    got = run(np.zeros_like(basis), y, bkg)
    assert (got[0] == 0).all() and got[3][0][3] == 0.0
    np.testing.assert_allclose(got[1][0], -(y - bkg).astype(np.float64))


def test_channel_subsets_weights_and_penalties():
    basis, y, bkg = make_problem(14, 1200, seed=5)
    rng = np.random.default_rng(1)
    idx = np.sort(rng.choice(1200, size=700, replace=False)).astype(np.int32)
    w = 1.0 / np.clip(np.abs(rng.standard_normal(700)) * 30.0, 1e-3, None)
    check(run(basis, y, bkg, idx), exact(basis, y, bkg, idx, None, 0.0, False, 1), 700, False)
    check(run(basis, y, bkg, idx, w), exact(basis, y, bkg, idx, w, 0.0, False, 1), 700, False)
    check(run(basis, y, bkg, None, None, 0.02, False, 14), exact(basis, y, bkg, None, None, 0.02, False, 14), 1200, True)
    check(run(basis, y, bkg, idx, w, 0.05, True, 14), exact(basis, y, bkg, idx, w, 0.05, True, 14), 700, True)


def test_batches_match_single_problems():
    from mapstorch_b200 import robust as rb

    probs = [make_problem(12, 1401, seed=20 + k) for k in range(5)]
    basis = torch.as_tensor(np.stack([p[0] for p in probs]), device="cuda")
    bkg = torch.as_tensor(np.stack([p[2] for p in probs]), device="cuda")
    y = torch.as_tensor(probs[0][1], device="cuda")
    w = torch.rand((5, 1401), dtype=torch.float64, device="cuda") + 0.1
    xb, rb_, vb, sb = rb.spec_solve(basis, y, bkg, None, w)
    for k in range(5):
        x1, r1, v1, s1 = rb.spec_solve(basis[k], y, bkg[k], None, w[k])
        assert torch.equal(x1[0], xb[k]) and torch.equal(r1[0], rb_[k]) and torch.equal(v1[0], vb[k])
        assert torch.equal(s1[0], sb[k])
    # one shared background / weight row for the whole batch
    xs, _, _, _ = rb.spec_solve(basis, y, bkg[0], None, w[0])
    x0, _, _, _ = rb.spec_solve(basis[3], y, bkg[0], None, w[0])
    assert torch.equal(xs[3], x0[0])


def test_agrees_with_the_torch_route_it_replaced():
    from mapstorch_b200 import robust as rb

    basis, y, bkg = make_problem(12, 1401, seed=9)
    b, yy, bb = (torch.as_tensor(a, device="cuda") for a in (basis, y, bkg))
    x, res, vec, stats = rb.spec_solve(b, yy, bb)
    tm = (yy - bb).double()
    old = rb.weighted_nnls(b.double(), tm)  # amplitudes rounded to fp32 on the way
    np.testing.assert_allclose(x[0].cpu().numpy(), old.cpu().numpy(), rtol=0, atol=2e-7 * float(old.max()))
    assert abs(float(stats[0, 1]) - rb.objective(b.double(), tm, old, "mse", 0.0)) <= 1e-7 * float(stats[0, 1])


def test_singular_gram_matrix_is_reported_not_returned():
    basis, y, bkg = make_problem(6, 500, seed=2)
    basis[4] = basis[1]  # identical columns
    got = run(basis, y, bkg)
    assert got[3][0][3] == 2.0 and (got[0] == 0).all()
    from mapstorch_b200 import _native

    with pytest.raises(_native.NativeError):
        run(np.zeros((97, 64), np.float32), np.zeros(64, np.float32))


def test_warm_started_pivoting_gives_the_same_answer_in_fewer_steps():
    """free_set (in / out): the partition a neighbouring trial point ended with starts the next solve.  Whatever is
    handed in -- the right partition, an empty one, another problem's, garbage -- the solution is the same to the bit
    (the final partition, hence the final factorisation, is the same) and the right one needs a single pivot step."""
    from mapstorch_b200 import robust as rb

    basis, y, bkg = make_problem(40, 2048, seed=12)
    b, yy, bb = (torch.as_tensor(a, device="cuda") for a in (basis, y, bkg))
    cold_set = torch.full((1, 40), 255, dtype=torch.uint8, device="cuda")
    x0, r0, v0, s0 = rb.spec_solve(b, yy, bb, free_set=cold_set)
    ref = rb.spec_solve(b, yy, bb)
    assert torch.equal(x0, ref[0]) and int(s0[0, 4]) == int(ref[3][0, 4]) and int(s0[0, 4]) >= 2
    assert set(cold_set.unique().tolist()) <= {0, 1, 2} and int((cold_set == 0).sum()) >= 1
    assert torch.equal((cold_set[0] == 1), x0[0] > 0) or int(((cold_set[0] == 1) != (x0[0] > 0)).sum()) <= 1
    warm_set = cold_set.clone()
    x1, _, _, s1 = rb.spec_solve(b, yy, bb, free_set=warm_set)
    assert torch.equal(x1, x0) and int(s1[0, 4]) == 1 and torch.equal(warm_set, cold_set)
    other = make_problem(40, 2048, seed=13)
    wrong = torch.full((1, 40), 255, dtype=torch.uint8, device="cuda")
    rb.spec_solve(torch.as_tensor(other[0], device="cuda"), torch.as_tensor(other[1], device="cuda"), None, free_set=wrong)
    for start in (torch.zeros((1, 40), dtype=torch.uint8, device="cuda"), wrong,
                  torch.randint(0, 2, (1, 40), dtype=torch.uint8, device="cuda")):
        xs, _, _, ss = rb.spec_solve(b, yy, bb, free_set=start)
        assert float(ss[0, 3]) == 0.0
        np.testing.assert_allclose(xs.cpu().numpy(), x0.cpu().numpy(), rtol=0, atol=1e-9 * float(x0.max()))
        assert torch.equal(start, cold_set)
