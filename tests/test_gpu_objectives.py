"""The objectives of fit_spec beyond plain least squares: loss="l1", the l1_lambda
amplitude penalty and channel subsets, on the GPU.

Oracles: exact solvers of the same convex problems (oracle.fit_lad = linear programme,
oracle.fit_nnls_penalised = Lawson-Hanson on the shifted system) and the reference's own Adam runs in
tests/golden/objectives_golden.npz (6000 iterations).  Adam with a fixed step in log10(A) does not
settle exactly on a non-smooth optimum, so against the reference the bar is "our objective value is not
worse, amplitudes agree to the tolerance stated per test"; against the exact solvers it is tight."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    from conftest import GOLDEN

    g = np.load(GOLDEN / "objectives_golden.npz", allow_pickle=False)
    params = dict(zip([str(k) for k in g["param_names"]], [float(v) for v in g["param_vals"]]))
    return g, params


def run_case(g, params, tag):
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    names = [str(n) for n in g[f"{tag}_names"]]
    kw = eval(str(g[f"{tag}_kwargs"]))  # repr of a small dict written by make_golden.py
    idx = g[f"{tag}_indices"]
    er = tuple(int(v) for v in g["er"])
    spec = g["spec"]
    tensors, spec_fit, bkg, trace = opt.fit_spec(
        spec, er, elements_to_fit=names[:-2], init_param_vals=params, tune_params=False, use_snip=False,
        indices=None if idx.size == 0 else idx, progress_bar=False, **kw)
    mine = np.array([10.0 ** tensors[n].item() for n in names])
    ref = 10.0 ** g[f"{tag}_logamps"]
    B, onames = mo.basis(params, er, names[:-2])
    B = B.astype(np.float64).T  # [E, C]
    order = [list(onames).index(n) for n in names]
    B = B[order]
    r = spec.astype(np.float64)
    Bm, rm = (B, r) if idx.size == 0 else (B[:, idx], r[idx])
    lam = kw.get("l1_lambda", 0.0)
    if kw["loss"] == "mse":
        c = lam * float((rm ** 2).sum()) / len(names)
        exact = mo.fit_nnls_penalised(rm, Bm, None, c)
        value = lambda v: float(((v @ Bm - rm) ** 2).sum() + c * v.sum())  # noqa: E731
    else:
        c = lam * float(np.abs(rm).sum()) / len(names)
        exact, _ = mo.fit_lad(rm, Bm, None, c)
        value = lambda v: float(np.abs(v @ Bm - rm).sum() + c * v.sum())  # noqa: E731
    return dict(mine=mine, ref=ref, exact=exact, value=value, trace=trace, ref_trace=g[f"{tag}_trace"], c=c,
                spec_fit=spec_fit, bkg=bkg)


def test_l1_loss_reaches_the_lad_optimum(golden):
    g, params = golden
    r = run_case(g, params, "l1")
    best = r["value"](r["exact"])
    assert r["value"](r["mine"]) <= best * (1 + 1e-5)          # the exact LP optimum
    assert r["value"](r["mine"]) <= float(r["ref_trace"].min())  # never worse than the reference's Adam run
    assert abs(r["trace"][-1] - r["value"](r["mine"])) <= 1e-4 * best  # loss_trace is in the reference's units
    top = r["exact"].max()
    assert np.abs(r["mine"] - r["exact"]).max() <= 1e-2 * top  # LAD optima are flat; amplitudes to 1 % of max
    assert np.abs(r["mine"] - r["ref"]).max() <= 2e-2 * top


def _mse_solution(g, params):
    from mapstorch_b200 import opt

    names = [str(n) for n in g["l1_names"]]
    t, _, _, _ = opt.fit_spec(g["spec"], tuple(int(v) for v in g["er"]), elements_to_fit=names[:-2],
                              init_param_vals=params, tune_params=False, use_snip=False, progress_bar=False)
    return np.array([10.0 ** t[n].item() for n in names])


def test_l1_and_mse_answers_differ_on_outliers(golden):
    g, params = golden
    r = run_case(g, params, "l1")
    mse = _mse_solution(g, params)
    assert r["value"](r["mine"]) < 0.95 * r["value"](mse)


def test_mse_with_amplitude_penalty(golden):
    g, params = golden
    r = run_case(g, params, "mse_pen")
    exact, mine, ref = r["exact"], r["mine"], r["ref"]
    top = exact.max()
    interior = exact > 1e-3 * top
    assert interior.sum() >= 8 and (~interior).sum() >= 5        # the penalty switched components off
    assert np.abs(mine[interior] / exact[interior] - 1).max() <= 1e-4
    assert np.abs(mine[~interior] - exact[~interior]).max() <= 1e-4 * top
    assert r["value"](mine) <= float(r["ref_trace"][-1]) * (1 + 1e-6)
    # the reference's converged run: surviving components agree, switched-off ones are small there too
    assert np.abs(mine[interior] / ref[interior] - 1).max() <= 1e-3
    assert np.abs(ref[~interior] - exact[~interior]).max() <= 5e-3 * top


def test_l1_with_amplitude_penalty(golden):
    g, params = golden
    r = run_case(g, params, "l1_pen")
    best = r["value"](r["exact"])
    assert r["value"](r["mine"]) <= best * (1 + 1e-5)
    assert r["value"](r["mine"]) <= float(r["ref_trace"].min())
    assert np.abs(r["mine"] - r["exact"]).max() <= 1e-2 * r["exact"].max()
    assert (r["mine"] <= 1e-6 * r["exact"].max()).sum() >= 5


def test_channel_subset(golden):
    g, params = golden
    r = run_case(g, params, "mse_idx")
    exact, mine, ref = r["exact"], r["mine"], r["ref"]
    top = exact.max()
    interior = exact > 1e-3 * top
    assert np.abs(mine[interior] / exact[interior] - 1).max() <= 1e-4
    assert np.abs(mine[interior] / ref[interior] - 1).max() <= 1e-3
    assert np.abs(mine[~interior] - exact[~interior]).max() <= 1e-4 * top


@pytest.mark.parametrize("dtype,use_snip,with_idx", [(torch.float32, False, False), (torch.float16, False, True),
                                                     (torch.float32, True, False)])
def test_fit_map_penalty_per_pixel(dtype, use_snip, with_idx):
    """fit_map(l1_lambda=...) == the penalised problem solved exactly pixel by pixel (c depends on the pixel)."""
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_param_vals
    from oracle import maps_oracle as mo

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    fitted = "Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As".split()
    present = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
    er = (0, 2047)
    rng = np.random.default_rng(5)
    Bt, tnames = mo.basis(p, er, present)
    amps = 10.0 ** rng.uniform(1.0, 2.6, size=(67, Bt.shape[1]))
    y = rng.poisson(amps @ Bt.astype(np.float64).T + (3.0 if use_snip else 0.0)).astype(np.float32)
    assert y.max() < 2048
    idx = np.r_[60:900, 950:1500] if with_idx else None
    lam = 2e-5
    vol = torch.from_numpy(y).to(dtype).cuda()
    maps, plan = opt.fit_map(vol, er, fitted, p, use_snip=use_snip, indices=idx, l1_lambda=lam, return_plan=True)
    B, names = mo.basis(p, er, fitted)
    B = B.astype(np.float64).T
    got = np.stack([maps[n].cpu().numpy() for n in names], axis=1).astype(np.float64)
    args = [p[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET",
                           "FWHM_FANOPRIME")]
    bkg = mo.snip_bkg(y, er, *args) if use_snip else np.zeros_like(y)
    effect = 0.0
    for px in range(y.shape[0]):
        r = y[px].astype(np.float64) - bkg[px]
        Bm, rm = (B, r) if idx is None else (B[:, idx], r[idx])
        c = lam * float((rm ** 2).sum()) / len(names)
        exact = mo.fit_nnls_penalised(rm, Bm, None, c)
        plain = mo.fit_nnls_penalised(rm, Bm, None, 0.0)
        top = exact.max()
        interior = exact > 1e-3 * top
        assert np.abs(got[px][interior] / exact[interior] - 1).max(initial=0.0) <= 1e-4, px
        assert np.abs(got[px][~interior] - exact[~interior]).max(initial=0.0) <= 1e-4 * top, px
        effect = max(effect, float(np.abs(plain - exact).max() / top))
    assert effect > 1e-3  # the penalty did move the answer (100x the tolerance checked above)
    with pytest.raises(NotImplementedError):  # the L1 maps are not refined per pixel
        opt.fit_map(vol, er, fitted, p, use_snip=use_snip, loss="l1", refine=("ENERGY_OFFSET",))


def test_global_fit_with_l1_loss_recovers_the_calibration():
    """fit_spec(tune_params=True, loss="l1"): start from a shifted calibration, with outlier channels in the
    spectrum.  The result must be at least as good as the exact LAD optimum AT THE TRUE calibration, and the
    exact LAD optimum at the calibration it reports must reproduce its objective.  (The LAD-optimal offset of
    a Poisson spectrum is not the true one -- an oracle scan puts it ~0.7 eV higher for this seed -- so the
    calibration itself is only required to come back from the 4 eV / 6 % perturbation to within 1.2 eV / 5 %.)"""
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_param_vals
    from oracle import maps_oracle as mo

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    elems = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
    er = (0, 2047)
    rng = np.random.default_rng(21)
    B, names = mo.basis(p, er, elems)
    amps = 10.0 ** rng.uniform(2.5, 3.5, size=B.shape[1])
    y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
    y[400] += 20000.0
    y[1200] += 9000.0
    start = dict(p, ENERGY_OFFSET=p["ENERGY_OFFSET"] + 0.004, FWHM_OFFSET=p["FWHM_OFFSET"] * 1.06)
    tensors, spec_fit, bkg, trace = opt.fit_spec(
        y, er, elements_to_fit=elems, fitting_params=["ENERGY_OFFSET", "FWHM_OFFSET"], init_param_vals=start,
        tune_params=True, use_snip=False, loss="l1", n_iter=40, progress_bar=False)
    _, best = mo.fit_lad(y.astype(np.float64), B.astype(np.float64).T, None, 0.0)
    final = float(np.abs(spec_fit.astype(np.float64) - y).sum())
    assert trace[-1] <= trace[0] and abs(trace[-1] - final) <= 1e-3 * best
    assert final <= best * (1 + 1e-4)
    found = dict(p, ENERGY_OFFSET=tensors["ENERGY_OFFSET"].item(), FWHM_OFFSET=tensors["FWHM_OFFSET"].item())
    Bf, _ = mo.basis(found, er, elems)
    _, best_found = mo.fit_lad(y.astype(np.float64), Bf.astype(np.float64).T, None, 0.0)
    assert abs(final - best_found) <= 2e-4 * best_found
    assert abs(found["ENERGY_OFFSET"] - p["ENERGY_OFFSET"]) < 1.2e-3
    assert abs(found["FWHM_OFFSET"] / p["FWHM_OFFSET"] - 1) < 5e-2


@pytest.mark.parametrize("use_snip,masked", [(False, False), (True, False), (False, True)])
def test_l1_maps_reach_the_lad_optimum_per_pixel(use_snip, masked):
    """fit_map(loss="l1"): every pixel's objective within 2e-5 (relative; typically 1e-7..5e-6) of the exact linear programme
    (oracle.fit_lad on the oracle's basis and background), 256 pixels with outliers; and it differs visibly from the
    least-squares fit, whose L1 objective is worse."""
    from mapstorch_b200 import opt, util
    from mapstorch_b200.default import default_param_vals
    from oracle import maps_oracle as mo

    elems = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    er = (0, 2047)
    rng = np.random.default_rng(9)
    B, names = mo.basis(p, er, elems)
    B64 = B.astype(np.float64)
    n_px = 256
    amps = 10.0 ** rng.uniform(0.5, 3.0, size=(n_px, B.shape[1]))
    y = rng.poisson(amps @ B64.T + 1.0).astype(np.float32)
    for i in range(n_px):  # a few hot channels per pixel: what the robust objective is for
        y[i, rng.integers(0, 2048, 4)] += rng.uniform(50, 400, 4).astype(np.float32)
    idx = None
    if masked:
        rr = util.get_peak_ranges(names, 12.0, p["COMPTON_ANGLE"], p["ENERGY_OFFSET"], p["ENERGY_SLOPE"], p["ENERGY_QUADRATIC"], er)
        idx = util.peak_range_indices(rr, 2048)
    maps = opt.fit_map(torch.from_numpy(y).reshape(16, 16, -1), er, elems, init_param_vals=p, use_snip=use_snip,
                       indices=idx, loss="l1")
    lsq = opt.fit_map(torch.from_numpy(y).reshape(16, 16, -1), er, elems, init_param_vals=p, use_snip=use_snip, indices=idx)
    got = torch.stack([maps[n] for n in names], dim=-1).reshape(n_px, -1).cpu().numpy().astype(np.float64)
    ls = torch.stack([lsq[n] for n in names], dim=-1).reshape(n_px, -1).cpu().numpy().astype(np.float64)
    loss = maps["loss"].reshape(-1).cpu().numpy()
    assert (got >= 0).all() and maps["Fe"].shape == (16, 16)
    bkg = np.zeros_like(y)
    if use_snip:
        bkg = mo.snip_bkg(y, er, *[p[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH",
                                                    "FWHM_OFFSET", "FWHM_FANOPRIME")])
    sel = np.arange(2048) if idx is None else np.asarray(idx)
    worst, gain = 0.0, []
    for i in range(0, n_px, 4):  # the LP takes ~0.3 s per pixel: every fourth pixel
        t = (y[i] - bkg[i]).astype(np.float64)[sel]
        _, best = mo.fit_lad(t, B64[sel].T)
        mine = np.abs(B64[sel] @ got[i] - t).sum()
        assert abs(mine - loss[i]) <= 2e-4 * best, i   # the kernel's own objective is that of the returned amplitudes
        worst = max(worst, (mine - best) / best)
        gain.append(np.abs(B64[sel] @ ls[i] - t).sum() / mine)
    print(f"L1 maps (snip={use_snip}, masked={masked}): worst relative gap to the LP optimum {worst:.2e}; "
          f"least-squares amplitudes are {np.mean(gain) - 1:.2%} worse in the L1 objective")
    # measured on B200: 1.9e-6 (plain), 5.2e-6 (SNIP residuals), ~1e-5 (peak-window masks); median ~2e-7
    assert worst <= 2e-5
    assert np.mean(gain) > 1.0005


def test_l1_maps_with_amplitude_penalty_reach_the_lp_optimum():
    """fit_map(loss="l1", l1_lambda=...): sum|B^T x - t| + c_p sum(x), c_p = l1_lambda * sum|t| / n_elements per pixel
    (opt.py:372-381 with the L1 loss), against the exact linear programme (oracle.fit_lad with that c) on every eighth
    of 128 pixels with outliers; the penalty is strong enough to move the answer."""
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_param_vals
    from oracle import maps_oracle as mo

    elems = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    er = (0, 2047)
    rng = np.random.default_rng(19)
    B, names = mo.basis(p, er, elems)
    B64 = B.astype(np.float64)
    n_px, lam = 128, 2e-3  # between "nothing moves" (1e-4) and "everything is switched off" (1e-2) on these pixels
    amps = 10.0 ** rng.uniform(0.5, 3.0, size=(n_px, B.shape[1]))
    y = rng.poisson(amps @ B64.T + 1.0).astype(np.float32)
    for i in range(n_px):
        y[i, rng.integers(0, 2048, 4)] += rng.uniform(50, 400, 4).astype(np.float32)
    vol = torch.from_numpy(y).reshape(8, 16, -1)
    maps = opt.fit_map(vol, er, elems, init_param_vals=p, use_snip=False, loss="l1", l1_lambda=lam)
    plain = opt.fit_map(vol, er, elems, init_param_vals=p, use_snip=False, loss="l1")
    got = torch.stack([maps[n] for n in names], dim=-1).reshape(n_px, -1).cpu().numpy().astype(np.float64)
    unpen = torch.stack([plain[n] for n in names], dim=-1).reshape(n_px, -1).cpu().numpy().astype(np.float64)
    loss = maps["loss"].reshape(-1).cpu().numpy()
    assert (got >= 0).all()
    worst, moved, n_off = 0.0, 0.0, 0
    for i in range(0, n_px, 8):
        t = y[i].astype(np.float64)
        c = lam * np.abs(t).sum() / len(names)
        exact, best = mo.fit_lad(t, B64.T, None, c)
        n_off += int((exact <= 1e-9 * exact.max()).sum())
        assert exact.max() > 0
        mine = np.abs(B64 @ got[i] - t).sum() + c * got[i].sum()
        assert abs(mine - loss[i]) <= 2e-4 * best, i
        worst = max(worst, (mine - best) / best)
        other = np.abs(B64 @ unpen[i] - t).sum() + c * unpen[i].sum()
        moved = max(moved, (other - mine) / mine)
    print(f"penalised L1 maps: worst relative gap to the LP optimum {worst:.2e}; the unpenalised answer is up to {moved:.2%} worse")
    assert worst <= 5e-5
    assert moved > 1e-4 and n_off >= 8  # the penalty switched components off without emptying the pixels
