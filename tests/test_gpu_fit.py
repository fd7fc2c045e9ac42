"""K3 (tensor-core contraction) + the batched non-negative solve + fit_map, on the GPU.

Oracle: float64 least squares / NNLS on the oracle's basis (what the reference's per-spectrum
Adam fit converges to) and the converged reference fits in tests/golden/fit_golden.npz.
Tolerance (north_star): amplitudes within 1e-4 relative on interior components."""
import numpy as np
import pytest
import torch

from conftest import golden_params

pytestmark = pytest.mark.gpu

ELEMS10 = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
ELEMS30 = ("Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge "
           "Ba_L Ce_L Gd_L W_L Pt_L Au_L Au_M Pt_M Hg_M W_M Si_Si").split()


def params12(n_ch=2048):
    from mapstorch_b200.default import default_param_vals

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    if n_ch == 4096:
        p.update(ENERGY_SLOPE=0.00599, ENERGY_QUADRATIC=0.0)
    return p


def synth(p, er, elems, n_px, seed, lo=1.5, hi=3.0, flat=0.0, n_full=None):
    from oracle import maps_oracle as mo

    rng = np.random.default_rng(seed)
    n_full = n_full or (er[1] + 1)
    B, names = mo.basis(p, (0, n_full - 1), elems)
    amps = 10.0 ** rng.uniform(lo, hi, size=(n_px, B.shape[1]))
    lam = amps @ B.astype(np.float64).T + flat
    return rng.poisson(lam).astype(np.float32), names


def check_against(x, ref, tol=1e-4, interior_frac=1e-3):
    """interior components (> interior_frac of the pixel maximum) within tol relative; the rest
    within tol * pixel maximum."""
    scale = np.abs(ref).max(axis=1, keepdims=True)
    interior = ref > interior_frac * scale
    rel = np.abs(x - ref) / np.where(interior, np.abs(ref), scale)
    assert rel.max() <= tol, rel.max()


@pytest.mark.parametrize("dtype,n_ch,elems,n_px", [
    (torch.float16, 2048, ELEMS10, 300), (torch.float32, 2048, ELEMS10, 257),
    (torch.float16, 4096, ELEMS30, 513), (torch.float32, 4096, ELEMS30, 128),
    (torch.float32, 2048, ELEMS10, 1), (torch.float16, 2048, ELEMS10, 127),
])
def test_unconstrained_solution_matches_float64_lstsq(dtype, n_ch, elems, n_px):
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12(n_ch)
    er = (0, n_ch - 1)
    y, names = synth(p, er, elems, n_px, seed=n_px)
    assert y.max() < 2048
    plan = opt.MapFitPlan(er, elems, p, use_snip=False)
    assert plan.names == names
    B, _ = mo.basis(p, er, elems)
    ref = mo.fit_lstsq(y, B)
    dev = torch.from_numpy(y).to("cuda", dtype)
    x = plan.fit(dev, nonneg=False).cpu().numpy().T
    check_against(x, ref, 1e-4)
    # integer counts x integer digit planes: the tensor-core accumulation is exact, what is left is
    # the fp32 recombination and the 30-bit quantisation of the pseudo-inverse
    assert (np.abs(x - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 5e-7
    xs = plan.fit(dev, nonneg=False, engine="simt").cpu().numpy().T
    check_against(xs, ref, 1e-4, interior_frac=1e-2)  # plain fp32 FMA chains: ~3e-7 of the pixel maximum


def test_energy_range_crop_and_unaligned_rows():
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12()
    er = (50, 1450)
    y, names = synth(p, er, ELEMS10, 200, seed=5, n_full=2048)
    B, _ = mo.basis(p, er, ELEMS10)
    ref = mo.fit_lstsq(y[:, 50:1451], B)
    plan = opt.MapFitPlan(er, ELEMS10, p, use_snip=False)
    x = plan.fit(torch.from_numpy(y).cuda(), nonneg=False).cpu().numpy().T
    check_against(x, ref, 1e-4)
    # rows whose pitch is not a multiple of 16 bytes cannot go through TMA: CUDA-core path
    odd = torch.from_numpy(np.ascontiguousarray(y[:, :2046])).cuda()
    x2 = plan.fit(odd, nonneg=False).cpu().numpy().T
    check_against(x2, ref, 1e-4, interior_frac=1e-2)


def test_nonnegative_solve_matches_scipy_nnls():
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12(4096)
    er = (0, 4095)
    y, names = synth(p, er, ELEMS30, 96, seed=21, lo=-0.5, hi=1.5)  # low counts: many clamped components
    B, _ = mo.basis(p, er, ELEMS30)
    ref = mo.fit_nnls(y, B)
    assert (ref == 0).mean() > 0.05
    plan = opt.MapFitPlan(er, ELEMS30, p, use_snip=False)
    n_fixed = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = plan.fit(torch.from_numpy(y).half().cuda(), nonneg=True, n_fixed=n_fixed).cpu().numpy().T
    assert (x >= 0).all() and int(n_fixed) > 0
    scale = ref.max(axis=1, keepdims=True)
    assert (np.abs(x - ref) / scale).max() <= 1e-4
    interior = ref > 1e-3 * scale
    assert (np.abs(x - ref)[interior] / ref[interior]).max() <= 2e-3
    # objective value: never worse than scipy's
    r_mine = ((y - x @ B.T.astype(np.float64)) ** 2).sum(axis=1)
    r_ref = ((y - ref @ B.T.astype(np.float64)) ** 2).sum(axis=1)
    assert np.all(r_mine <= r_ref * (1 + 1e-6))


def test_fit_map_matches_converged_reference_fits(fit_golden):
    """fit_golden holds fit_spec(tune_params=False, optimizer='adam', n_iter=4000) of the reference."""
    from mapstorch_b200 import opt

    g = fit_golden
    for tag in [str(t) for t in g["fit_runs"]]:
        p = golden_params(g, tag)
        er = tuple(int(v) for v in g[f"{tag}_er"])
        names = [str(n) for n in g[f"{tag}_names"]]
        use_snip = bool(g[f"{tag}_use_snip"])
        spec = g[f"{tag}_spec"]
        maps = opt.fit_map(spec.reshape(1, 1, -1), er, names[:-2], init_param_vals=p, use_snip=use_snip, device="cpu")
        ref = np.array([10.0 ** v for v in g[f"{tag}_logamps"]])
        mine = np.array([float(maps[n]) for n in names])
        interior = ref > 1e-3 * ref.max()
        assert (np.abs(mine - ref)[interior] / ref[interior]).max() <= 1e-4, tag
        assert np.all(mine[~interior] <= ref[~interior] + 1e-4 * ref.max()), tag
        tensors, spec_fit, bkg, trace = opt.fit_spec(spec, er, names[:-2], init_param_vals=p, tune_params=False,
                                                     use_snip=use_snip)
        assert np.abs(spec_fit - g[f"{tag}_spec_fit"]).max() <= 2e-4 * g[f"{tag}_spec_fit"].max()
        assert np.abs(bkg - g[f"{tag}_bkg"]).max() <= 1e-5 * max(g[f"{tag}_bkg"].max(), 1e-30)
        assert trace[-1] <= g[f"{tag}_trace"][-1] * (1 + 1e-5)
        assert abs(float(tensors["Fe"]) - dict(zip(names, g[f"{tag}_logamps"]))["Fe"]) < 1e-4


def test_fit_map_with_snip_background_and_shards():
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12()
    er = (0, 2047)
    y, names = synth(p, er, ELEMS10, 8 * 12, seed=9, flat=3.0)
    vol = y.reshape(8, 12, -1)
    pr = [p[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME")]
    bkg = mo.snip_bkg(y, er, *pr)
    B, _ = mo.basis(p, er, ELEMS10)
    ref = mo.fit_nnls(y, B, bkg)
    maps, plan = opt.fit_map(torch.from_numpy(vol), er, ELEMS10, init_param_vals=p, use_snip=True, return_plan=True)
    x = torch.stack([maps[n] for n in names], dim=-1).reshape(-1, len(names)).cpu().numpy()
    check_against(x, ref, 1e-4)
    assert maps["Fe"].shape == (8, 12) and maps["Fe"].is_cuda
    # row shards give bit-identical maps (pixels are independent): the multi-GPU partitioning
    parts = [opt.fit_map(torch.from_numpy(vol[r0:r1]), er, ELEMS10, plan=plan) for r0, r1 in ((0, 3), (3, 8))]
    for n in names:
        assert torch.equal(torch.cat([q[n] for q in parts], dim=0), maps[n]), n
    # restricting the objective to peak windows (`indices`) == least squares on the masked channels
    from mapstorch_b200 import util

    rr = util.get_peak_ranges(names, 12.0, p["COMPTON_ANGLE"], p["ENERGY_OFFSET"], p["ENERGY_SLOPE"], p["ENERGY_QUADRATIC"], er)
    idx = util.peak_range_indices(rr, 2048)
    m2 = opt.fit_map(torch.from_numpy(vol), er, ELEMS10, init_param_vals=p, use_snip=True, indices=idx)
    ref2 = mo.fit_nnls(y[:, idx], B[idx], bkg[:, idx])
    x2 = torch.stack([m2[n] for n in names], dim=-1).reshape(-1, len(names)).cpu().numpy()
    check_against(x2, ref2, 1e-4)


def test_gated_off_component_and_escape_factor():
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12()
    er = (0, 2047)
    elems = ELEMS10 + ["Pb_L"]  # every Pb L edge lies above 12 keV: column is identically zero
    y, _ = synth(p, er, ELEMS10, 40, seed=2)
    maps = opt.fit_map(torch.from_numpy(y), er, elems, init_param_vals=p, use_snip=False, escape_factor=0.05)
    assert float(maps["Pb_L"].abs().max()) == 0.0
    B, names = mo.basis(p, er, ELEMS10, escape_factor=0.05)
    ref = mo.fit_nnls(y, B)
    x = torch.stack([maps[n] for n in names], dim=-1).cpu().numpy()
    check_against(x, ref, 1e-4)


def test_full_size_linearity_property():
    """BASELINE config 2 size (256x256x2048, 10 elements): fit(Y1 + Y2) == fit(Y1) + fit(Y2) for the
    unconstrained solve, and duplicated pixels give identical answers wherever they sit."""
    from mapstorch_b200 import opt

    p = params12()
    er = (0, 2047)
    plan = opt.MapFitPlan(er, ELEMS10, p, use_snip=False)
    gen = torch.Generator(device="cuda").manual_seed(3)
    lam = plan.basis.T @ (10 ** (1.0 + 1.5 * torch.rand(len(plan.names), 256, device="cuda", generator=gen)))
    y1 = torch.poisson(lam.T.repeat(256, 1), generator=gen)
    y2 = torch.poisson(lam.T.repeat(256, 1).flip(0), generator=gen)
    assert y1.shape == (65536, 2048) and float((y1 + y2).max()) < 2048
    a, b, c = (plan.fit(t, nonneg=False) for t in (y1, y2, y1 + y2))
    scale = c.abs().max(dim=0).values
    assert float(((a + b - c).abs() / scale).max()) <= 2e-5
    y1[40000] = y1[3]
    d = plan.fit(y1.half(), nonneg=False)
    assert torch.equal(d[:, 40000], d[:, 3])
    assert float(((d - a)[:, :40000].abs() / scale[:40000]).max()) <= 2e-5


def test_fp32_counts_above_2048_take_the_exact_split_path():
    """tf32 operands hold integers below 2048 exactly; brighter fp32 volumes must not lose bits."""
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12()
    er = (0, 2047)
    y, names = synth(p, er, ELEMS10, 300, seed=31, lo=4.5, hi=5.3)
    assert y.max() > 10000
    B, _ = mo.basis(p, er, ELEMS10)
    ref = mo.fit_lstsq(y, B)
    plan = opt.MapFitPlan(er, ELEMS10, p, use_snip=False)
    dev = torch.from_numpy(y).cuda()
    x = plan.fit(dev, nonneg=False).cpu().numpy().T
    check_against(x, ref, 1e-4)
    assert (np.abs(x - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 2e-6
    # the unchecked direct path truncates operands: visibly worse, which is why the check exists
    bad = plan.fit(dev, nonneg=False, exact_counts=True).cpu().numpy().T
    assert (np.abs(bad - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() > 1e-5
    # host volumes are streamed slab by slab; only the slab with large counts is redone exactly
    mixed = np.concatenate([np.minimum(y, 100.0), y[:50]])
    xh = plan.fit_host(torch.from_numpy(mixed), nonneg=False, chunk_px=128).cpu().numpy().T
    check_against(xh, mo.fit_lstsq(mixed, B), 1e-4)


def test_pixel_major_layout_is_the_same_fit():
    from mapstorch_b200 import opt

    p = params12(4096)
    er = (0, 4095)
    y, names = synth(p, er, ELEMS30, 700, seed=77, lo=0.0, hi=2.5)
    plan = opt.MapFitPlan(er, ELEMS30, p, use_snip=False)
    dev = torch.from_numpy(y).half().cuda()
    plan.epilogue_fixup = False  # same (fp64, list-driven) non-negative solve for both layouts
    ep = plan.fit(dev)
    pe = plan.fit(dev, layout="pe")
    assert pe.shape == (700, plan.e_pad) and torch.equal(pe[:, : len(names)].T, ep)
    # the in-epilogue fp32 fix-up of small zero sets agrees with the fp64 solver to fp32 rounding
    plan.epilogue_fixup = True
    fused = plan.fit(dev)
    assert float(((fused - ep).abs() / ep.abs().max(dim=0, keepdim=True).values).max()) <= 3e-5
    basis = plan.basis.double()
    y64 = dev.double()
    obj = lambda x: ((y64 - x.double().T @ basis) ** 2).sum(dim=1)
    assert bool((obj(fused) <= obj(ep) * (1 + 1e-9) + 1e-6).all())
    ep_s = plan.fit(dev, engine="simt")
    pe_s = plan.fit(dev, engine="simt", layout="pe")
    assert torch.equal(pe_s[:, : len(names)].T, ep_s)
    # 12 components: e_pad = 16, padding columns stay zero
    y2, names2 = synth(params12(), (0, 2047), ELEMS10, 130, seed=78)
    plan2 = opt.MapFitPlan((0, 2047), ELEMS10, params12(), use_snip=False)
    pe2 = plan2.fit(torch.from_numpy(y2).cuda(), layout="pe", exact_counts=True)
    assert pe2.shape == (130, 16) and float(pe2[:, 12:].abs().max()) == 0.0
    assert torch.equal(pe2[:, :12].T, plan2.fit(torch.from_numpy(y2).cuda(), exact_counts=True))


def test_all_default_components_at_high_incident_energy():
    """85 live components at 17.5 keV: more than one tensor-core tile (80) -> contracted in two groups,
    non-negative solve by the general kernel.  Checked against float64 NNLS on the oracle basis."""
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_fitting_elems, default_param_vals
    from oracle import maps_oracle as mo

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=17.5)
    er = (0, 2047)
    elems = [e for e in default_fitting_elems if e not in ("COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE")]
    plan = opt.MapFitPlan(er, elems, p, use_snip=False)
    assert plan.n_live == 85 and len(plan.groups) == 2 and len(plan.names) == 93
    B, names = mo.basis(p, er, elems)
    assert names == plan.names
    rng = np.random.default_rng(12)
    live = np.abs(B).sum(axis=0) > 0
    amps = np.where(live, 10.0 ** rng.uniform(1.0, 3.0, size=(24, len(names))), 0.0)
    y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
    x = plan.fit(torch.from_numpy(y).cuda(), nonneg=False).cpu().numpy().T
    ref = mo.fit_lstsq(y, B[:, live])
    assert (np.abs(x[:, live] - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 2e-5  # cond(G) ~ 9e3 here
    assert not x[:, ~live].any()
    xn = plan.fit(torch.from_numpy(y).cuda(), nonneg=True).cpu().numpy().T
    refn = mo.fit_nnls(y, B[:, live])
    assert (xn >= 0).all()
    assert (np.abs(xn[:, live] - refn) / refn.max(axis=1, keepdims=True)).max() <= 1e-3
    r_mine = ((y - xn @ B.T.astype(np.float64)) ** 2).sum(axis=1)
    r_ref = ((y - refn @ B[:, live].T.astype(np.float64)) ** 2).sum(axis=1)
    assert np.all(r_mine <= r_ref * (1 + 1e-4))


def test_fit_map_with_override_file_constants(tmp_path):
    """e_consts edited by a MAPS override file (Fe K-beta x1.2, Ge absorption column, ...) reach the basis,
    the Gram matrix and the per-pixel refinement tables: fit a volume generated WITH those constants and
    compare with float64 NNLS on the oracle's basis for the same constants; the default constants must fit
    the same volume visibly worse."""
    from mapstorch_b200 import constant, io as mio, opt
    from mapstorch_b200.default import default_elem_info
    from oracle import maps_oracle as mo

    path = tmp_path / "maps_fit_parameters_override.txt"
    mio.write_override_params_file(str(path), {"COHERENT_SCT_ENERGY": 12.0}, ["Fe", "Cu", "Zn", "Gd_L", "Pt_L"], "Ge", 0.0)
    consts = constant.read_constants(str(path), default_elem_info)
    cfg = mio.parse_override_params_file(str(path))
    p = params12()
    elems, er = cfg["elements_to_fit"], (0, 2047)
    assert cfg["detector_material"] == "Ge" and elems == ["Fe", "Cu", "Zn", "Gd_L", "Pt_L"]
    B, names = mo.basis(p, er, elems, e_consts=consts, elem_info=default_elem_info)
    rng = np.random.default_rng(8)
    amps = 10.0 ** rng.uniform(2.0, 3.0, size=(96, B.shape[1]))
    y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
    maps, plan = opt.fit_map(torch.from_numpy(y).cuda(), er, elems, p, use_snip=False, e_consts=consts,
                             return_plan=True)
    assert plan.e_consts is consts
    got = np.stack([maps[n].cpu().numpy() for n in names], axis=1)
    ref = mo.fit_nnls(y, B)
    check_against(got, ref)
    plain = opt.fit_map(torch.from_numpy(y).cuda(), er, elems, p, use_snip=False)
    worse = np.stack([plain[n].cpu().numpy() for n in names], axis=1)
    resid = lambda a: float((((a.astype(np.float64) @ B.astype(np.float64).T) - y) ** 2).sum())  # noqa: E731
    assert resid(worse) > 1.05 * resid(got)
    # refinement builds its line tables from the plan's constants too (with ~1e3 counts per pixel the calibration
    # itself is statistically loose, so the check is on the objective and the bounds)
    fine = opt.fit_map(torch.from_numpy(y[:16]).cuda(), er, elems, p, use_snip=False, e_consts=consts,
                       refine=("ENERGY_OFFSET", "FWHM_OFFSET"))
    assert float((fine["ENERGY_OFFSET"] - p["ENERGY_OFFSET"]).abs().max()) <= 0.03 + 1e-6
    assert float((fine["FWHM_OFFSET"] / p["FWHM_OFFSET"] - 1).abs().max()) <= 0.30 + 1e-6
    sse = (((got[:16].astype(np.float64) @ B.astype(np.float64).T) - y[:16]) ** 2).sum(axis=1)
    assert np.all(fine["loss"].cpu().numpy() <= sse * (1 + 1e-5))  # refining can only lower the objective


@pytest.mark.parametrize("case", ["snip", "snip_penalty", "wide", "refine", "host_penalty"])
def test_chunked_paths_equal_the_single_chunk_result(monkeypatch, case):
    """Volumes larger than one scratch slab are processed slab by slab (SNIP residual, exact operand split,
    penalty shift, refinement, host streaming).  Pixels are independent, so forcing tiny slabs must reproduce
    the single-slab maps bit for bit -- this is what catches a pitch or offset mixed up between a slab and
    the whole map."""
    from mapstorch_b200 import opt

    p = params12()
    er = (0, 2047)
    bright = case == "wide"
    y, names = synth(p, er, ELEMS10, 101, seed=13, lo=4.2 if bright else 1.5, hi=5.0 if bright else 3.0,
                     flat=0.0 if bright else 3.0)
    dev = torch.from_numpy(y).cuda()
    kw = dict(init_param_vals=p, use_snip=case.startswith("snip"))
    if case.endswith("penalty"):
        kw["l1_lambda"] = 2e-5
    if case == "refine":
        kw.update(refine=("ENERGY_OFFSET", "FWHM_OFFSET"), refine_iters=8)
    whole = opt.fit_map(dev, er, ELEMS10, **kw)
    monkeypatch.setattr(opt, "SNIP_CHUNK_PX", 24)
    if case == "host_penalty":
        parts = opt.fit_map(torch.from_numpy(y), er, ELEMS10, chunk_px=24, **kw)
    else:
        parts = opt.fit_map(dev, er, ELEMS10, **kw)
    assert set(whole) == set(parts)
    for k in whole:
        assert torch.equal(whole[k], parts[k]), (case, k)
    if bright:
        assert y.max() >= 2048  # the exact-split path was the one exercised


def test_bright_fp16_volume_accuracy():
    """Counts near the top of fp16's exact-integer range (2047) in many channels: sum_k y_k |digit_k| exceeds 2^24,
    so the tensor-core accumulation is no longer exact integer arithmetic and rounds like ordinary fp32
    (csrc/mt_fit.cu header).  The result must still sit well inside the 1e-4 bar."""
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12(4096)
    er = (0, 4095)
    B, names = mo.basis(p, er, ELEMS30)
    rng = np.random.default_rng(404)
    amps = 10.0 ** rng.uniform(2.0, 3.0, size=(384, B.shape[1]))
    lam = amps @ B.astype(np.float64).T
    lam *= 1900.0 / lam.max(axis=1, keepdims=True)  # brightest channel of every pixel near 1900 counts
    y = np.minimum(rng.poisson(lam + 600.0), 2047).astype(np.float32)  # + a bright flat floor: every channel is large
    assert (y > 512).mean() > 0.9 and y.max() == 2047
    ref = mo.fit_lstsq(y, B)
    plan = opt.MapFitPlan(er, ELEMS30, p, use_snip=False)
    x = plan.fit(torch.from_numpy(y).half().cuda(), nonneg=False).cpu().numpy().T
    scale = np.abs(ref).max(axis=1, keepdims=True)
    err = (np.abs(x - ref) / scale).max()
    interior = np.abs(ref) > 1e-3 * scale
    rel = (np.abs(x - ref)[interior] / np.abs(ref)[interior]).max()
    print(f"bright fp16: max error {err:.3g} of the pixel maximum, {rel:.3g} relative on interior components")
    assert err <= 2e-5 and rel <= 1e-4


def test_non_integer_fp32_volume_is_not_truncated():
    """Normalised / dead-time corrected data are not integer counts: values with more than 11 significant bits are
    detected on the device (mt_operand_inexact) and take the exact split path, whatever their magnitude."""
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12()
    er = (0, 2047)
    y, names = synth(p, er, ELEMS10, 300, seed=41, lo=1.5, hi=3.0)
    y = (y * np.float32(0.3712)).astype(np.float32)  # all values < 2048, almost none tf32-exact
    assert y.max() < 2048
    B, _ = mo.basis(p, er, ELEMS10)
    ref = mo.fit_lstsq(y, B)
    plan = opt.MapFitPlan(er, ELEMS10, p, use_snip=False)
    dev = torch.from_numpy(y).cuda()
    x = plan.fit(dev, nonneg=False).cpu().numpy().T
    assert plan._last_wide
    assert (np.abs(x - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 2e-6
    # integer counts below 2048 keep the single-pass path
    plan.fit(torch.from_numpy(np.rint(y)).cuda(), nonneg=False)
    assert not plan._last_wide
    # streamed host volumes: only the inexact slabs are redone
    mixed = np.concatenate([np.rint(y[:128]), y[:60]])
    xh = plan.fit_host(torch.from_numpy(mixed), nonneg=False, chunk_px=64).cpu().numpy().T
    refm = mo.fit_lstsq(mixed, B)
    assert (np.abs(xh - refm) / np.abs(refm).max(axis=1, keepdims=True)).max() <= 2e-6


def test_host_results_of_consecutive_calls_are_independent():
    """fit_map(device="cpu") returns tensors the caller owns: plans (and their buffers) are cached across calls, so a
    second fit with the same calibration and shape must not overwrite the first result."""
    from mapstorch_b200 import opt

    p = params12()
    er = (0, 2047)
    y1, names = synth(p, er, ELEMS10, 64, seed=51)
    y2, _ = synth(p, er, ELEMS10, 64, seed=52)
    kw = dict(init_param_vals=p, use_snip=False, device="cpu")
    first = opt.fit_map(torch.from_numpy(y1).reshape(8, 8, -1), er, ELEMS10, **kw)
    keep = {k: v.clone() for k, v in first.items()}
    second = opt.fit_map(torch.from_numpy(y2).reshape(8, 8, -1), er, ELEMS10, **kw)
    for k in keep:
        assert torch.equal(first[k], keep[k]), k
        assert not torch.equal(first[k], second[k]), k
        assert first[k].data_ptr() != second[k].data_ptr()


def test_caller_buffer_rows_of_dead_components_are_zeroed():
    from mapstorch_b200 import opt

    p = params12()
    er = (0, 2047)
    elems = ELEMS10 + ["Pb_L"]  # no open Pb L line at 12 keV
    y, _ = synth(p, er, ELEMS10, 32, seed=61)
    plan = opt.MapFitPlan(er, elems, p, use_snip=False)
    out = torch.full((len(plan.names), 32), 7.0, device="cuda")
    plan.fit(torch.from_numpy(y).cuda(), out=out)
    assert float(out[plan.names.index("Pb_L")].abs().max()) == 0.0


@pytest.mark.parametrize("n_px", [64, 100, 148 * 64 + 37, 148 * 64 * 2 - 17, 148 * 64 * 3 - 5, 148 * 64 * 5 + 130,
                                  148 * 64 * 8 + 64 * 77 + 17, 148 * 64 * 9 + 64 * 3 + 1])
def test_partial_last_tiles_of_every_granule_count(n_px):
    """K3's tile schedule (csrc/mt_fit.cu): whole 256-pixel tiles round-robin, then the rest of the map in 64-pixel
    granules, one partial tile per CTA -- one, two or three granules, i.e. a 64-row MMA, a 128-row MMA, or both, with the
    64-row accumulator in lanes 0..15 of every TMEM quarter.  The two largest sizes give every CTA at least two whole
    tiles, which switches on the clusters of two CTAs that share the pack slabs by TMA multicast and walk their tiles in
    lockstep -- with partners whose last tiles differ in size, and (the last case) partners of which only one HAS a
    last tile, so that the other runs a round without spectra.  Every pixel must still be the float64 least-squares fit
    (exact integer accumulation: 5e-7 of the pixel maximum), wherever in the schedule it falls."""
    from mapstorch_b200 import opt
    from oracle import maps_oracle as mo

    p = params12(2048)
    er = (0, 2047)
    y, names = synth(p, er, ELEMS10, 2048, seed=3)
    reps = -(-n_px // y.shape[0])
    idx = np.random.default_rng(n_px).integers(0, y.shape[0], size=n_px)  # pixels drawn with repetition, in random order
    yy = y[idx]
    plan = opt.MapFitPlan(er, ELEMS10, p, use_snip=False)
    B, _ = mo.basis(p, er, ELEMS10)
    ref = mo.fit_lstsq(y, B)[idx]
    for dtype in (torch.float16, torch.float32):
        x = plan.fit(torch.from_numpy(yy).to("cuda", dtype), nonneg=False).cpu().numpy().T
        assert x.shape == ref.shape and reps >= 1
        assert (np.abs(x - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 5e-7
    # and the non-negative path through the epilogue fix-up on the same partial tiles
    xn = plan.fit(torch.from_numpy(yy).cuda(), nonneg=True).cpu().numpy().T
    assert xn.min() >= 0 and (np.abs(xn - np.maximum(ref, 0)) / np.abs(ref).max(axis=1, keepdims=True)).max() <= 1e-3
