"""Per-pixel nonlinear refinement (kernel K4) and the global-parameter fit of one spectrum.

``refine_map`` frees up to four calibration parameters per pixel (any of the eight with a non-zero gradient on
model_spec's path: ENERGY_OFFSET / SLOPE / QUADRATIC, FWHM_OFFSET / FANOPRIME, COHERENT_SCT_ENERGY, COMPTON_ANGLE,
COMPTON_FWHM_CORR) together with all amplitudes; step terms (F_STEP_*, COMPTON_F_STEP) are part of the model -- the
per-pixel version of the reference's ``fit_spec(..., fitting_params=[...], tune_params=True)`` (opt.py:165-327) --
starting from the linear solve of ``MapFitPlan.fit``.  ``fit_spec_global`` is ``fit_spec`` with
``tune_params=True`` for one (integrated) spectrum: variable projection over the tunable calibration
parameters, every model evaluation on the GPU (K1 basis + the non-negative solve).
"""
import ctypes

import numpy as np
import torch

from mapstorch_b200 import _native, tables
from mapstorch_b200.default import default_energy_consts, default_elem_info, default_param_vals

f32 = np.float32

REFINE_WARPS = 4  # warps per CTA of mt_refine (csrc/mt_refine.cu: kRefThreads / 32)

DEFAULT_BOUNDS = {
    "ENERGY_OFFSET": ("abs", 0.03),        # +- keV around the global value
    "ENERGY_SLOPE": ("rel", 0.01),         # +- 1 %
    "FWHM_OFFSET": ("rel", 0.30),          # +- 30 %
    "FWHM_FANOPRIME": ("rel", 0.50),       # +- 50 %
    "ENERGY_QUADRATIC": ("abs", 2e-8),     # keV / channel^2
    "COHERENT_SCT_ENERGY": ("abs", 0.15),  # +- keV
    "COMPTON_ANGLE": ("abs", 15.0),        # +- degrees
    "COMPTON_FWHM_CORR": ("rel", 0.40),    # +- 40 %
}


def model_lines(params, names, e_consts=default_energy_consts, elem_info=default_elem_info):
    """(energy, factor, comp, sigma_scale, e296, step, kind) of every line on model_spec's path (map.py:260-309:
    use_step=True, use_tail=False), amplitudes excluded.  factor = ratio / (1 + f_step) as model_elem_spec forms it
    (map.py:183-207), step = f_step of the line (map.py:169-172), COMPTON_F_STEP for the Compton peak (map.py:94-114)."""
    p32 = {k: tables.scalar32(v) for k, v in params.items() if k in default_param_vals}
    e0 = p32["COHERENT_SCT_ENERGY"]
    out = []
    for comp, name in enumerate(names):
        if name == "COHERENT_SCT_AMPLITUDE":
            out.append((e0, f32(1.0), comp, f32(1.0), e0 * f32(2.96), f32(0.0), _native.MT_LINE_ELASTIC))
        elif name == "COMPTON_AMPLITUDE":
            fs = p32["COMPTON_F_STEP"] if p32["COMPTON_F_STEP"] > 0 else f32(0.0)
            ce = tables.compton_energy(params)
            out.append((ce, f32(1.0) / (f32(1.0) + fs), comp, p32["COMPTON_FWHM_CORR"], ce * f32(2.96), fs,
                        _native.MT_LINE_COMPTON))
        else:
            for s in e_consts[name]:
                if s.ratio == 0 or s.energy <= 0 or not s.check_binding_energy(elem_info, e0):
                    continue
                mu, e = f32(s.mu_fraction), f32(s.energy)
                f_step = np.abs(mu * (p32["F_STEP_OFFSET"] + p32["F_STEP_LINEAR"] * e))
                fac = f32(s.ratio)
                if s.ptype.startswith("Ka") or s.ptype.startswith("L") or s.ptype.startswith("Kb"):
                    fac = fac / (f32(1.0) + f_step)
                out.append((e, fac, comp, f32(1.0), f32(s.energy * 2.96), f32(f_step), _native.MT_LINE_ELEMENT))
    return out


def gaussian_lines(params, names, e_consts=default_energy_consts, elem_info=default_elem_info):
    """(energy, factor, comp, sigma_scale, e296) of the Gaussian terms (kept for callers of the round-1 tables)."""
    return [t[:5] for t in model_lines(params, names, e_consts, elem_info)]


def refine_tables(params, names, er, free, bounds=None, k_sigma=5.5, max_theta=_native.MT_MAX_THETA_REFINE, **kw):
    """Host tables of mt_refine / mt_model_jacobian for the components `names` over channels er[0]..er[1]."""
    bounds = dict(DEFAULT_BOUNDS, **(bounds or {}))
    free = list(free)
    if len(free) > max_theta or any(f not in _native.THETA_IDS for f in free):
        raise ValueError(f"free parameters: at most {max_theta} of {sorted(_native.THETA_IDS)}, got {free}")
    lines = sorted(model_lines(params, names, **kw), key=lambda t: float(t[0]))
    n_ch = er[1] - er[0] + 1
    p32 = {k: float(tables.scalar32(params[k])) for k in default_param_vals}
    cfg = _native.MtRefineConfig()
    cfg.n_ch, cfg.ch0, cfg.n_comp, cfg.n_lines, cfg.n_theta = n_ch, er[0], len(names), len(lines), len(free)
    lo_hi = {}
    for k, name in enumerate(free):
        kind, width = bounds[name]
        v = p32[name]
        span = width if kind == "abs" else abs(v) * width
        cfg.theta_id[k] = _native.THETA_IDS[name]
        cfg.theta0[k], cfg.theta_lo[k], cfg.theta_hi[k] = v, v - span, v + span
        lo_hi[name] = (v - span, v + span)
    cfg.energy_offset, cfg.energy_slope, cfg.energy_quad = p32["ENERGY_OFFSET"], p32["ENERGY_SLOPE"], p32["ENERGY_QUADRATIC"]
    cfg.fwhm_offset, cfg.fwhm_fanoprime = p32["FWHM_OFFSET"], p32["FWHM_FANOPRIME"]
    cfg.coherent_sct_energy, cfg.compton_angle = p32["COHERENT_SCT_ENERGY"], p32["COMPTON_ANGLE"]
    cfg.compton_fwhm_corr = p32["COMPTON_FWHM_CORR"]
    # windows: widest sigma and largest energy-axis / line-energy excursion the bounds allow
    fo_hi = lo_hi.get("FWHM_OFFSET", (0, p32["FWHM_OFFSET"]))[1]
    fa_hi = lo_hi.get("FWHM_FANOPRIME", (0, p32["FWHM_FANOPRIME"]))[1]
    corr_hi = lo_hi.get("COMPTON_FWHM_CORR", (0, p32["COMPTON_FWHM_CORR"]))[1]
    ch = np.arange(er[0], er[1] + 1, dtype=np.float64)
    ev = p32["ENERGY_OFFSET"] + p32["ENERGY_SLOPE"] * ch + p32["ENERGY_QUADRATIC"] * ch * ch
    shift = 0.0
    if "ENERGY_OFFSET" in lo_hi:
        shift += lo_hi["ENERGY_OFFSET"][1] - p32["ENERGY_OFFSET"]
    if "ENERGY_SLOPE" in lo_hi:
        shift += (lo_hi["ENERGY_SLOPE"][1] - p32["ENERGY_SLOPE"]) * ch.max()
    if "ENERGY_QUADRATIC" in lo_hi:
        shift += (lo_hi["ENERGY_QUADRATIC"][1] - p32["ENERGY_QUADRATIC"]) * ch.max() ** 2
    # how far the scatter lines can move: the elastic line with COHERENT_SCT_ENERGY, the Compton line with both
    move = {_native.MT_LINE_ELEMENT: 0.0, _native.MT_LINE_ELASTIC: 0.0, _native.MT_LINE_COMPTON: 0.0}
    if "COHERENT_SCT_ENERGY" in lo_hi:
        span = lo_hi["COHERENT_SCT_ENERGY"][1] - p32["COHERENT_SCT_ENERGY"]
        move[_native.MT_LINE_ELASTIC] += span
        move[_native.MT_LINE_COMPTON] += span
    if "COMPTON_ANGLE" in lo_hi:
        e0 = p32["COHERENT_SCT_ENERGY"] + move[_native.MT_LINE_ELASTIC]
        ce = [e0 / (1 + (e0 / 511.0) * (1 - np.cos(np.deg2rad(a)))) for a in np.linspace(*lo_hi["COMPTON_ANGLE"], 33)]
        move[_native.MT_LINE_COMPTON] += max(abs(c - float(tables.compton_energy(params))) for c in ce)
    arr = (_native.MtRefineLine * len(lines))()
    window = np.zeros(len(lines), dtype=np.uint32)
    first = np.full(n_ch, len(lines), dtype=np.int64)
    last = np.zeros(n_ch, dtype=np.int64)
    for i, (e, fac, comp, scale, e296, step, kind) in enumerate(lines):
        arr[i].energy, arr[i].factor, arr[i].comp, arr[i].sigma_scale, arr[i].e296 = float(e), float(fac), comp, float(scale), float(e296)
        arr[i].step, arr[i].kind = float(step), int(kind)
        widest = max(float(scale), corr_hi) if kind == _native.MT_LINE_COMPTON else float(scale)
        e_hi = float(e) + move[kind]
        sigma = widest * np.sqrt((fo_hi / 2.3548) ** 2 + e_hi * 2.96 * fa_hi)
        reach = k_sigma * sigma + shift + move[kind]
        inside = np.nonzero(np.abs(ev - float(e)) <= reach)[0]
        if step > 0:
            # the step term is a plateau below the line (erfc -> 2): its window runs down to the first channel
            top = np.nonzero(ev <= float(e) + reach)[0]
            inside = np.arange(0, (int(top[-1]) + 1) if top.size else 0)
        if inside.size:
            lo_c, hi_c = int(inside[0]), int(inside[-1]) + 1
            window[i] = lo_c | (hi_c << 16)
            first[lo_c:hi_c] = np.minimum(first[lo_c:hi_c], i)
            last[lo_c:hi_c] = np.maximum(last[lo_c:hi_c], i + 1)
    first = np.minimum(first, last)
    chan_lines = (first | (last << 16)).astype(np.uint32)
    order = sorted(range(len(lines)), key=lambda i: lines[i][2])
    comp_lines = np.array(order, dtype=np.int32)
    counts = np.bincount([lines[i][2] for i in order], minlength=len(names))
    comp_start = np.zeros(len(names) + 1, dtype=np.int32)
    np.cumsum(counts, out=comp_start[1:])
    # schedule of the component-major pass: the CTA's 4 warps take entries w, w + 4, ... of comp_order; deal the
    # components out by decreasing cost (summed window lengths) in serpentine order so that the warps finish together
    lengths = (window >> 16).astype(np.int64) - (window & 0xFFFF).astype(np.int64)
    cost = np.zeros(len(names), dtype=np.int64)
    for i, line in enumerate(lines):
        cost[line[2]] += lengths[i] + 16  # + per-line overhead
    ranked = sorted(range(len(names)), key=lambda c: (-int(cost[c]), c))
    comp_order = np.zeros(len(names), dtype=np.int32)
    for r0 in range(0, len(ranked), REFINE_WARPS):
        chunk = ranked[r0: r0 + REFINE_WARPS]
        if (r0 // REFINE_WARPS) % 2:
            chunk = chunk[::-1] if len(chunk) == REFINE_WARPS else chunk
        comp_order[r0: r0 + len(chunk)] = chunk
    return cfg, arr, chan_lines, comp_start, comp_lines, window, comp_order


def model_jacobian(params, energy_range, names, amps, free, e_consts=default_energy_consts, elem_info=default_elem_info):
    """Model spectrum of model_spec's path and its analytic derivatives w.r.t. the calibration parameters `free` (any of
    _native.THETA_IDS, up to 8) at fixed linear amplitudes `amps` (CUDA fp32 [len(names)]): what torch.autograd gives
    the reference for d model_spec / d param (opt.py:249-317).  Returns (model [C], d_theta [len(free), C]) on the device."""
    er = (int(energy_range[0]), int(energy_range[1]))
    # windows wide enough for exact values everywhere that matters: no bounds (the parameters do not move here)
    zero = {k: ("abs", 0.0) for k in DEFAULT_BOUNDS}
    cfg, lines, chan_lines, _, _, _, _ = refine_tables(params, names, er, free, zero, k_sigma=7.0,
                                                       max_theta=_native.MT_MAX_THETA, e_consts=e_consts, elem_info=elem_info)
    n_ch = er[1] - er[0] + 1
    x = amps.detach().to("cuda", torch.float32).contiguous()
    model = torch.empty(n_ch, dtype=torch.float32, device="cuda")
    d = torch.empty((max(len(free), 1), n_ch), dtype=torch.float32, device="cuda")
    _native.call("mt_model_jacobian", ctypes.byref(cfg), ctypes.cast(lines, ctypes.c_void_p),
                 chan_lines.ctypes.data_as(ctypes.c_void_p), x.data_ptr(), model.data_ptr(), d.data_ptr(), d.stride(0),
                 _native.stream_ptr())
    return model, d[: len(free)]


def refine_map(plan, spec2d, x, free=("ENERGY_OFFSET", "FWHM_OFFSET"), bounds=None, max_iter=12, tol=1e-6,
               bkg=None, stream=None):
    """Refine amplitudes x [E, P] (CUDA, from plan.fit) and the `free` parameters per pixel in place.

    Returns (theta [len(free), P], loss [P], iterations [P]) on the device."""
    if plan.n_live != len(plan.names):
        raise NotImplementedError("refinement with components that have no open line")
    # A channel mask (fit_spec's `indices`) enters the kernel as 0 / 1 weights of the residual and derivative rows, escape
    # peaks (escape_factor > 0) as the shifted, scaled copy every column carries (map.py:127-150; the bin shift of the
    # plan's calibration); the plan's G^-1 (the fixed amplitude curvature) was built from that same masked / augmented basis
    escape_factor = float(getattr(plan, "escape_factor", 0.0))
    escape_bins = 0
    if escape_factor > 0:
        from mapstorch_b200 import tables

        escape_bins = int(tables.escape_bins(tables.energy_axis(plan.params, plan.er), getattr(plan, "detector_type", "Si")))
    free = list(free)
    unknown = [f for f in free if f not in _native.THETA_IDS]
    if unknown or len(free) > _native.MT_MAX_THETA_REFINE:
        raise ValueError(f"free parameters: at most {_native.MT_MAX_THETA_REFINE} of {sorted(_native.THETA_IDS)} per pixel, "
                         f"got {free}")
    # the host tables depend only on the plan and the free parameters: built once, reused for every slab
    key = (tuple(free), tuple(sorted((bounds or {}).items())))
    cache = plan.__dict__.setdefault("_refine_tables", {})
    if key not in cache:
        cache[key] = refine_tables(plan.params, plan.names, plan.er, free, bounds,
                                   e_consts=getattr(plan, "e_consts", default_energy_consts),
                                   elem_info=getattr(plan, "elem_info", default_elem_info))
    cfg, lines, chan_lines, comp_start, comp_lines, window, comp_order = cache[key]
    cfg.max_iter, cfg.tol = int(max_iter), float(tol)
    n_px = spec2d.shape[0]
    if not hasattr(plan, "_g0inv32"):
        plan._g0inv32 = torch.from_numpy(plan.hinv.astype(np.float32)).cuda()
    theta = torch.empty((max(len(free), 1), n_px), dtype=torch.float32, device="cuda")
    loss = torch.empty(n_px, dtype=torch.float32, device="cuda")
    iters = torch.empty(n_px, dtype=torch.int32, device="cuda")
    if bkg is not None and (bkg.dtype != torch.float32 or bkg.stride(1) != 1):
        bkg = bkg.float().contiguous()
    _native.call("mt_refine", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), n_px, spec2d.stride(0),
                 bkg.data_ptr() if bkg is not None else None, bkg.stride(0) if bkg is not None else 0,
                 plan.chan_mask.data_ptr() if plan.chan_mask is not None else None, escape_bins, escape_factor,
                 ctypes.byref(cfg), ctypes.cast(lines, ctypes.c_void_p), chan_lines.ctypes.data_as(ctypes.c_void_p),
                 comp_start.ctypes.data_as(ctypes.c_void_p), comp_lines.ctypes.data_as(ctypes.c_void_p),
                 window.ctypes.data_as(ctypes.c_void_p), comp_order.ctypes.data_as(ctypes.c_void_p),
                 plan._g0inv32.data_ptr(), x.data_ptr(), x.stride(0),
                 theta.data_ptr(), theta.stride(0), loss.data_ptr(), iters.data_ptr(), _native.stream_ptr(stream))
    return theta[: len(free)], loss, iters


# ---- fit_spec(tune_params=True): global calibration fit of one spectrum -------------------------
LM_REFRESH = 5  # accepted steps between full forward differencings (Broyden updates in between)
LAST_STATS = {"evaluations": 0, "snip_launches": 0, "jacobians": 0}  # work done by the last fit_spec_global call


def fit_spec_global(int_spec, energy_range, elements_to_fit, fitting_params, init_param_vals, fixed_param_vals,
                    indices, use_snip, n_iter, e_consts, elem_info, detector_type, escape_factor, device, verbose,
                    loss="mse", l1_lambda=0.0):
    """``fit_spec`` with ``tune_params=True`` (opt.py:165-327) on the GPU.

    The reference runs Adam on all tunable calibration parameters and log-amplitudes at once.  Here the
    amplitudes are projected out (for given calibration parameters they are the non-negative
    least-squares solution) and the remaining small problem is solved by
    Levenberg-Marquardt.  Every evaluation runs K1 (basis), K2 (SNIP, which depends on the calibration too) and
    the non-negative solve on the device.  Without SNIP the Jacobian is ANALYTIC -- one launch of
    mt_model_jacobian (the derivative code of the per-pixel refinement kernel, checked against torch.autograd of
    the reference) gives d model / d parameter for all free parameters, projected orthogonally to the unclamped
    basis columns (Kaufman's variable-projection Jacobian) -- so an iteration costs one derivative launch plus the
    trial evaluations instead of 1 + len(free) evaluations.  With SNIP, escape peaks or the l1_lambda penalty the
    residual is differenced as a whole (forward differences).

    ``loss="l1"`` wraps that in iteratively re-weighted rounds (weights 1/max(|res|, eps), eps annealed);
    the ``l1_lambda`` penalty c*sum(A) enters the amplitude solve as a shifted right-hand side and the
    Levenberg-Marquardt residual vector as extra entries sqrt(c*A) (A >= 0), see robust.py."""
    from mapstorch_b200 import bkg as _bkg, map as _map, robust
    from mapstorch_b200.default import default_fitting_params, default_learning_rates
    from mapstorch_b200.opt import create_tensors

    _native.require_device()
    er = (int(energy_range[0]), int(energy_range[1]))
    spec = int_spec.detach().float().cpu().numpy() if isinstance(int_spec, torch.Tensor) else np.asarray(int_spec, dtype=np.float32)
    y = torch.from_numpy(np.ascontiguousarray(spec[er[0]: er[1] + 1])).cuda()
    params = dict(default_param_vals)
    params.update({k: (v.item() if isinstance(v, torch.Tensor) else float(v)) for k, v in init_param_vals.items()
                   if k in default_param_vals})
    params.update({k: (v.item() if isinstance(v, torch.Tensor) else float(v)) for k, v in fixed_param_vals.items()
                   if k in default_param_vals})
    free = [p for p in default_fitting_params if p in set(fitting_params) and p not in fixed_param_vals]
    elements = list(dict.fromkeys(list(elements_to_fit)))
    idx = None if indices is None else torch.as_tensor(np.asarray(indices), device="cuda")
    scale = np.array([10.0 * default_learning_rates.get(p, default_learning_rates["default_lr"]) for p in free])
    robust_loss = loss == "l1"
    bkg_cache = {}
    LAST_STATS.update(evaluations=0, snip_launches=0, jacobians=0)

    if idx is not None and idx.dtype == torch.bool:
        idx = idx.nonzero().flatten()
    idx32 = None if idx is None else idx.to(torch.int32).contiguous()
    n_k = y.numel() if idx is None else int(idx.numel())
    no_bkg = torch.zeros_like(y)
    warm = {}

    def dense(info):
        """fp64 basis and target on the fitted channels (the analytic Jacobian and the final polish want them)."""
        if "bm" not in info:
            target = (y - info["bkg"]).double()
            info["bm"] = info["basis"].double() if idx is None else info["basis"].double()[:, idx]
            info["tm"] = target if idx is None else target[idx]
        return info["bm"], info["tm"]

    def evaluate(theta, w=None, defer=None):
        """Residual vector and solution at `theta`.  defer: a list -- the evaluation's device-side statistics are appended to it
        instead of being read back (no synchronisation; the trial points of a forward-difference Jacobian)."""
        LAST_STATS["evaluations"] += 1
        pp = dict(params, **dict(zip(free, (float(t) for t in theta))))
        basis, names = _map.element_basis(pp, er, elements, e_consts, elem_info, detector_type, escape_factor)
        if use_snip:
            # the background depends on six of the parameters only: trial points that move the others (scatter energy,
            # Compton angle and width) reuse it instead of rebuilding the window tables and relaunching K2
            key = tuple(float(pp[k]) for k in (
                "ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME"))
            bkg = bkg_cache.get(key)
            if bkg is None:
                if len(bkg_cache) > 64:
                    bkg_cache.clear()
                LAST_STATS["snip_launches"] += 1
                bkg = bkg_cache[key] = _bkg.snip_bkg(y, er, *(torch.tensor(v) for v in key), device="cuda")
        else:
            bkg = no_bkg
        if basis.shape[0] <= _native.MT_SPEC_MAX_COMP and n_k <= _native.MT_SPEC_MAX_CH:
            # one launch (csrc/mt_spec.cu): Gram matrix, non-negative solve, residual vector, objective; one 64-byte
            # read-back per trial point
            if warm.get("n") != basis.shape[0]:  # the active set of the last trial point starts the next solve
                warm["n"], warm["set"] = basis.shape[0], torch.full((1, basis.shape[0]), 255, dtype=torch.uint8, device="cuda")
            x, res, vec, stats = robust.spec_solve(basis, y, bkg if use_snip else None, idx32, w, l1_lambda, robust_loss,
                                                   len(names), free_set=warm["set"])
            if defer is not None:
                defer.append(stats)
                return (vec[0] if l1_lambda > 0 else vec[0, :n_k]), None
            st = stats[0].tolist()
            if st[3] == 2.0:
                st[0] = st[1] = float("inf")  # collinear components at this trial point: never accepted as a step
            return (vec[0] if l1_lambda > 0 else vec[0, :n_k]), dict(
                x=x[0], names=names, basis=basis, bkg=bkg, params=pp, res=res[0], c=st[2], objective=st[1], value=st[0])
        info = dict(names=names, basis=basis, bkg=bkg, params=pp)
        bm, tm = dense(info)
        c = robust.penalty_constant(tm, loss, l1_lambda, len(names))
        x = robust.weighted_nnls(bm, tm, w, c if robust_loss else 0.5 * c)
        res = x @ bm - tm
        vec = res if w is None else torch.sqrt(w) * res
        if l1_lambda > 0:
            vec = torch.cat([vec, torch.sqrt((2.0 if robust_loss else 1.0) * c * x)])
        info.update(x=x, res=res, c=c, objective=robust.objective(bm, tm, x, loss, c), value=float(vec @ vec))
        return vec, info

    # Parameters that move the model: the tail fractions of default_fitting_params never enter model_spec
    # (use_tail=False, map.py:290,302), their gradient is zero in the reference as well and they keep their values.
    active = [k for k, name in enumerate(free) if name in _native.THETA_IDS]
    # what a finite difference can see at all: the model parameters above plus the SNIP width
    moving = set(active) | {k for k, name in enumerate(free) if name == "SNIP_WIDTH" and use_snip}
    # analytic Jacobian (mt_model_jacobian) wherever the device model is the whole model: no escape peaks, no penalty
    # entries in the residual vector; otherwise forward differences (1 + len(free) evaluations per iteration)
    # ... and no SNIP: the background follows the width and energy-axis parameters through integer window tables, a
    # dependence of the objective that only a finite difference of the WHOLE residual sees (measured on BASELINE
    # config 1: final loss 7.6e5 with forward differences, 1.4e6 with the model-only Jacobian, reference 9.3e5)
    analytic = (escape_factor == 0.0 and not (l1_lambda > 0) and len(active) <= _native.MT_MAX_THETA and not use_snip)

    state = {"fd": not analytic}

    def jacobian(theta, w, resid, info):
        if not state["fd"]:
            return jacobian_analytic(theta, w, resid, info)
        # forward differences: 1 + len(moving) evaluations.  Parameters that never enter model_spec or the background
        # (the tail fractions, see `active`) have an identically zero column: no evaluation is spent on them.
        cols = []
        zero = None
        pending = []  # device-side statistics of the trial points: nothing is read back until all of them are queued
        for k in range(len(free)):
            if k not in moving:
                zero = torch.zeros_like(resid) if zero is None else zero
                cols.append(zero)
                continue
            t2 = theta.copy()
            t2[k] += 0.05 * scale[k]
            r2, _ = evaluate(t2, w, defer=pending)
            cols.append((r2 - resid) / 0.05)  # derivative w.r.t. the scaled variable theta_k / scale_k
        jac = torch.stack(cols, dim=1)
        if pending and bool((torch.cat(pending)[:, 3] == 2.0).any()):
            # a trial point with collinear components has no usable residual: its column would be garbage
            raise RuntimeError("fit_spec: the Gram matrix of the model components is singular next to "
                               f"{dict(zip(free, theta))}")
        return jac

    def jacobian_analytic(theta, w, resid, info):
        # variable projection (Kaufman): d r / d theta_k = P (d model / d theta_k at fixed amplitudes), P the projector
        # orthogonal to the basis columns of the components that are not clamped (weighted for the robust loss)
        x = info["x"]
        _, d = model_jacobian(info["params"], er, info["names"], x.float(), [free[k] for k in active], e_consts, elem_info)
        d = d.double() if idx is None else d.double()[:, idx]
        bm, tm = dense(info)
        bf = bm[x > 0]
        wv = torch.ones_like(tm) if w is None else w
        if bf.shape[0]:
            gram = (bf * wv) @ bf.T
            coef = torch.linalg.solve(gram + 1e-12 * torch.diag(torch.diagonal(gram)), (bf * wv) @ d.T)
            d = d - coef.T @ bf
        jac = torch.zeros((resid.shape[0], len(free)), dtype=torch.float64, device=d.device)
        for j, k in enumerate(active):
            jac[:, k] = torch.sqrt(wv) * d[j] * scale[k]
        return jac

    # physical ranges (the optimiser never leaves them) and a trust region in the scaled variables: one iteration
    # moves a parameter by at most `radius` x 10 learning rates (default.py:102-116) -- what the reference's Adam
    # covers in ten steps.  Exact derivatives make the raw Gauss-Newton step long on the first, badly conditioned
    # iterations (a Compton angle 70 degrees away was accepted once: lower loss, wrong basin).
    limits = {"ENERGY_SLOPE": (1e-6, np.inf), "FWHM_OFFSET": (1e-4, np.inf), "FWHM_FANOPRIME": (1e-9, np.inf),
              "COMPTON_ANGLE": (1.0, 179.0), "COMPTON_FWHM_CORR": (0.1, 20.0), "COHERENT_SCT_ENERGY": (0.5, 300.0)}
    lo = np.array([limits.get(name, (-np.inf, np.inf))[0] for name in free])
    hi = np.array([limits.get(name, (-np.inf, np.inf))[1] for name in free])

    def run_lm(theta, w, max_it, trace):
        resid, info = evaluate(theta, w)
        value = info["value"]
        lam = 1e-3
        radius = 1.0
        # Forward differences cost one evaluation per moving parameter.  Between full differencings the Jacobian is
        # carried along by Broyden's rank-one secant update (free: it uses the accepted trial point), refreshed every
        # `refresh` accepted steps and whenever a step computed from a carried Jacobian fails.
        jac, age, refresh = None, 0, LM_REFRESH
        it = 0
        while it < (max_it if free else 0):
            fresh = jac is None or not state["fd"] or age >= refresh
            if fresh:
                LAST_STATS["jacobians"] += 1
                jac = jacobian(theta, w, resid, info)
                age = 0
            jtj = (jac.T @ jac).cpu().numpy()
            jtr = (jac.T @ resid).cpu().numpy()
            improved = False
            for _ in range(8 if fresh else 3):
                try:
                    damp = np.diag(np.diag(jtj)) + 1e-12 * np.eye(len(free))
                    dead = np.diag(jtj) <= 0  # parameters without influence (zero column): no step
                    step = np.linalg.solve(jtj + lam * damp + np.diag(dead.astype(float)), -jtr)
                except np.linalg.LinAlgError:
                    lam *= 10
                    continue
                longest = float(np.max(np.abs(step))) if step.size else 0.0
                clipped = longest > radius and not state["fd"]  # trust region: analytic mode only
                if clipped:
                    step = step * (radius / longest)
                cand = np.clip(theta + step * scale, lo, hi)
                r_new, info_new = evaluate(cand, w)
                value_new = info_new["value"]
                if value_new < value:
                    rel = (value - value_new) / max(value, 1e-30)
                    if state["fd"] and r_new.shape == resid.shape:
                        sv = torch.from_numpy((cand - theta) / scale).to(jac)
                        denom = float(sv @ sv)
                        if denom > 0:
                            jac = jac + torch.outer(r_new - resid - jac @ sv, sv) / denom
                    theta, resid, info, value = cand, r_new, info_new, value_new
                    lam = max(lam / 5, 1e-9)
                    if clipped:
                        radius = min(radius * 2.0, 16.0)
                        rel = max(rel, 1.0)  # a clipped step says nothing about convergence
                    improved = True
                    age += 1
                    break
                lam *= 5
                radius = max(radius * 0.5, 1e-3)
            if (not improved or rel < 1e-10) and not fresh:
                jac = None  # the carried Jacobian has gone stale: difference again before drawing conclusions
                lam = min(lam, 1e-3)
                continue
            it += 1
            trace.append(info["objective"])
            if verbose:
                print(f"fit_spec_global iter {it}: objective {info['objective']:.6g} lambda {lam:.2g}")
            if (not improved or rel < 1e-10) and not state["fd"]:
                # The analytic Jacobian leaves out what is not differentiable: the SNIP background follows the width
                # and energy-axis parameters through integer window tables, and the non-negative solve changes its
                # active set.  Where that stalls the iteration, finish with forward differences of the whole residual.
                state["fd"] = True
                lam, radius = 1e-3, 1.0
                jac = None
                continue
            if not improved or rel < 1e-10:
                break
        return theta, info

    theta = np.array([params[p] for p in free], dtype=np.float64)
    trace = []
    budget = max(1, min(int(n_iter), 60))
    # The host side of an evaluation is a few dozen torch CPU ops on ~1400-element tensors (the reference's own op
    # sequence for the SNIP window tables).  With a large intra-op thread pool every one of them pays a fork / join:
    # 0.29 s per call became 0.88 s inside a process that had set 32 threads.  One thread for the duration of the fit.
    host_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        if not robust_loss:
            _, info0 = evaluate(theta)
            trace.append(info0["objective"])
            theta, info = run_lm(theta, None, budget, trace)
        else:
            _, info = evaluate(theta)
            trace.append(info["objective"])
            top = max(float(info["res"].abs().max()), 1e-30)
            eps, floor = 1e-3 * top, 1e-10 * top
            for rnd in range(16):
                before = info["objective"]
                theta, info = run_lm(theta, robust.irls_weights(info["res"], eps), min(budget, 12), trace)
                if eps <= floor and abs(before - info["objective"]) <= 1e-8 * abs(before):
                    break
                eps = max(eps * 0.2, floor)
            # calibration settled: converge the amplitudes of the final model all the way
            bm, tm = dense(info)
            polished = robust.solve_amplitudes(bm, tm, loss, info["c"])
            value = robust.objective(bm, tm, polished, loss, info["c"])
            if value <= info["objective"]:
                info = dict(info, x=polished, objective=value)
                trace.append(value)

        if not np.isfinite(info["objective"]):
            raise RuntimeError("fit_spec: the Gram matrix of the model components is singular (collinear components) at "
                               f"{ {k: info['params'][k] for k in free} }")
        pp, names, x = info["params"], info["names"], info["x"]
        tensors, _ = create_tensors(names, free, pp, {k: pp[k] for k in pp if k not in free}, tune_params=True, device=device)
        for n, a in zip(names, x.float()):
            tensors[n] = torch.log10(torch.clamp(a, min=1e-30)).to(device).requires_grad_(True)
        spec_fit = (x.float() @ info["basis"]).reshape(-1)
        return tensors, spec_fit.cpu().numpy(), info["bkg"].cpu().numpy(), trace
    finally:
        torch.set_num_threads(host_threads)
