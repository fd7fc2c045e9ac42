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


def spec_solve(basis, y, bkg=None, chan_idx=None, weights=None, l1_lambda=0.0, robust=False, n_elements=1, free_set=None):
    """The amplitude solve of one spectrum (or a batch of trial points) in ONE launch of mt_spec_solve
    (csrc/mt_spec.cu): Gram matrix, non-negative solve, residual vector and objective, no host synchronisation.

    basis fp32 CUDA [E, C] or [N, E, C]; y fp32 CUDA [C]; bkg fp32 CUDA [C] / [N, C] / None; chan_idx int32 CUDA [K] or
    None (all channels); weights fp64 CUDA [K] / [N, K] or None; free_set uint8 CUDA [N, E] or None: starting partition of
    the pivoting, overwritten with the final one (warm start between neighbouring trial points).  Returns (x [N, E], res [N, K], vec [N, K + E],
    stats [N, 8]) as fp64 CUDA tensors, see include/mapstorch_b200.h."""
    b3 = basis if basis.dim() == 3 else basis.unsqueeze(0)
    if b3.stride(2) != 1 or b3.stride(1) < b3.shape[2]:
        b3 = b3.contiguous()
    n_prob, n_comp, n_ch = b3.shape
    n_k = int(chan_idx.numel()) if chan_idx is not None else n_ch
    if n_comp > _native.MT_SPEC_MAX_COMP or n_k > _native.MT_SPEC_MAX_CH or n_k < 1:
        raise _native.NativeError(_native.MT_ERR_UNSUPPORTED, f"mt_spec_solve: {n_comp} components x {n_k} channels")
    ld_bkg = 0
    if bkg is not None:
        bkg = bkg.contiguous()
        ld_bkg = bkg.stride(0) if bkg.dim() == 2 and bkg.shape[0] > 1 else 0
    ld_w = 0
    if weights is not None:
        weights = weights.contiguous()
        ld_w = weights.stride(0) if weights.dim() == 2 and weights.shape[0] > 1 else 0
    out = torch.empty((n_prob, n_comp + n_k + (n_k + n_comp) + 8), dtype=torch.float64, device=b3.device)
    x, res, vec, stats = torch.split(out, [n_comp, n_k, n_k + n_comp, 8], dim=1)
    # outputs are rows of one allocation: per-problem pitches differ from the packed layout the entry point expects
    # only for n_prob > 1, so batches get their own packed tensors
    if n_prob > 1:
        x, res, vec, stats = (torch.empty((n_prob, n), dtype=torch.float64, device=b3.device)
                              for n in (n_comp, n_k, n_k + n_comp, 8))
    _native.call("mt_spec_solve", b3.data_ptr(), b3.stride(1), b3.stride(0) if n_prob > 1 else 0, n_comp, y.data_ptr(),
                 bkg.data_ptr() if bkg is not None else None, ld_bkg,
                 chan_idx.data_ptr() if chan_idx is not None else None, n_k,
                 weights.data_ptr() if weights is not None else None, ld_w, float(l1_lambda), int(bool(robust)),
                 int(n_elements), x.data_ptr(), res.data_ptr(), vec.data_ptr(), stats.data_ptr(),
                 free_set.data_ptr() if free_set is not None else None, n_prob, _native.stream_ptr())
    return x, res, vec, stats


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


# ---- per-pixel loss="l1" maps ---------------------------------------------------------------------------------------
# ADMM schedule: (iterations, factor applied to the threshold kappa at the start of the stage).  A large threshold moves
# the primal variables fast, a small one resolves the optimum; the scaled dual follows the threshold (u <- u * factor).
# Measured in float64 against the exact LP on 64 pixels with outliers (worst / median relative gap of the objective):
# plain 1.9e-6 / 1.2e-7, SNIP residuals 4.7e-6 / 1.5e-7, peak-window masks 6.7e-6 / 3.2e-7.  The slowest pixels need the
# long first stage at a LARGE threshold (3 x the mean residual): with 1 x they stall at 3e-5 whatever comes after.
LAD_SCHEDULE = ((200, 1.0), (100, 1.0 / 3.0), (100, 1.0 / 3.0), (100, 1.0 / 3.0))
LAD_ITERATIONS = sum(n for n, _ in LAD_SCHEDULE)
LAD_KAPPA = 3.0       # first-stage threshold per pixel = LAD_KAPPA x mean |residual| of the least-squares start
LAD_CHUNK_PX = 32768  # t, u and the operand planes cost 16 bytes per (pixel, channel)


def fit_map_lad(plan, spec2d, nonneg=True, schedule=LAD_SCHEDULE, kappa=LAD_KAPPA, chunk_px=LAD_CHUNK_PX, l1_lambda=0.0,
                loader=None, n_px=None):
    """argmin_{x >= 0} sum_c |x . B[:, c] + bkg_c - y_c| for every pixel of spec2d [P, C_full] (CUDA or host; the
    reference's loss="l1" with the calibration fixed, opt.py:232-233, 371-375; restricted to plan's `indices`).

    ADMM on z = model - target: the x-update is the map fit itself -- mt_fit_contract with the pixel-independent
    pseudo-inverse + the batched non-negative fix-up, applied to the modified target t + z - u (incrementally: the
    contraction sees the change of the target, the unconstrained solution is accumulated) -- and mt_lad_update
    does the rest in one pass (model, soft threshold, dual update, operand split, L1 objective of the current x).
    The best iterate per pixel is kept; `schedule` lists (iterations, threshold factor) stages.  Returns (x [E, P] fp32, objective [P]) on the device.

    l1_lambda > 0 adds the reference's amplitude penalty c_p * sum(x) with c_p = l1_lambda * sum|y - bkg| / n_elements per
    pixel (opt.py:372-381 with the L1 loss).  On x >= 0 the penalty is linear, so it only moves the unconstrained
    solution of the x-update: with the threshold kappa = 1 / rho the update minimises c sum(x) + |B^T x - v|^2 / (2 kappa),
    i.e. x = nonneg(xu - c_p kappa_p G^-1 1).  The reported objective includes the penalty.

    loader (with n_px): a slab source of opt.fit_map -- loader(p0, n, dst) fills dst [n, C_full] on the device with pixels
    p0 .. p0 + n (whole rows) -- instead of spec2d: lazily read files and channel-first volumes, chunk by chunk."""
    import ctypes  # noqa: F401

    from mapstorch_b200 import bkg as _bkg

    n_px, n_all = (int(n_px) if loader is not None else spec2d.shape[0]), len(plan.names)
    n_ch, er0 = plan.n_ch, plan.er[0]
    live = torch.as_tensor(plan.live, device="cuda")
    basis = plan.basis[live].contiguous()
    if plan.chan_mask is not None:
        basis = (basis * plan.chan_mask[None, :]).contiguous()
    n_active = float(plan.chan_mask.sum()) if plan.chan_mask is not None else float(n_ch)
    out = torch.zeros((n_all, n_px), dtype=torch.float32, device="cuda")
    objective = torch.zeros(n_px, dtype=torch.float32, device="cuda")
    chunk = max(1, min(int(chunk_px), n_px))
    slab = None
    if loader is not None:
        # chunks of whole loader slabs (which are whole rows of the map)
        chunk = min(max(1, int(chunk_px) // loader.chunk_px) * loader.chunk_px, -(-n_px // loader.chunk_px) * loader.chunk_px)
        slab = torch.empty((chunk, loader.n_ch), dtype=loader.dtype, device="cuda")
    pitch = -(-n_ch // 4) * 4
    t = torch.zeros((chunk, pitch), dtype=torch.float32, device="cuda")
    u = torch.zeros((chunk, pitch), dtype=torch.float32, device="cuda")
    qs = torch.zeros((chunk, pitch), dtype=torch.float32, device="cuda")
    q = torch.zeros((chunk, 2 * pitch), dtype=torch.float32, device="cuda")[:, : 2 * n_ch]
    stream = _native.stream_ptr()
    penalised = float(l1_lambda) > 0
    if penalised and not nonneg:
        raise ValueError("the l1_lambda penalty needs the non-negative solve")
    hinv1 = torch.from_numpy(plan.hinv.sum(axis=1).astype("float32")).cuda() if penalised else None  # (G^-1 1), live components

    def update(tn, xn, un, kap, qsn, qn, obj, mode):
        ptr = lambda a: a.data_ptr() if a is not None else None  # noqa: E731
        ld = lambda a: a.stride(0) if a is not None else 0  # noqa: E731
        _native.call("mt_lad_update", tn.data_ptr(), tn.stride(0), xn.data_ptr(), xn.stride(0), basis.data_ptr(),
                     basis.stride(0), plan.n_live, n_ch, tn.shape[0], ptr(un), ld(un), ptr(kap), ptr(qsn), ld(qsn), ptr(qn),
                     ld(qn), obj.data_ptr(), int(mode), stream)

    def nonneg_copy(xu):
        x = xu.clone()
        if nonneg:
            _native.call("mt_nnls_fixup", x.data_ptr(), x.stride(0), _native.MT_X_ELEMENT_MAJOR, x.shape[1], plan.n_live,
                         plan.gram_dev.data_ptr(), plan.hinv_dev.data_ptr(), None, None, stream)
        return x

    for p0 in range(0, n_px, chunk):
        n = min(chunk, n_px - p0)
        if loader is not None:
            for q0 in range(0, n, loader.chunk_px):
                m = min(loader.chunk_px, n - q0)
                loader(p0 + q0, m, slab[q0: q0 + m])
            part = slab[:n]
        else:
            part = spec2d[p0: p0 + n]
            if part.device.type != "cuda":
                part = part.cuda(non_blocking=True)
        tn, un, qsn, qn = t[:n], u[:n], qs[:n], q[:n]
        # target t = y - bkg on the fitted window
        if plan.use_snip:
            _bkg.snip_launch(part, er0, n_ch, plan._snip[0], plan._snip[1], plan.boxcar_size, tn, _native.MT_SNIP_RESIDUAL)
        else:
            tn[:, :n_ch] = part[:, er0: er0 + n_ch]
        if plan.chan_mask is not None:
            tn[:, :n_ch] *= plan.chan_mask[None, :]
        un.zero_()
        qsn.copy_(tn)
        # least-squares start: xu = unconstrained solution of the target itself (the exact integer-count path where
        # it applies), x = its non-negative fix-up
        xu = plan.fit(part, nonneg=False)[live].contiguous()
        dxu = torch.empty_like(xu)
        x = nonneg_copy(xu)
        obj = torch.empty(n, dtype=torch.float32, device="cuda")
        update(tn, x, None, None, None, None, obj, 0)
        kap = (float(kappa) / n_active) * obj
        pen = (float(l1_lambda) / max(n_all, 1)) * tn[:, :n_ch].abs().sum(dim=1) if penalised else None  # c_p

        def solve(kap_now):
            if pen is None:
                return nonneg_copy(xu)
            return nonneg_copy(xu - hinv1[:, None] * (pen * kap_now)[None, :])

        def total(x_now):  # the kernel's L1 objective of x_now (in obj) plus the penalty
            if pen is not None:
                obj.add_(pen * x_now.sum(dim=0))
            return obj

        total(x)
        best_obj, best_x = obj.clone(), x.clone()
        for n_stage, factor in schedule:
            if factor != 1.0:
                un.mul_(factor)
                kap = kap * factor
            for _ in range(int(n_stage)):
                update(tn, x, un, kap, qsn, qn, obj, 1)  # obj = objective of the x that enters
                total(x)
                better = obj < best_obj
                best_obj = torch.where(better, obj, best_obj)
                best_x = torch.where(better[None, :], x, best_x)
                plan.fit_operand_hilo(qn, dxu, nonneg=False)
                xu += dxu
                x = solve(kap)
        update(tn, x, None, None, None, None, obj, 0)
        total(x)
        better = obj < best_obj
        best_obj = torch.where(better, obj, best_obj)
        best_x = torch.where(better[None, :], x, best_x)
        out[live, p0: p0 + n] = best_x
        objective[p0: p0 + n] = best_obj
    return out, objective
