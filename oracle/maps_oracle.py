"""CPU oracle of the per-pixel XRF map-fitting hot path.  TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference algorithm (xyin-anl/MapsTorch), written from the
reference sources; every function cites the file:line it follows.  It exists so that the
CUDA path can be checked on the GPU box, where /root/reference does not exist.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package ``mapstorch_b200`` never does.

Parity pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks oracle == golden on every CPU run).

The element dictionaries come from ``mapstorch_b200.default`` (host-side data, themselves
verified field-by-field against the reference dump in ``tests/test_dictionaries.py``).
"""
import math

import numpy as np
from scipy.special import erfc as _erfc

f32 = np.float32
M_PI = math.pi
M_SQRT2 = math.sqrt(2)
SQRT_2XPI = math.sqrt(2 * math.pi)
SCATTER = ("COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE")


def _defaults():
    from mapstorch_b200 import default

    return default


def p32(params, key):
    v = params[key]
    return f32(v.item() if hasattr(v, "item") else v)


# ---- forward model ---------------------------------------------------------------------------
def energy_axis(params, energy_range):
    """map.py:271-281: ev = off + slope*c + quad*c**2, all fp32, evaluated left to right."""
    c = np.arange(energy_range[0], energy_range[1] + 1, dtype=np.float32)
    return (p32(params, "ENERGY_OFFSET") + p32(params, "ENERGY_SLOPE") * c) + p32(params, "ENERGY_QUADRATIC") * (c * c)


def erfc32(x):
    return _erfc(x.astype(np.float64)).astype(np.float32)


def peak(gain, sigma, delta):
    """map.py:57-58."""
    q = delta / sigma
    return (gain / (sigma * f32(SQRT_2XPI))) * np.exp(f32(-0.5) * (q * q))


def step(gain, sigma, delta, peak_e):
    """map.py:61-62."""
    return (gain / f32(2.0) / f32(peak_e)) * erfc32(delta / (f32(M_SQRT2) * sigma))


def tail(gain, sigma, delta, gamma):
    """map.py:65-69."""
    v1 = erfc32(delta / (f32(M_SQRT2) * sigma) + (f32(1.0) / (gamma * f32(M_SQRT2))))
    with np.errstate(over="ignore", invalid="ignore"):
        v2 = np.exp(delta / (gamma * sigma)) * v1
    d_e = np.where(delta < 0, v2, v1)
    norm = np.exp((f32(1.0) / (gamma * gamma)) * f32(-0.5))
    return (gain / f32(2.0) / gamma / sigma / norm) * d_e


def _sigma(params, e_times_296):
    a = p32(params, "FWHM_OFFSET") / f32(2.3548)
    return np.sqrt(a * a + e_times_296 * p32(params, "FWHM_FANOPRIME"))


def compton_energy(params):
    """map.py:83-87."""
    e0 = p32(params, "COHERENT_SCT_ENERGY")
    ang = p32(params, "COMPTON_ANGLE") * f32(2) * f32(M_PI) / f32(360.0)
    return e0 / (f32(1) + (e0 / f32(511)) * (f32(1) - np.cos(ang)))


def elastic_peak(params, ev, gain):
    """map.py:72-79."""
    e0 = p32(params, "COHERENT_SCT_ENERGY")
    sigma = _sigma(params, e0 * f32(2.96))
    amp = np.power(f32(10), p32(params, "COHERENT_SCT_AMPLITUDE"))
    return amp * peak(f32(gain), sigma, ev - e0)


def compton_peak(params, ev, gain, use_step=True, use_tail=False):
    """map.py:82-124."""
    gain = f32(gain)
    ce = compton_energy(params)
    sigma = _sigma(params, ce * f32(2.96))
    delta = ev - ce
    f_step, f_tail, hi_tail = (p32(params, k) for k in ("COMPTON_F_STEP", "COMPTON_F_TAIL", "COMPTON_HI_F_TAIL"))
    norm = f32(1.0)
    if use_step and f_step > 0:
        norm = norm + f_step
    if use_tail:
        if f_tail > 0:
            norm = norm + f_tail
        if hi_tail > 0:
            norm = norm + hi_tail
    faktor = np.power(f32(10), p32(params, "COMPTON_AMPLITUDE")) / norm
    counts = faktor * peak(gain, sigma * p32(params, "COMPTON_FWHM_CORR"), delta)
    if use_step and f_step > 0:
        counts = counts + (faktor * f_step) * step(gain, sigma, delta, ce)
    if use_tail:
        if f_tail > 0:
            counts = counts + (faktor * f_tail) * tail(gain, sigma, delta, p32(params, "COMPTON_GAMMA"))
        if hi_tail > 0:
            counts = counts + (faktor * hi_tail) * tail(gain, sigma, -delta, p32(params, "COMPTON_HI_GAMMA"))
    return counts.astype(np.float32)


_TAIL_LINES = ("Ka1", "Ka2", "La1", "La2", "Lb1", "Lb2", "Lb3", "Lb4", "Lg1", "Lg2", "Lg3", "Lg4", "Ll", "Ln")


def model_elem_spec(params, element, ev, use_step=True, use_tail=True, e_consts=None, elem_info=None):
    """map.py:152-257, fp32, same accumulation order."""
    d = _defaults()
    e_consts = d.default_energy_consts if e_consts is None else e_consts
    elem_info = d.default_elem_info if elem_info is None else elem_info
    gain = p32(params, "ENERGY_SLOPE")
    e0 = p32(params, "COHERENT_SCT_ENERGY")
    spec = np.zeros_like(ev, dtype=np.float32)
    pre = np.power(f32(10), p32(params, element))
    for s in e_consts[element]:
        sigma = _sigma(params, f32(s.energy * 2.96))
        mu, en = f32(s.mu_fraction), f32(s.energy)
        f_step = np.abs(mu * (p32(params, "F_STEP_OFFSET") + p32(params, "F_STEP_LINEAR") * en))
        f_tail = np.abs(p32(params, "F_TAIL_OFFSET") + p32(params, "F_TAIL_LINEAR") * mu)
        kb_f_tail = np.abs(p32(params, "KB_F_TAIL_OFFSET") + p32(params, "KB_F_TAIL_LINEAR") * mu)
        if s.ratio == 0 or s.energy <= 0:
            continue
        delta = ev - en
        faktor = f32(s.ratio) * pre
        if s.check_binding_energy(elem_info, e0):
            if s.ptype.startswith("Ka") or s.ptype.startswith("L"):
                faktor = faktor / ((f32(1.0) + f_step + f_tail) if use_tail else (f32(1.0) + f_step))
            elif s.ptype.startswith("Kb"):
                faktor = faktor / ((f32(1.0) + f_step + kb_f_tail) if use_tail else (f32(1.0) + f_step))
        else:
            faktor = f32(0)
        spec = spec + faktor * peak(gain, sigma, delta)
        if f_step > 0 and use_step:
            spec = spec + (faktor * f_step) * step(gain, sigma, delta, s.energy)
        if use_tail and (s.ptype in ("Kb1", "Kb2") or s.ptype in _TAIL_LINES):
            gamma = np.abs(p32(params, "GAMMA_OFFSET") + p32(params, "GAMMA_LINEAR") * en) * f32(s.width_multi)
            frac = kb_f_tail if s.ptype in ("Kb1", "Kb2") else f_tail
            spec = spec + (faktor * frac) * tail(gain, sigma, delta, gamma)
    return spec.astype(np.float32)


def escape_peak(spectrum, ev, escape_factor, detector_type="Si"):
    """map.py:127-150."""
    esc_e = 9.886 if detector_type == "Ge" else 1.73998
    out = np.zeros_like(spectrum)
    if ev.size < 2:
        return out
    delta_e = float(np.mean((ev[1:] - ev[:-1]).astype(np.float32), dtype=np.float32))
    if delta_e <= 0:
        return out
    bins = int(round(esc_e / delta_e))
    if bins < ev.size:
        out[: ev.size - bins] = spectrum[bins:] * f32(escape_factor)
    return out


def model_spec(params, energy_range, elements_to_fit, e_consts=None, elem_info=None, detector_type="Si",
               escape_factor=0.0):
    """map.py:260-309 (elements with use_step=True, use_tail=False)."""
    d = _defaults()
    e_consts = d.default_energy_consts if e_consts is None else e_consts
    ev = energy_axis(params, energy_range)
    agr = np.zeros_like(ev)
    for e in elements_to_fit:
        if e in e_consts and e not in SCATTER:
            agr = agr + model_elem_spec(params, e, ev, True, False, e_consts, elem_info)
    gain = p32(params, "ENERGY_SLOPE")
    agr = agr + elastic_peak(params, ev, gain)
    agr = agr + compton_peak(params, ev, gain, True, False)
    if escape_factor > 0.0:
        agr = agr + escape_peak(agr, ev, escape_factor, detector_type)
    return agr.astype(np.float32)


def component_names(elements_to_fit, e_consts=None):
    d = _defaults()
    e_consts = d.default_energy_consts if e_consts is None else e_consts
    return [e for e in elements_to_fit if e in e_consts and e not in SCATTER] + [
        "COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE"]


def basis(params, energy_range, elements_to_fit, e_consts=None, elem_info=None, detector_type="Si",
          escape_factor=0.0):
    """Unit-amplitude columns [C, E] of model_spec (the model is linear in 10**a)."""
    names = component_names(elements_to_fit, e_consts)
    ev = energy_axis(params, energy_range)
    unit = dict(params)
    cols = []
    for n in names:
        unit[n] = 0.0
        if n == "COHERENT_SCT_AMPLITUDE":
            col = elastic_peak(unit, ev, p32(params, "ENERGY_SLOPE"))
        elif n == "COMPTON_AMPLITUDE":
            col = compton_peak(unit, ev, p32(params, "ENERGY_SLOPE"), True, False)
        else:
            col = model_elem_spec(unit, n, ev, True, False, e_consts, elem_info)
        if escape_factor > 0.0:
            col = col + escape_peak(col, ev, escape_factor, detector_type)
        cols.append(col.astype(np.float32))
    return np.stack(cols, axis=1), names


# ---- evaluation of compiled term tables (mirror of csrc/mt_model.cu arithmetic) --------------
def eval_terms(terms, col_start, ev):
    ev = np.asarray(ev, dtype=np.float32)
    out = np.zeros((len(col_start) - 1, ev.size), dtype=np.float32)
    for col in range(len(col_start) - 1):
        acc = np.zeros_like(ev)
        for t in terms[col_start[col]: col_start[col + 1]]:
            d = ev - t["e"]
            if t["sign"] < 0:
                d = -d
            if t["kind"] == 0:
                q = d / t["p0"]
                v = t["amp"] * (t["p1"] * np.exp(f32(-0.5) * (q * q)))
            elif t["kind"] == 1:
                v = t["amp"] * (t["p1"] * erfc32(d / t["p0"]))
            else:
                v1 = erfc32(d / t["p0"] + t["p1"])
                with np.errstate(over="ignore", invalid="ignore"):
                    v2 = np.exp(d / t["p2"]) * v1
                v = t["amp"] * (t["p3"] * np.where(d < 0, v2, v1))
            acc = acc + v.astype(np.float32)
        out[col] = acc
    return out


# ---- SNIP background -------------------------------------------------------------------------
def snip_widths(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
    """bkg.py:83-91 in fp32."""
    idx = np.arange(n_ch, dtype=np.float32) + f32(ch0)
    energy = (f32(e_offset) + f32(e_slope) * idx) + f32(e_quad) * (idx * idx)
    a = f32(fwhm_offset) / f32(2.3548)
    sigma_sq = np.maximum(a * a + (energy * f32(2.96)) * f32(fwhm_fanoprime), f32(0))
    return (f32(snip_width) * f32(2.35)) * np.sqrt(sigma_sq) / f32(e_slope)


def snip_indices(n_ch, width):
    """bkg.py:57-62: clamp in fp32, then truncate."""
    i = np.arange(n_ch, dtype=np.float32)
    lo = np.clip(i - width, f32(0), f32(n_ch - 1)).astype(np.int64)
    hi = np.clip(i + width, f32(0), f32(n_ch - 1)).astype(np.int64)
    return lo, hi


def snip_pass_tables(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
    """Pass schedule of bkg.py:104-112 -> list of (lo, hi)."""
    w = snip_widths(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime)
    passes = [snip_indices(n_ch, w), snip_indices(n_ch, w)]
    while float(np.max(w)) >= 0.5:
        passes.append(snip_indices(n_ch, w))
        w = (w / f32(M_SQRT2)).astype(np.float32)
    return passes


def snip_bkg(spec, er, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime, boxcar_size=5):
    """bkg.py:69-114 for spec[..., C]."""
    spec = np.asarray(spec, dtype=np.float32)
    n_ch = spec.shape[-1]
    flat = spec.reshape(-1, n_ch)
    half = boxcar_size // 2
    padded = np.pad(flat, ((0, 0), (half, half)))
    bg = np.zeros_like(flat)
    w = f32(1.0) / f32(boxcar_size)
    for j in range(boxcar_size):
        bg = bg + padded[:, j: j + n_ch] * w
    with np.errstate(all="ignore"):
        bg = np.log1p(bg)
    bg = np.nan_to_num(bg, nan=0.0, posinf=0.0, neginf=0.0)
    for lo, hi in snip_pass_tables(n_ch, er[0], e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
        bg = np.minimum((bg[:, lo] + bg[:, hi]) / f32(2), bg)
    with np.errstate(all="ignore"):
        bg = np.expm1(bg)
    return np.nan_to_num(bg, nan=0.0, posinf=0.0, neginf=0.0).reshape(spec.shape).astype(np.float32)


# ---- initial amplitudes, peak windows --------------------------------------------------------
def init_elem_amps(elem, spec, e_sct, e_offset, e_slope, e_consts=None, compton_angle=None):
    """opt.py:63-97."""
    e_consts = _defaults().default_energy_consts if e_consts is None else e_consts
    try:
        if elem in SCATTER:
            if elem == "COMPTON_AMPLITUDE" and compton_angle is not None:
                th = np.deg2rad(compton_angle)
                e_energy = e_sct / (1.0 + (e_sct / 511.0) * (1.0 - np.cos(th)))
            else:
                e_energy = e_sct
            min_e, max_e = e_energy - 0.4, e_energy + 0.4
        else:
            e_energy = e_consts[elem][0].energy
            min_e, max_e = e_energy - 0.1, e_energy + 0.1
        er_min = max(0, round((min_e - e_offset) / e_slope))
        er_max = min(spec.shape[-1] - 1, round((max_e - e_offset) / e_slope))
        er_size = (er_max + 1) - er_min
        e_sum = np.sum(spec[..., er_min: er_max + 1], axis=-1) / er_size
        return np.log10(np.maximum(e_sum * 8.0 + 0.01, 1.0))
    except Exception:
        return 0.0


def get_peak_ranges(elements, coherent_sct_energy, compton_angle, energy_offset, energy_slope, energy_quadratic,
                    energy_range, e_consts=None):
    """util.py:60-108 (float64 energy axis, first channel with ev >= bound)."""
    d = _defaults()
    e_consts = d.default_energy_consts if e_consts is None else e_consts
    ch = np.arange(energy_range[0], energy_range[1] + 1, dtype=float)
    ev = energy_offset + energy_slope * ch + energy_quadratic * (ch ** 2)
    centers = {}
    for e in elements:
        if e == "COMPTON_AMPLITUDE":
            centers[e] = coherent_sct_energy / (
                1 + (coherent_sct_energy / 511) * (1 - math.cos(compton_angle * 2 * M_PI / 360.0)))
        elif e == "COHERENT_SCT_AMPLITUDE":
            centers[e] = coherent_sct_energy
        elif e in d.default_fitting_elems:
            for i, er in enumerate(e_consts[e]):
                if er.energy > 0:
                    centers[e + "_" + str(i)] = er.energy

    def chan(energy):
        hits = np.nonzero(ev >= energy)[0]
        return int(hits[0]) if hits.size else len(ev) - 1

    out = {}
    for name, c in centers.items():
        width = math.sqrt(150.0 ** 2 + (c * 12.0) ** 2) / 2000
        out[name] = (chan(c - width), chan(c + width))
    return out


# ---- the fit ---------------------------------------------------------------------------------
def fit_lstsq(spectra, B, bkg=None):
    """float64 unconstrained least squares of (spectra - bkg) on the basis columns B [C,E]."""
    y = np.asarray(spectra, dtype=np.float64).reshape(-1, B.shape[0])
    if bkg is not None:
        y = y - np.asarray(bkg, dtype=np.float64).reshape(y.shape)
    sol, *_ = np.linalg.lstsq(B.astype(np.float64), y.T, rcond=None)
    return sol.T


def fit_nnls(spectra, B, bkg=None):
    """float64 NNLS per spectrum: what the reference's 10**a parametrisation converges to
    (SURVEY.md 0.4)."""
    from scipy.optimize import nnls

    y = np.asarray(spectra, dtype=np.float64).reshape(-1, B.shape[0])
    if bkg is not None:
        y = y - np.asarray(bkg, dtype=np.float64).reshape(y.shape)
    B64 = B.astype(np.float64)
    return np.stack([nnls(B64, row, maxiter=50 * B.shape[1])[0] for row in y])


def fit_spec_adam(spectrum, B, bkg=None, init_log_amps=None, n_iter=500, lr=1e-2, loss="mse", betas=(0.9, 0.999),
                  eps=1e-8):
    """The reference's per-spectrum fit (opt.py:165-327) with tune_params=False: Adam on the
    log10 amplitudes, objective sum((B @ 10**a + bkg - y)**2) (opt.py:371-375), torch.optim.Adam
    update rule, fp32 state.  Gradients are analytic (the model is linear in 10**a)."""
    y = np.asarray(spectrum, dtype=np.float32)
    bk = np.zeros_like(y) if bkg is None else np.asarray(bkg, dtype=np.float32)
    a = np.zeros(B.shape[1], dtype=np.float32) if init_log_amps is None else np.asarray(init_log_amps, np.float32).copy()
    m = np.zeros_like(a)
    v = np.zeros_like(a)
    ln10 = f32(math.log(10.0))
    trace = []
    for t in range(1, n_iter + 1):
        amp = np.power(f32(10), a)
        r = (B @ amp + bk) - y
        if loss == "mse":
            trace.append(float(np.sum(r * r)))
            g_amp = f32(2) * (B.T @ r)
        else:
            trace.append(float(np.sum(np.abs(r))))
            g_amp = B.T @ np.sign(r)
        g = (g_amp * amp * ln10).astype(np.float32)
        m = f32(betas[0]) * m + f32(1 - betas[0]) * g
        v = f32(betas[1]) * v + f32(1 - betas[1]) * (g * g)
        bc1 = 1 - betas[0] ** t
        bc2 = 1 - betas[1] ** t
        step_size = lr / bc1
        denom = np.sqrt(v) / f32(math.sqrt(bc2)) + f32(eps)
        a = (a - f32(step_size) * (m / denom)).astype(np.float32)
    return a, trace


def fit_nnls_penalised(spectrum, B, bkg=None, c=0.0):
    """argmin_{x>=0} |x B - (y - bkg)|^2 + c*sum(x): the MSE objective with the l1_lambda penalty of
    _calculate_loss (opt.py:372-396; c = l1_lambda * loss(bkg, y) / n_elements).  Solved exactly as a
    Lawson-Hanson problem on the Cholesky factor of the Gram matrix (|L^T x - L^-1 (h - c/2)|^2 has the
    same minimiser).  B: [E, C] float64."""
    from scipy.optimize import nnls

    r = np.asarray(spectrum, dtype=np.float64) - (0.0 if bkg is None else np.asarray(bkg, dtype=np.float64))
    B = np.asarray(B, dtype=np.float64)
    live = (B * B).sum(axis=1) > 0
    x = np.zeros(B.shape[0])
    Bl = B[live]
    L = np.linalg.cholesky(Bl @ Bl.T)
    rhs = np.linalg.solve(L, Bl @ r - 0.5 * c)
    x[live] = nnls(L.T, rhs, maxiter=50 * L.shape[0])[0]
    return x


def fit_lad(spectrum, B, bkg=None, c=0.0):
    """argmin_{x>=0} sum|x B - (y - bkg)| + c*sum(x): the loss="l1" objective (opt.py:232-233, 364-396),
    as the exact linear programme  min sum(t) + c*sum(x)  s.t. -t <= x B - r <= t.  Returns (x, objective)."""
    from scipy.optimize import linprog

    r = np.asarray(spectrum, dtype=np.float64) - (0.0 if bkg is None else np.asarray(bkg, dtype=np.float64))
    B = np.asarray(B, dtype=np.float64)
    n_e, n_c = B.shape
    cost = np.r_[np.full(n_e, float(c)), np.ones(n_c)]
    eye = np.eye(n_c)
    a_ub = np.block([[B.T, -eye], [-B.T, -eye]])
    b_ub = np.r_[r, -r]
    sol = linprog(cost, A_ub=a_ub, b_ub=b_ub, bounds=[(0, None)] * (n_e + n_c), method="highs")
    if sol.status != 0:
        raise RuntimeError(f"LAD programme did not solve: {sol.message}")
    return sol.x[:n_e], float(sol.fun)


def fit_map_composed(spectra, B, snip_args=None, er=None):
    """"Reference functions, best-case composition" (BASELINE.md section 4.2): batched snip_bkg,
    then (Y - bkg) @ pinv(B) in fp32 BLAS, then scipy NNLS only for the pixels whose unconstrained
    solution has a negative component.  This is the CPU baseline bench.py times; it is far faster
    than the reference's actual per-spectrum Adam loop (fit_spec_adam above)."""
    from scipy.optimize import nnls

    y = np.asarray(spectra, dtype=np.float32).reshape(-1, B.shape[0])
    if snip_args is not None:
        y = y - snip_bkg(y, er, *snip_args)
    B64 = B.astype(np.float64)
    pinv = np.linalg.solve(B64.T @ B64, B64.T).astype(np.float32)  # [E, C]
    x = y @ pinv.T
    bad = np.nonzero((x < 0).any(axis=1))[0]
    for i in bad:
        x[i] = nnls(B64, y[i].astype(np.float64), maxiter=50 * B.shape[1])[0]
    return x


# ---- nonlinear refinement (BASELINE config 4) --------------------------------------------------
def gaussian_model_jacobian(params, energy_range, elements_to_fit, amps):
    """Analytic Jacobian of model_spec's default (pure Gaussian) path w.r.t. ENERGY_OFFSET,
    ENERGY_SLOPE, FWHM_OFFSET, FWHM_FANOPRIME and the linear amplitudes, in float64.
    Returns (model [C], d_theta [4, C], basis [C, E], names).  Formulas as in csrc/mt_refine.cu."""
    d = _defaults()
    names = component_names(elements_to_fit)
    P = {k: float(p32(params, k)) for k in d.default_param_vals}
    ch = np.arange(energy_range[0], energy_range[1] + 1, dtype=np.float64)
    ev = P["ENERGY_OFFSET"] + P["ENERGY_SLOPE"] * ch + P["ENERGY_QUADRATIC"] * ch * ch
    s0 = P["FWHM_OFFSET"] / 2.3548
    basis = np.zeros((ch.size, len(names)))
    d_ev = np.zeros(ch.size)
    d_fo = np.zeros(ch.size)
    d_fa = np.zeros(ch.size)
    e0 = P["COHERENT_SCT_ENERGY"]
    for j, name in enumerate(names):
        if name == "COHERENT_SCT_AMPLITUDE":
            lines = [(e0, 1.0, 1.0, e0 * 2.96)]
        elif name == "COMPTON_AMPLITUDE":
            ce = float(compton_energy(params))
            lines = [(ce, 1.0, P["COMPTON_FWHM_CORR"], ce * 2.96)]
        else:
            lines = [(s.energy, s.ratio, 1.0, s.energy * 2.96) for s in d.default_energy_consts[name]
                     if s.ratio != 0 and s.energy > 0 and s.check_binding_energy(d.default_elem_info, e0)]
        for energy, fac, scale, e296 in lines:
            base = np.sqrt(s0 * s0 + e296 * P["FWHM_FANOPRIME"])
            sigma = base * scale
            u = (ev - energy) / sigma
            b = fac * P["ENERGY_SLOPE"] / (sigma * SQRT_2XPI) * np.exp(-0.5 * u * u)
            basis[:, j] += b
            v = amps[j] * b
            d_ev += -v * u / sigma
            dsg = v * (u * u - 1.0) / sigma
            d_fo += dsg * scale * s0 / (2.3548 * base)
            d_fa += dsg * scale * e296 / (2.0 * base)
    model = basis @ np.asarray(amps, dtype=np.float64)
    d_theta = np.stack([d_ev, d_ev * ch + model / P["ENERGY_SLOPE"], d_fo, d_fa])
    return model, d_theta, basis, names


def refine_spectrum(spectrum, params, energy_range, elements_to_fit, free=("ENERGY_OFFSET", "FWHM_OFFSET"), bkg=None,
                    indices=None, escape=None):
    """Variable projection in float64: minimise over the free calibration parameters the NNLS
    residual of the spectrum on the Gaussian basis -- the minimiser the reference's
    fit_spec(fitting_params=free, tune_params=True) heads for.  `indices` restricts the objective to those channels
    of the window, as fit_spec(indices=...) does (opt.py:344-347); `escape` = (escape_factor, bins) adds every column's
    detector escape peak, b'(c) = b(c) + factor * b(c + bins) (escape_peak, map.py:127-150).  Returns (theta, amplitudes, loss)."""
    from scipy.optimize import least_squares, nnls

    y = np.asarray(spectrum, dtype=np.float64)[energy_range[0]: energy_range[1] + 1]
    if bkg is not None:
        y = y - np.asarray(bkg, dtype=np.float64)
    if indices is not None:
        y = y[np.asarray(indices)]
    start = np.array([float(p32(params, k)) for k in free])
    scale = np.where(start != 0, np.abs(start), 1.0)
    n_comp = len(component_names(elements_to_fit))

    def solve(theta):
        pp = dict(params, **dict(zip(free, theta)))
        # the full path of model_spec: Gaussians, step terms, scatter lines following their parameters
        _, _, B = model_jacobian(pp, energy_range, elements_to_fit, np.ones(n_comp), free=())
        if escape is not None and escape[1] > 0:
            shifted = np.zeros_like(B)
            shifted[: B.shape[0] - escape[1]] = B[escape[1]:] * escape[0]
            B = B + shifted
        if indices is not None:
            B = B[np.asarray(indices)]
        x, _ = nnls(B, y, maxiter=50 * n_comp)
        return B, x

    def resid(t):
        B, x = solve(t * scale)
        return B @ x - y

    sol = least_squares(resid, start / scale, method="trf", diff_step=1e-5, xtol=1e-12, ftol=1e-14, gtol=1e-12,
                        max_nfev=400)
    theta = sol.x * scale
    B, x = solve(theta)
    return theta, x, float(np.sum((B @ x - y) ** 2))


# ---- general analytic Jacobian of model_spec (Gaussians + step terms, all effective calibration parameters) ------
THETA_NAMES = ("ENERGY_OFFSET", "ENERGY_SLOPE", "FWHM_OFFSET", "FWHM_FANOPRIME", "ENERGY_QUADRATIC",
               "COHERENT_SCT_ENERGY", "COMPTON_ANGLE", "COMPTON_FWHM_CORR")


def model_lines(params, elements_to_fit, e_consts=None, elem_info=None):
    """Every term of model_spec's path (map.py:260-309: use_step=True, use_tail=False) as a line record
    (component, kind, energy, factor, step fraction): kind 0 = element line, 1 = elastic peak (energy =
    COHERENT_SCT_ENERGY), 2 = Compton peak.  factor holds ratio / (1 + f_step) as model_elem_spec forms it
    (map.py:183-207); gates are evaluated at the global COHERENT_SCT_ENERGY."""
    d = _defaults()
    e_consts = d.default_energy_consts if e_consts is None else e_consts
    elem_info = d.default_elem_info if elem_info is None else elem_info
    P = {k: float(p32(params, k)) for k in d.default_param_vals}
    e0 = P["COHERENT_SCT_ENERGY"]
    out = []
    for comp, name in enumerate(component_names(elements_to_fit, e_consts)):
        if name == "COHERENT_SCT_AMPLITUDE":
            out.append((comp, 1, e0, 1.0, 0.0))
        elif name == "COMPTON_AMPLITUDE":
            fs = P["COMPTON_F_STEP"] if P["COMPTON_F_STEP"] > 0 else 0.0
            out.append((comp, 2, float(compton_energy(params)), 1.0 / (1.0 + fs), fs))
        else:
            for s in e_consts[name]:
                if s.ratio == 0 or s.energy <= 0 or not s.check_binding_energy(elem_info, e0):
                    continue
                f_step = abs(float(s.mu_fraction) * (P["F_STEP_OFFSET"] + P["F_STEP_LINEAR"] * float(s.energy)))
                fac = float(s.ratio)
                if s.ptype.startswith("Ka") or s.ptype.startswith("L") or s.ptype.startswith("Kb"):
                    fac = fac / (1.0 + f_step)
                out.append((comp, 0, float(s.energy), fac, f_step))
    return out


def model_jacobian(params, energy_range, elements_to_fit, amps, free=THETA_NAMES, e_consts=None, elem_info=None):
    """float64 model of model_spec's path and its analytic derivatives w.r.t. the calibration parameters in `free`
    (any of THETA_NAMES) at fixed linear amplitudes `amps`.  Returns (model [C], d_theta [len(free), C], basis [C, E]).
    The formulas are the ones csrc/mt_refine.cu evaluates; checked against torch.autograd of the reference
    (tests/golden/make_golden.py --only jacobian)."""
    from scipy.special import erfc

    d = _defaults()
    P = {k: float(p32(params, k)) for k in d.default_param_vals}
    names = component_names(elements_to_fit, e_consts)
    lines = model_lines(params, elements_to_fit, e_consts, elem_info)
    ch = np.arange(energy_range[0], energy_range[1] + 1, dtype=np.float64)
    off, slope, quad = P["ENERGY_OFFSET"], P["ENERGY_SLOPE"], P["ENERGY_QUADRATIC"]
    ev = off + slope * ch + quad * ch * ch
    s0 = P["FWHM_OFFSET"] / 2.3548
    fa = P["FWHM_FANOPRIME"]
    e0, ang, corr = P["COHERENT_SCT_ENERGY"], P["COMPTON_ANGLE"], P["COMPTON_FWHM_CORR"]
    rad = ang * 2.0 * M_PI / 360.0
    a = (e0 / 511.0) * (1.0 - np.cos(rad))
    basis = np.zeros((ch.size, len(names)))
    s_ev = np.zeros(ch.size)  # sum of d(term)/d(ev)
    acc = {k: np.zeros(ch.size) for k in ("FWHM_OFFSET", "FWHM_FANOPRIME", "COHERENT_SCT_ENERGY", "COMPTON_ANGLE",
                                          "COMPTON_FWHM_CORR")}
    amps = np.asarray(amps, dtype=np.float64)
    for comp, kind, energy, fac, fstep in lines:
        # line energy and its derivatives w.r.t. the incident energy / scattering angle
        d_e = {"COHERENT_SCT_ENERGY": 0.0, "COMPTON_ANGLE": 0.0}
        if kind == 1:
            energy, d_e["COHERENT_SCT_ENERGY"] = e0, 1.0
        elif kind == 2:
            energy = e0 / (1.0 + a)
            d_e["COHERENT_SCT_ENERGY"] = 1.0 / (1.0 + a) ** 2
            d_e["COMPTON_ANGLE"] = -e0 / (1.0 + a) ** 2 * (e0 / 511.0) * np.sin(rad) * (2.0 * M_PI / 360.0)
        sb = np.sqrt(s0 * s0 + energy * 2.96 * fa)  # width of the line (step terms use it as it is)
        scale = corr if kind == 2 else 1.0
        sg = sb * scale                              # width of the Gaussian
        # d(sb)/d(theta) / sb
        c_sb = {"FWHM_OFFSET": s0 / (2.3548 * sb * sb), "FWHM_FANOPRIME": energy * 2.96 / (2.0 * sb * sb),
                "COHERENT_SCT_ENERGY": 2.96 * fa / (2.0 * sb * sb) * d_e["COHERENT_SCT_ENERGY"],
                "COMPTON_ANGLE": 2.96 * fa / (2.0 * sb * sb) * d_e["COMPTON_ANGLE"], "COMPTON_FWHM_CORR": 0.0}
        c_sg = dict(c_sb, COMPTON_FWHM_CORR=(1.0 / corr if kind == 2 else 0.0))
        dl = ev - energy
        g = np.exp(-0.5 * (dl / sg) ** 2)
        b = fac * slope / (sg * SQRT_2XPI) * g
        v = amps[comp] * b
        t1 = v * dl / (sg * sg)                 # -d v / d ev
        t2 = v * ((dl / sg) ** 2 - 1.0)         # sg * d v / d sg
        s_ev += -t1
        for k in acc:
            acc[k] += t2 * c_sg[k] + t1 * d_e.get(k, 0.0)
        if fstep > 0:
            z = dl / (M_SQRT2 * sb)
            kk = fac * fstep * slope / (2.0 * energy)
            bs = kk * erfc(z)
            b = b + bs
            s = amps[comp] * bs
            q = amps[comp] * kk * (2.0 / np.sqrt(M_PI)) * np.exp(-z * z) / (M_SQRT2 * sb)  # -d s / d ev
            s_ev += -q
            for k in acc:
                acc[k] += q * (dl / sb) * sb * c_sb[k] + (q - s / energy) * d_e.get(k, 0.0)
        basis[:, comp] += b
    model = basis @ amps
    full = {"ENERGY_OFFSET": s_ev, "ENERGY_SLOPE": s_ev * ch + model / slope, "ENERGY_QUADRATIC": s_ev * ch * ch}
    full.update(acc)
    return model, (np.stack([full[k] for k in free]) if len(free) else np.zeros((0, ch.size))), basis
