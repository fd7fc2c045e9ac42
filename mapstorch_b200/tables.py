"""Host-side "compile step": parameter/element dictionaries -> flat tables for the kernels.

Everything here depends only on the ~30 global calibration parameters, never on pixel data,
so it runs once per fit on the host and reproduces the reference's fp32 scalar arithmetic
(and its float->int conversions) exactly; the CUDA kernels then only consume tables.

* :func:`energy_axis`     -- ``ev[c]`` of map.py:271-281 (fp32, same op order, no FMA)
* :func:`compile_terms`   -- line tables of model_elem_spec / elastic_peak / compton_peak
                             (map.py:72-124, 152-257) as ``MtTerm`` records
* :func:`snip_tables`     -- per-pass gather indices of snip_bkg/snip_op (bkg.py:53-66, 83-112)
* :func:`escape_bins`     -- the integer bin shift of escape_peak (map.py:127-148)
* :func:`amp_window`      -- the channel window of init_elem_amps (opt.py:74-90)
"""
import numpy as np
import torch

from mapstorch_b200.constant import M_PI, M_SQRT2, SQRT_2XPI
from mapstorch_b200._native import MT_TERM_GAUSS, MT_TERM_STEP, MT_TERM_TAIL

TERM_DTYPE = np.dtype([("kind", "<i4"), ("sign", "<f4"), ("e", "<f4"), ("p0", "<f4"), ("p1", "<f4"),
                       ("p2", "<f4"), ("p3", "<f4"), ("amp", "<f4")])
assert TERM_DTYPE.itemsize == 32

SCATTER_NAMES = ("COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE")
_TAIL_LINES = frozenset("Ka1 Ka2 La1 La2 Lb1 Lb2 Lb3 Lb4 Lg1 Lg2 Lg3 Lg4 Ll Ln".split())
_KB_TAIL_LINES = frozenset(("Kb1", "Kb2"))

f32 = np.float32
_SQRT2 = f32(M_SQRT2)
_SQRT2PI = f32(SQRT_2XPI)


def scalar32(v):
    """A parameter (python float, numpy scalar or 0-dim tensor) as the fp32 value torch would use."""
    if isinstance(v, torch.Tensor):
        return f32(v.detach().to(torch.float32).item())
    return f32(v)


def _t32(v):
    return torch.tensor(float(v), dtype=torch.float32)


def _exp32(v):
    return f32(torch.exp(_t32(v)).item())


def _cos32(v):
    return f32(torch.cos(_t32(v)).item())


def pow10(v):
    """``10 ** tensor`` as torch evaluates it on the CPU for an fp32 0-dim tensor (map.py:78,103,163)."""
    return f32((10 ** _t32(scalar32(v))).item())


def energy_axis(params, energy_range):
    """fp32 energy of channels energy_range[0]..energy_range[1] (inclusive), CPU tensor."""
    ch = torch.arange(energy_range[0], energy_range[1] + 1, dtype=torch.float32)
    off, slope, quad = (_t32(scalar32(params[k])) for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC"))
    return off + slope * ch + quad * (ch ** 2)


def _sigma(fwhm_offset, e_times_296_f32, fano):
    a = fwhm_offset / f32(2.3548)
    return np.sqrt(a * a + e_times_296_f32 * fano)


def _gauss(e, sigma, gain, amp):
    return (MT_TERM_GAUSS, 1.0, e, sigma, gain / (sigma * _SQRT2PI), 0.0, 0.0, amp)


def _step(e, sigma, gain, peak_e, amp):
    return (MT_TERM_STEP, 1.0, e, _SQRT2 * sigma, gain / f32(2.0) / peak_e, 0.0, 0.0, amp)


def _tail(e, sigma, gain, gamma, amp, sign=1.0):
    inv = f32(1.0) / (gamma * _SQRT2)
    norm_exp = _exp32((f32(1.0) / (gamma * gamma)) * f32(-0.5))
    scale = gain / f32(2.0) / gamma / sigma / norm_exp
    return (MT_TERM_TAIL, sign, e, _SQRT2 * sigma, inv, gamma * sigma, scale, amp)


def element_terms(params, slots, elem_info, amp, use_step, use_tail):
    """Terms of one element column; `amp` is the linear amplitude (10**logA) as fp32."""
    P = {k: scalar32(params[k]) for k in (
        "FWHM_OFFSET", "FWHM_FANOPRIME", "F_STEP_OFFSET", "F_STEP_LINEAR", "F_TAIL_OFFSET", "F_TAIL_LINEAR",
        "KB_F_TAIL_OFFSET", "KB_F_TAIL_LINEAR", "GAMMA_OFFSET", "GAMMA_LINEAR", "ENERGY_SLOPE",
        "COHERENT_SCT_ENERGY")}
    gain, e0 = P["ENERGY_SLOPE"], P["COHERENT_SCT_ENERGY"]
    out = []
    with np.errstate(all="ignore"):
        for s in slots:
            if s.ratio == 0 or s.energy <= 0:
                continue
            if not s.check_binding_energy(elem_info, e0):
                continue  # reference multiplies the line by faktor = 0 (map.py:206-207)
            e = f32(s.energy)
            mu = f32(s.mu_fraction)
            sigma = _sigma(P["FWHM_OFFSET"], f32(s.energy * 2.96), P["FWHM_FANOPRIME"])
            f_step = np.abs(mu * (P["F_STEP_OFFSET"] + P["F_STEP_LINEAR"] * e))
            f_tail = np.abs(P["F_TAIL_OFFSET"] + P["F_TAIL_LINEAR"] * mu)
            kb_f_tail = np.abs(P["KB_F_TAIL_OFFSET"] + P["KB_F_TAIL_LINEAR"] * mu)
            faktor = f32(s.ratio) * amp
            family = s.ptype[:2]
            if family in ("Ka", "Kb") or s.ptype.startswith("L"):
                norm = f32(1.0) + f_step
                if use_tail:
                    norm = norm + (kb_f_tail if family == "Kb" else f_tail)
                faktor = faktor / norm
            out.append(_gauss(e, sigma, gain, faktor))
            if f_step > 0 and use_step:
                out.append(_step(e, sigma, gain, e, faktor * f_step))
            if use_tail and (s.ptype in _KB_TAIL_LINES or s.ptype in _TAIL_LINES):
                gamma = np.abs(P["GAMMA_OFFSET"] + P["GAMMA_LINEAR"] * e) * f32(s.width_multi)
                frac = kb_f_tail if s.ptype in _KB_TAIL_LINES else f_tail
                out.append(_tail(e, sigma, gain, gamma, faktor * frac))
    return out


def compton_energy(params):
    e0 = scalar32(params["COHERENT_SCT_ENERGY"])
    angle = scalar32(params["COMPTON_ANGLE"])
    c = _cos32(angle * f32(2) * f32(M_PI) / f32(360.0))
    return e0 / (f32(1) + (e0 / f32(511)) * (f32(1) - c))


def elastic_terms(params, gain, amp):
    e0 = scalar32(params["COHERENT_SCT_ENERGY"])
    sigma = _sigma(scalar32(params["FWHM_OFFSET"]), e0 * f32(2.96), scalar32(params["FWHM_FANOPRIME"]))
    return [_gauss(e0, sigma, gain, amp)]


def compton_terms(params, gain, amp, use_step=True, use_tail=False):
    P = {k: scalar32(params[k]) for k in ("COMPTON_F_STEP", "COMPTON_F_TAIL", "COMPTON_HI_F_TAIL",
                                          "COMPTON_FWHM_CORR", "COMPTON_GAMMA", "COMPTON_HI_GAMMA")}
    ce = compton_energy(params)
    sigma = _sigma(scalar32(params["FWHM_OFFSET"]), ce * f32(2.96), scalar32(params["FWHM_FANOPRIME"]))
    with np.errstate(all="ignore"):
        norm = f32(1.0)
        if use_step and P["COMPTON_F_STEP"] > 0:
            norm = norm + P["COMPTON_F_STEP"]
        if use_tail:
            if P["COMPTON_F_TAIL"] > 0:
                norm = norm + P["COMPTON_F_TAIL"]
            if P["COMPTON_HI_F_TAIL"] > 0:
                norm = norm + P["COMPTON_HI_F_TAIL"]
        faktor = amp / norm
        out = [_gauss(ce, sigma * P["COMPTON_FWHM_CORR"], gain, faktor)]
        if use_step and P["COMPTON_F_STEP"] > 0:
            out.append(_step(ce, sigma, gain, ce, faktor * P["COMPTON_F_STEP"]))
        if use_tail:
            if P["COMPTON_F_TAIL"] > 0:
                out.append(_tail(ce, sigma, gain, P["COMPTON_GAMMA"], faktor * P["COMPTON_F_TAIL"]))
            if P["COMPTON_HI_F_TAIL"] > 0:
                out.append(_tail(ce, sigma, gain, P["COMPTON_HI_GAMMA"], faktor * P["COMPTON_HI_F_TAIL"], sign=-1.0))
    return out


def pack_columns(columns):
    """list of term lists -> (structured term array, int32 col_start[n_cols+1])."""
    flat = [t for col in columns for t in col]
    terms = np.zeros(max(len(flat), 1), dtype=TERM_DTYPE)
    for i, t in enumerate(flat):
        terms[i] = t
    starts = np.zeros(len(columns) + 1, dtype=np.int32)
    np.cumsum([len(c) for c in columns], out=starts[1:])
    return terms[: max(len(flat), 1)], starts


def fitted_components(elements_to_fit, e_consts):
    """Column order of model_spec: element columns in the given order (map.py:282-300), then the
    elastic and the Compton peak (map.py:301-304)."""
    cols = [e for e in elements_to_fit if e in e_consts and e not in SCATTER_NAMES]
    return cols + ["COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE"]


def compile_terms(params, components, e_consts, elem_info, use_step=True, use_tail=False, log_amps=None):
    """Term lists for `components` (element names and/or the two scatter names).

    log_amps: optional dict name -> log10 amplitude; default 0 (unit amplitude -> basis columns)."""
    gain = scalar32(params["ENERGY_SLOPE"])
    columns = []
    for name in components:
        amp = pow10(log_amps[name]) if log_amps is not None else f32(1.0)
        if name == "COHERENT_SCT_AMPLITUDE":
            columns.append(elastic_terms(params, gain, amp))
        elif name == "COMPTON_AMPLITUDE":
            columns.append(compton_terms(params, gain, amp, use_step, use_tail))
        else:
            columns.append(element_terms(params, e_consts[name], elem_info, amp, use_step, use_tail))
    return columns


# ---- vectorised line-table compiler --------------------------------------------------------------------------------
# fit_spec(tune_params=True) rebuilds the term table at every trial point of its Levenberg-Marquardt loop; the scalar
# code above costs ~0.4 ms for ten elements (NumPy scalar arithmetic per line).  compile_packed does the same fp32
# operations on arrays over all lines at once (IEEE single-precision adds, multiplies, divisions and square roots round
# identically for scalars and arrays) and writes the packed table directly; everything that does not depend on the
# parameters (line energies, ratios, absorption edges, column membership) is gathered once per component list.
_LINE_SETS = {}


class _LineSet:
    """Static per-line arrays of the element columns of one component list."""

    def __init__(self, components, e_consts, elem_info):
        e, e296, mu, ratio, kl, col, edge = [], [], [], [], [], [], []
        self.columns = []  # per component: ("lines", first, last) or the scatter name
        for name in components:
            if name in SCATTER_NAMES:
                self.columns.append(name)
                continue
            first = len(e)
            for s in e_consts[name]:
                if s.ratio == 0 or s.energy <= 0:
                    continue
                edges = s.binding_edge(elem_info)
                if edges is None:
                    continue  # never excited (check_binding_energy is False for every incident energy)
                e.append(f32(s.energy))
                e296.append(f32(s.energy * 2.96))
                mu.append(f32(s.mu_fraction))
                ratio.append(f32(s.ratio))
                kl.append(s.ptype[:2] in ("Ka", "Kb") or s.ptype.startswith("L"))
                col.append(len(self.columns))
                # the gate compares each edge with the fp32 incident energy: Python floats enter that comparison as fp32
                pair = [f32(v) for v in edges] + [f32(-np.inf)] * (2 - len(edges))
                edge.append(pair[:2])
                if len(edges) > 2:
                    raise ValueError("more than two absorption edges per line")
            self.columns.append(("lines", first, len(e)))
        self.e = np.asarray(e, dtype=np.float32)
        self.e296 = np.asarray(e296, dtype=np.float32)
        self.mu = np.asarray(mu, dtype=np.float32)
        self.ratio = np.asarray(ratio, dtype=np.float32)
        self.kl = np.asarray(kl, dtype=bool)
        self.col = np.asarray(col, dtype=np.int64)
        self.edge = np.asarray(edge, dtype=np.float32).reshape(-1, 2)
        self.consts = (e_consts, elem_info)  # keeps the keyed objects alive (the cache key holds their ids)


def _line_set(components, e_consts, elem_info):
    key = (tuple(components), id(e_consts), id(elem_info))
    ls = _LINE_SETS.get(key)
    if ls is None:
        if len(_LINE_SETS) > 32:
            _LINE_SETS.clear()
        ls = _LINE_SETS[key] = _LineSet(components, e_consts, elem_info)
    return ls


def compile_packed(params, components, e_consts, elem_info, use_step=True, use_tail=False, log_amps=None):
    """pack_columns(compile_terms(...)) in one vectorised step: (structured term array, int32 col_start[n_cols+1]).
    Bit-identical to the scalar route (tests/test_tables.py); tail terms take the scalar route."""
    if use_tail:
        return pack_columns(compile_terms(params, components, e_consts, elem_info, use_step, use_tail, log_amps))
    ls = _line_set(components, e_consts, elem_info)
    gain = scalar32(params["ENERGY_SLOPE"])
    e0 = scalar32(params["COHERENT_SCT_ENERGY"])
    n_cols = len(ls.columns)
    amps = np.ones(n_cols, dtype=np.float32)
    if log_amps is not None:
        for k, name in enumerate(components):
            amps[k] = pow10(log_amps[name])
    n = ls.e.shape[0]
    with np.errstate(all="ignore"):
        a = scalar32(params["FWHM_OFFSET"]) / f32(2.3548)
        sigma = np.sqrt(a * a + ls.e296 * scalar32(params["FWHM_FANOPRIME"]))
        f_step = np.abs(ls.mu * (scalar32(params["F_STEP_OFFSET"]) + scalar32(params["F_STEP_LINEAR"]) * ls.e))
        faktor = ls.ratio * amps[ls.col]
        faktor = np.where(ls.kl, faktor / (f32(1.0) + f_step), faktor)
        on = (ls.edge[:, 0] < e0) & (ls.edge[:, 1] < e0)
        rows = np.zeros((n, 2), dtype=TERM_DTYPE)  # [line, gauss | step]
        rows["sign"] = 1.0
        rows["e"] = ls.e[:, None]
        g, st = rows[:, 0], rows[:, 1]
        g["kind"] = MT_TERM_GAUSS
        g["p0"] = sigma
        g["p1"] = gain / (sigma * _SQRT2PI)
        g["amp"] = faktor
        st["kind"] = MT_TERM_STEP
        st["p0"] = _SQRT2 * sigma
        st["p1"] = gain / f32(2.0) / ls.e
        st["amp"] = faktor * f_step
        keep = np.stack([on, on & (f_step > 0) & bool(use_step)], axis=1)
    per_col = np.bincount(ls.col, weights=keep.sum(axis=1), minlength=n_cols).astype(np.int64) if n else \
        np.zeros(n_cols, dtype=np.int64)
    line_terms = rows[keep]  # row-major: line order, Gaussian before its step
    scatter = {}
    for k, c in enumerate(ls.columns):
        if c == "COHERENT_SCT_AMPLITUDE":
            scatter[k] = elastic_terms(params, gain, amps[k])
        elif c == "COMPTON_AMPLITUDE":
            scatter[k] = compton_terms(params, gain, amps[k], use_step, use_tail)
        if k in scatter:
            per_col[k] = len(scatter[k])
    starts = np.zeros(n_cols + 1, dtype=np.int32)
    np.cumsum(per_col, out=starts[1:])
    total = int(starts[-1])
    terms = np.zeros(max(total, 1), dtype=TERM_DTYPE)
    first_scatter = min(scatter) if scatter else n_cols
    if all(k >= first_scatter for k in scatter) and len(scatter) == n_cols - first_scatter:
        # model_spec's column order: element columns, then the scatter peaks -- the line terms are one contiguous run
        terms[: line_terms.shape[0]] = line_terms
        for k, col in scatter.items():
            for i, t in enumerate(col):
                terms[int(starts[k]) + i] = t
        return terms, starts
    src = 0
    for k, c in enumerate(ls.columns):  # element columns are contiguous runs of line_terms in column order
        lo, hi = int(starts[k]), int(starts[k + 1])
        if k in scatter:
            for i, t in enumerate(scatter[k]):
                terms[lo + i] = t
        else:
            terms[lo:hi] = line_terms[src: src + hi - lo]
            src += hi - lo
    return terms, starts


def snip_tables(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
    """Gather indices of every SNIP clipping pass: uint32 [n_pass, n_ch], lo | hi << 16.

    The fp32 operation sequence of bkg.py:83-91 (widths), 104-112 (pass schedule) and 57-62 (clamp, then truncate to
    int) in NumPy: IEEE single-precision adds, multiplies, divisions and square roots round identically in both
    libraries, parameters given as 0-dim tensors enter as float32 scalars and Python floats as weakly typed scalars
    exactly as under torch's promotion rules, and the indices of all passes are formed in one vectorised step.
    Bit-identical to snip_tables_torch (the reference's ops verbatim) on every case of tests/test_tables.py; a
    fifth of its time -- it runs once per trial point of fit_spec(tune_params=True)."""
    if n_ch > 65535:
        raise ValueError("snip tables hold 16-bit channel indices")

    width = snip_widths(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime)
    idx = np.arange(n_ch, dtype=np.float32)
    widths = [width, width]
    while float(width.max()) >= 0.5:
        widths.append(width)
        width = width / f32(M_SQRT2)
        if len(widths) > 4096:
            raise ValueError("SNIP width does not shrink (check ENERGY_SLOPE / SNIP_WIDTH)")
    w = np.stack(widths)
    lo = np.clip(idx[None, :] - w, f32(0), f32(n_ch - 1)).astype(np.int64)
    hi = np.clip(idx[None, :] + w, f32(0), f32(n_ch - 1)).astype(np.int64)
    return np.ascontiguousarray((lo | (hi << 16)).astype(np.uint32))


def snip_widths(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
    """First-pass clipping half-width per channel (bkg.py:83-91), float32 [n_ch]; see snip_tables."""

    def operand(v):
        return f32(v.detach().cpu().to(torch.float32).item()) if isinstance(v, torch.Tensor) else v

    e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime = map(
        operand, (e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime))
    idx = np.arange(n_ch, dtype=np.float32)
    gidx = idx + f32(ch0)
    energy = (e_offset + (e_slope * gidx) + (e_quad * (gidx * gidx))).astype(np.float32)
    if isinstance(fwhm_offset, np.floating):
        t = fwhm_offset / f32(2.3548)
        first = t * t
    else:
        first = (fwhm_offset / 2.3548) ** 2  # Python floats: double arithmetic, rounded when it meets the tensor
    sigma_sq = np.maximum((first + energy * 2.96 * fwhm_fanoprime).astype(np.float32), f32(0.0))
    return ((snip_width * 2.35 * np.sqrt(sigma_sq)).astype(np.float32) / f32(e_slope)).astype(np.float32)


def snip_pass_count(width_max):
    """Number of clipping passes (bkg.py:104-112) from the largest first-pass width alone: every pass divides all widths
    by the same constant and rounding is monotone, so the largest width stays the largest."""
    w, n = f32(width_max), 2
    while float(w) >= 0.5:
        n += 1
        w = w / f32(M_SQRT2)
        if n > 4096:
            raise ValueError("SNIP width does not shrink (check ENERGY_SLOPE / SNIP_WIDTH)")
    return n


def snip_tables_torch(n_ch, ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime):
    """The table builder written with the reference's own torch CPU ops (bkg.py:83-91, 104-112, 57-62): what
    snip_tables is checked against (tests/test_tables.py); uint32 [n_pass, n_ch], lo | hi << 16.

    Same torch-CPU op sequence as bkg.py:83-91 (widths), 104-112 (pass schedule) and 57-62
    (clamp, then truncate to int), so the indices are bit-identical to the reference's."""
    if n_ch > 65535:
        raise ValueError("snip tables hold 16-bit channel indices")

    def as_operand(v):
        return v.detach().cpu().to(torch.float32) if isinstance(v, torch.Tensor) else v

    e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime = map(
        as_operand, (e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime))
    idx = torch.arange(n_ch, dtype=torch.float32)
    gidx = idx + ch0
    energy = e_offset + (e_slope * gidx) + (e_quad * (gidx ** 2))
    sigma_sq = torch.clamp((fwhm_offset / 2.3548) ** 2 + energy * 2.96 * fwhm_fanoprime, min=0.0)
    width = snip_width * 2.35 * torch.sqrt(sigma_sq) / e_slope

    def indices(w):
        lo = torch.clamp(idx - w, 0, n_ch - 1).to(torch.int64)
        hi = torch.clamp(idx + w, 0, n_ch - 1).to(torch.int64)
        return (lo | (hi << 16)).numpy().astype(np.uint32)

    passes = [indices(width), indices(width)]
    while torch.max(width).item() >= 0.5:
        passes.append(indices(width))
        width = width / M_SQRT2
        if len(passes) > 4096:
            raise ValueError("SNIP width does not shrink (check ENERGY_SLOPE / SNIP_WIDTH)")
    return np.ascontiguousarray(np.stack(passes))


def escape_bins(ev, detector_type="Si"):
    """Bin shift of the escape peak, or 0 if it falls outside the spectrum (map.py:127-148)."""
    escape_e = 9.886 if detector_type == "Ge" else 1.73998
    if ev.numel() < 2:
        return 0
    delta_e = torch.mean(ev[1:] - ev[:-1]).item()
    if delta_e <= 0:
        return 0
    bins = int(round(escape_e / delta_e))
    return bins if bins < ev.numel() else 0


def amp_window(elem, n_ch, e_sct, e_offset, e_slope, e_consts, compton_angle=None):
    """Inclusive channel window [lo, hi] used for the initial amplitude guess (opt.py:74-90)."""
    if elem in SCATTER_NAMES:
        if elem == "COMPTON_AMPLITUDE" and compton_angle is not None:
            theta = np.deg2rad(compton_angle)
            centre = e_sct / (1.0 + (e_sct / 511.0) * (1.0 - np.cos(theta)))
        else:
            centre = e_sct
        half = 0.4
    else:
        centre = e_consts[elem][0].energy
        half = 0.1
    lo = max(0, round((centre - half - e_offset) / e_slope))
    hi = min(n_ch - 1, round((centre + half - e_offset) / e_slope))
    return lo, hi
