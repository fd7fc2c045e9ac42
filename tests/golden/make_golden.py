#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Only runs in the build container, where the unmodified reference is mounted read-only at
/root/reference (it cannot travel to the GPU box; the fixtures do):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py [--skip-fit]

Outputs (np.savez_compressed):
  model_golden.npz  model_elem_spec / compton_peak / elastic_peak / model_spec outputs
  snip_golden.npz   snip_bkg inputs+outputs and the per-pass gather indices
  fit_golden.npz    per-spectrum fit_spec(tune_params=False) runs (converged Adam) + loss traces
  jacobian_golden.npz  d model_spec / d calibration parameter by torch.autograd of the reference (--only jacobian)
  fit_c3_golden.npz the same at BASELINE config 3's shape: 4096 channels, 30 elements, SNIP (--only fit_c3)
  host_golden.npz   init_elem_amps, get_peak_ranges, escape bins
  objectives_golden.npz  fit_spec(tune_params=False) with loss="l1", l1_lambda > 0 and `indices` (--only objectives)
  override_golden.npz  override files written by the reference, what its parser and read_constants make of
                       them, and element spectra evaluated with the edited branching ratios
The reference ships no tests or fixtures of its own, so these are the pins.
"""
import argparse
import os
import sys
import types
import warnings
from pathlib import Path

warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402

if os.environ.get("PYTHONHASHSEED") != "0":
    print("re-exec with PYTHONHASHSEED=0 (fit_spec iterates a set() of element names, opt.py:194-200)")
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

sys.path.insert(0, "/root/reference")
# mapstorch.util imports two UI packages that are not installed; stub them (only the pure helpers are used)
for _name in ("anywidget", "traitlets"):
    _m = types.ModuleType(_name)
    sys.modules[_name] = _m
sys.modules["anywidget"].AnyWidget = type("AnyWidget", (), {})


class _Trait:
    def __init__(self, *a, **k):
        pass

    def tag(self, **k):
        return self


for _t in ("Dict", "List", "Int", "Unicode", "Bool", "Float"):
    setattr(sys.modules["traitlets"], _t, _Trait)

sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # mapstorch.io imports it; only the text-file helpers are used

from mapstorch import bkg as rbkg, constant as rconst, io as rio, map as rmap, opt as ropt, util as rutil  # noqa: E402
from mapstorch.default import default_param_vals, default_fitting_elems, default_energy_consts  # noqa: E402

HERE = Path(__file__).resolve().parent
ELEMS10 = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
ELEMS20 = "Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As".split()
ELEMS30 = ELEMS20[:19] + "Ba_L Ce_L Gd_L W_L Pt_L Au_L Au_M Pt_M Hg_M W_M Si_Si".split()
MODEL_ELEMS = ["Fe", "Ca", "Zn", "Si", "Pb_L", "Au_L", "Ba_L", "Au_M", "Pt_M", "Si_Si", "Si_Cl", "Hf_L"]


def params_a():
    p = dict(default_param_vals)
    p["COHERENT_SCT_ENERGY"] = 12.0
    return p


def params_b():
    p = dict(default_param_vals)
    p.update({"COHERENT_SCT_ENERGY": 17.5, "F_STEP_OFFSET": 0.002, "F_STEP_LINEAR": 0.0004,
              "COMPTON_F_STEP": 0.01, "GAMMA_LINEAR": 0.05, "F_TAIL_LINEAR": 1e-4, "KB_F_TAIL_LINEAR": 2e-4,
              "ENERGY_OFFSET": 0.0113, "FWHM_OFFSET": 0.0912})
    return p


def params_4096():
    p = params_a()
    p.update({"ENERGY_SLOPE": 0.00599, "ENERGY_QUADRATIC": 0.0})
    return p


def tens(p, extra=None):
    t = {k: torch.tensor(float(v)) for k, v in p.items()}
    for k, v in (extra or {}).items():
        t[k] = torch.tensor(float(v))
    return t


def ev_of(t, er):
    ch = torch.arange(er[0], er[1] + 1, dtype=torch.float32)
    return t["ENERGY_OFFSET"] + t["ENERGY_SLOPE"] * ch + t["ENERGY_QUADRATIC"] * (ch ** 2)


def gen_model(out):
    for tag, p, er in (("a", params_a(), (0, 2047)), ("b", params_b(), (50, 1450)), ("c", params_4096(), (0, 4095))):
        out[f"{tag}_er"] = np.array(er)
        out[f"{tag}_param_names"] = np.array(list(p.keys()))
        out[f"{tag}_param_vals"] = np.array([p[k] for k in p], dtype=np.float64)
        for logamp in (0.0, 2.3):
            t = tens(p, {e: logamp for e in default_fitting_elems})
            ev = ev_of(t, er)
            if logamp == 0.0:
                out[f"{tag}_ev"] = ev.numpy()
            for us in (False, True):
                for ut in (False, True):
                    if tag == "c" and (logamp != 0.0):
                        continue
                    key = f"{tag}_amp{logamp}_s{int(us)}_t{int(ut)}"
                    cols = [rmap.model_elem_spec(t, e, ev, us, ut).numpy() for e in MODEL_ELEMS]
                    out[key + "_elems"] = np.stack(cols)
                    out[key + "_compton"] = rmap.compton_peak(t, ev, t["ENERGY_SLOPE"], us, ut).numpy()
            out[f"{tag}_amp{logamp}_elastic"] = rmap.elastic_peak(t, ev, t["ENERGY_SLOPE"]).numpy()
        # every default component at unit amplitude, default flags of model_spec: compact statistics
        t = tens(p, {e: 0.0 for e in default_fitting_elems})
        ev = ev_of(t, er)
        names = [e for e in default_fitting_elems if e in default_energy_consts]
        stats = []
        for e in names:
            c = rmap.model_elem_spec(t, e, ev, True, False).numpy().astype(np.float64)
            stats.append([c.sum(), c.max(), float(c.argmax()), (c * c).sum()])
        out[f"{tag}_all_names"] = np.array(names)
        out[f"{tag}_all_stats"] = np.array(stats)
        # model_spec with amplitudes, with and without escape peaks
        rng = np.random.default_rng(7)
        for set_name, elems in (("e10", ELEMS10), ("e30", ELEMS30)):
            amps = {e: float(rng.uniform(1.5, 3.0)) for e in elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]}
            t = tens(p, amps)
            out[f"{tag}_{set_name}_logamps"] = np.array([amps[e] for e in elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]])
            out[f"{tag}_{set_name}_spec"] = rmap.model_spec(t, er, elems).numpy()
            out[f"{tag}_{set_name}_spec_si"] = rmap.model_spec(t, er, elems, detector_type="Si", escape_factor=0.05).numpy()
            out[f"{tag}_{set_name}_spec_ge"] = rmap.model_spec(t, er, elems, detector_type="Ge", escape_factor=0.02).numpy()
    out["model_elems"] = np.array(MODEL_ELEMS)
    out["elems10"], out["elems30"] = np.array(ELEMS10), np.array(ELEMS30)


def synth_spectrum(p, er_full, elems, rng, lo=1.5, hi=3.0, flat=2.0):
    amps = {e: float(rng.uniform(lo, hi)) for e in elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]}
    t = tens(p, amps)
    lam = rmap.model_spec(t, er_full, elems).numpy().astype(np.float64) + flat
    return rng.poisson(lam).astype(np.float32), amps


def gen_snip(out):
    rng = np.random.default_rng(11)
    cases = (("c2048", params_a(), (0, 2047), ELEMS10), ("c1401", params_a(), (50, 1450), ELEMS10),
             ("c4096", params_4096(), (0, 4095), ELEMS30))
    for tag, p, er, elems in cases:
        n_full = 2048 if tag != "c4096" else 4096
        specs = np.stack([synth_spectrum(p, (0, n_full - 1), elems, rng, flat=f)[0] for f in (0.0, 2.0, 40.0)])
        # a few awkward rows: all zero, single spike, negative values
        odd = np.zeros((3, n_full), dtype=np.float32)
        odd[1, n_full // 3] = 1000.0
        odd[2] = rng.normal(0, 3, n_full).astype(np.float32)
        specs = np.concatenate([specs, odd])
        crop = specs[:, er[0]: er[1] + 1]
        t = tens(p)
        args = (t["ENERGY_OFFSET"], t["ENERGY_SLOPE"], t["ENERGY_QUADRATIC"], t["SNIP_WIDTH"], t["FWHM_OFFSET"],
                t["FWHM_FANOPRIME"])
        batch = rbkg.snip_bkg(torch.tensor(crop), er, *args).numpy()
        single = rbkg.snip_bkg(torch.tensor(crop[1]), er, *args).numpy()
        assert np.array_equal(batch[1], single)
        out[f"{tag}_er"] = np.array(er)
        out[f"{tag}_spec"] = crop
        out[f"{tag}_bkg"] = batch
        out[f"{tag}_params"] = np.array([float(a) for a in args], dtype=np.float64)
        # per-pass gather indices, replaying bkg.py:83-112 / 57-62 with the reference's own ops
        n_ch = crop.shape[-1]
        local = torch.arange(n_ch, dtype=torch.float32)
        gidx = local + er[0]
        energy = args[0] + (args[1] * gidx) + (args[2] * (gidx ** 2))
        sig = torch.clamp((args[4] / 2.3548) ** 2 + energy * 2.96 * args[5], min=0.0)
        w = args[3] * 2.35 * torch.sqrt(sig) / args[1]
        tabs = []

        def idx(w):
            lo = torch.clamp(local - w, 0, n_ch - 1).to(torch.int64)
            hi = torch.clamp(local + w, 0, n_ch - 1).to(torch.int64)
            return torch.stack([lo, hi]).numpy().astype(np.int32)

        tabs += [idx(w), idx(w)]
        from mapstorch.constant import M_SQRT2
        while torch.max(w).item() >= 0.5:
            tabs.append(idx(w))
            w = w / M_SQRT2
        out[f"{tag}_lohi"] = np.stack(tabs)


def gen_host(out):
    rng = np.random.default_rng(5)
    p = params_a()
    spec, _ = synth_spectrum(p, (0, 2047), ELEMS10, rng)
    vol = rng.poisson(5.0, size=(3, 4, 2048)).astype(np.float32) + spec
    out["init_spec"], out["init_vol"] = spec, vol
    names = list(default_fitting_elems)
    out["init_names"] = np.array(names)
    out["init_amps_spec"] = np.array([float(ropt.init_elem_amps(e, spec, 12.0, p["ENERGY_OFFSET"], p["ENERGY_SLOPE"],
                                                               compton_angle=p["COMPTON_ANGLE"])) for e in names])
    out["init_amps_vol"] = np.stack([np.asarray(ropt.init_elem_amps(e, vol, 12.0, p["ENERGY_OFFSET"], p["ENERGY_SLOPE"],
                                                                    compton_angle=p["COMPTON_ANGLE"]), dtype=np.float64)
                                     * np.ones((3, 4)) for e in names])
    for tag, pp, er in (("a", params_a(), (0, 2047)), ("b", params_b(), (50, 1450)), ("c", params_4096(), (0, 4095))):
        rr = rutil.get_peak_ranges(names, pp["COHERENT_SCT_ENERGY"], pp["COMPTON_ANGLE"], pp["ENERGY_OFFSET"],
                                   pp["ENERGY_SLOPE"], pp["ENERGY_QUADRATIC"], er)
        out[f"ranges_{tag}_keys"] = np.array(list(rr.keys()))
        out[f"ranges_{tag}_vals"] = np.array(list(rr.values()), dtype=np.int64)
        t = tens(pp)
        ev = ev_of(t, er)
        bins = []
        for det, esc in (("Si", 1.73998), ("Ge", 9.886)):
            delta_e = torch.mean(ev[1:] - ev[:-1]).item()
            bins.append(int(round(esc / delta_e)))
        out[f"escape_bins_{tag}"] = np.array(bins)


def gen_fit(out, quick):
    torch.set_num_threads(8)
    rng = np.random.default_rng(3)
    runs = []
    p = params_a()
    for i, (lo, hi) in enumerate(((1.5, 3.0), (1.5, 3.0), (0.0, 2.0))):
        spec, amps = synth_spectrum(p, (0, 2047), ELEMS10, rng, lo, hi, flat=0.0)
        runs.append((f"e10_px{i}", p, (0, 2047), ELEMS10, spec, False, 500 if quick else 4000))
    spec, amps = synth_spectrum(p, (0, 2047), ELEMS10, rng, 1.5, 3.0, flat=3.0)
    runs.append(("e10_snip", p, (50, 1450), ELEMS10, spec, True, 500 if quick else 4000))
    if not quick:
        p4 = params_4096()
        spec, amps = synth_spectrum(p4, (0, 4095), ELEMS30, rng, 1.5, 3.0, flat=0.0)
        runs.append(("e30_px0", p4, (0, 4095), ELEMS30, spec, False, 4000))
    out["fit_runs"] = np.array([r[0] for r in runs])
    for tag, pp, er, elems, spec, use_snip, n_iter in runs:
        tensors, spec_fit, bkg, trace = ropt.fit_spec(
            spec, er, elements_to_fit=list(elems), init_param_vals=pp, tune_params=False, init_amp=True,
            use_snip=use_snip, loss="mse", optimizer="adam", n_iter=n_iter, progress_bar=False)
        names = list(elems) + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
        out[f"{tag}_spec"] = spec
        out[f"{tag}_er"] = np.array(er)
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_param_names"] = np.array(list(pp.keys()))
        out[f"{tag}_param_vals"] = np.array([pp[k] for k in pp], dtype=np.float64)
        out[f"{tag}_logamps"] = np.array([tensors[n].item() for n in names], dtype=np.float64)
        out[f"{tag}_spec_fit"] = spec_fit
        out[f"{tag}_bkg"] = bkg
        out[f"{tag}_trace"] = np.array(trace, dtype=np.float64)
        out[f"{tag}_use_snip"] = np.array(use_snip)
        out[f"{tag}_n_iter"] = np.array(n_iter)
        print(tag, "final loss", trace[-1], flush=True)


def gen_fit_c3(out):
    """BASELINE config 3's shape: 4096 channels, 30 elements + 2 scatter peaks, SNIP background -- two pixels
    (high- and low-count) through the reference's own fit_spec(tune_params=False, use_snip=True), Adam, 4000 steps."""
    torch.set_num_threads(8)
    rng = np.random.default_rng(33)
    p4 = params_4096()
    runs = []
    for i, (lo, hi, flat) in enumerate(((1.5, 3.0, 3.0), (0.3, 1.6, 1.0))):
        spec, amps = synth_spectrum(p4, (0, 4095), ELEMS30, rng, lo, hi, flat=flat)
        runs.append((f"e30_snip_px{i}", p4, (0, 4095), ELEMS30, spec, True, 4000))
    out["fit_runs"] = np.array([r[0] for r in runs])
    for tag, pp, er, elems, spec, use_snip, n_iter in runs:
        tensors, spec_fit, bkg, trace = ropt.fit_spec(
            spec, er, elements_to_fit=list(elems), init_param_vals=pp, tune_params=False, init_amp=True,
            use_snip=use_snip, loss="mse", optimizer="adam", n_iter=n_iter, progress_bar=False)
        names = list(elems) + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
        out[f"{tag}_spec"] = spec
        out[f"{tag}_er"] = np.array(er)
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_param_names"] = np.array(list(pp.keys()))
        out[f"{tag}_param_vals"] = np.array([pp[k] for k in pp], dtype=np.float64)
        out[f"{tag}_logamps"] = np.array([tensors[n].item() for n in names], dtype=np.float64)
        out[f"{tag}_spec_fit"] = spec_fit
        out[f"{tag}_bkg"] = bkg
        out[f"{tag}_trace"] = np.array(trace[::10] + trace[-1:], dtype=np.float64)
        out[f"{tag}_use_snip"] = np.array(use_snip)
        out[f"{tag}_n_iter"] = np.array(n_iter)
        print(tag, "final loss", trace[-1], flush=True)


JAC_PARAMS = ("ENERGY_OFFSET", "ENERGY_SLOPE", "FWHM_OFFSET", "FWHM_FANOPRIME", "ENERGY_QUADRATIC",
              "COHERENT_SCT_ENERGY", "COMPTON_ANGLE", "COMPTON_FWHM_CORR")


def gen_jacobian(out):
    """d model_spec / d param by torch.autograd through the reference's own model (what its optimiser differentiates,
    opt.py:249-317), in float64 with fp32-representable parameter values, for the eight calibration parameters that
    have a gradient on that path; with and without step terms, full and cropped energy range."""
    from torch.autograd.functional import jacobian

    elems = ["K", "Ca", "Fe", "Cu", "Zn", "Au_L", "Pt_M", "Si_Si"]
    rng = np.random.default_rng(17)
    cases = (("plain", params_a(), (0, 2047)), ("steps", dict(params_b(), COHERENT_SCT_ENERGY=12.5), (50, 1450)),
             ("steps4096", dict(params_4096(), F_STEP_OFFSET=0.003, F_STEP_LINEAR=0.0002, COMPTON_F_STEP=0.02), (0, 4095)))
    out["jac_cases"] = np.array([c[0] for c in cases])
    out["jac_params"] = np.array(JAC_PARAMS)
    out["jac_elems"] = np.array(elems)
    names = elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
    for tag, p, er in cases:
        p = {k: float(np.float32(v)) for k, v in p.items()}
        logamp = {n: float(np.float32(rng.uniform(1.5, 3.0))) for n in names}
        base = {k: torch.tensor(v, dtype=torch.float64) for k, v in p.items()}
        base.update({n: torch.tensor(v, dtype=torch.float64) for n, v in logamp.items()})

        def model(*theta):
            t = dict(base)
            t.update(dict(zip(JAC_PARAMS, theta)))
            return rmap.model_spec(t, er, elems)

        point = tuple(base[k].clone() for k in JAC_PARAMS)
        jac = jacobian(model, point)
        out[f"jac_{tag}_er"] = np.array(er)
        out[f"jac_{tag}_param_names"] = np.array(list(p.keys()))
        out[f"jac_{tag}_param_vals"] = np.array([p[k] for k in p], dtype=np.float64)
        out[f"jac_{tag}_amp_names"] = np.array(names)
        out[f"jac_{tag}_logamps"] = np.array([logamp[n] for n in names], dtype=np.float64)
        out[f"jac_{tag}_model"] = model(*point).numpy()
        out[f"jac_{tag}_jac"] = np.stack([j.numpy() for j in jac])
        print(tag, "max |J| per parameter", [float(np.abs(j.numpy()).max()) for j in jac], flush=True)


def gen_random_tables(out, n_sets=40):
    """Index arithmetic under randomly perturbed calibrations: CRC32 of every SNIP gather table, pass
    counts, escape bins, amplitude-guess windows and peak ranges, all computed by the reference's own
    code (bkg.py:57-62,83-112; map.py:142-146; opt.py:74-90; util.py:87-108)."""
    import zlib
    from mapstorch.constant import M_SQRT2

    rng = np.random.default_rng(2024)
    base = params_a()
    recs = []
    elems = ["Si", "Ca", "Fe", "Zn", "Au_L", "Pt_M", "COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
    for i in range(n_sets):
        n_ch = int(rng.choice([1024, 2048, 4096]))
        slope = float(rng.uniform(0.8, 1.25)) * (0.012 * 2048 / n_ch)
        p = dict(base, ENERGY_OFFSET=float(rng.uniform(-0.05, 0.05)), ENERGY_SLOPE=slope,
                 ENERGY_QUADRATIC=float(rng.uniform(-4e-8, 4e-8)) * (2048 / n_ch) ** 2,
                 FWHM_OFFSET=float(rng.uniform(0.06, 0.12)), FWHM_FANOPRIME=float(rng.uniform(1e-4, 3e-4)),
                 SNIP_WIDTH=float(rng.uniform(0.3, 1.5)), COHERENT_SCT_ENERGY=float(rng.uniform(9.0, 18.0)),
                 COMPTON_ANGLE=float(rng.uniform(80.0, 130.0)))
        er0 = int(rng.integers(0, 80))
        er1 = int(n_ch - 1 - rng.integers(0, 200))
        t = tens(p)
        n = er1 - er0 + 1
        local = torch.arange(n, dtype=torch.float32)
        gidx = local + er0
        energy = t["ENERGY_OFFSET"] + (t["ENERGY_SLOPE"] * gidx) + (t["ENERGY_QUADRATIC"] * (gidx ** 2))
        sig = torch.clamp((t["FWHM_OFFSET"] / 2.3548) ** 2 + energy * 2.96 * t["FWHM_FANOPRIME"], min=0.0)
        w = t["SNIP_WIDTH"] * 2.35 * torch.sqrt(sig) / t["ENERGY_SLOPE"]
        crc, n_pass = 0, 0

        def crc_of(w, crc):
            lo = torch.clamp(local - w, 0, n - 1).to(torch.int64)
            hi = torch.clamp(local + w, 0, n - 1).to(torch.int64)
            return zlib.crc32((lo | (hi << 16)).numpy().astype(np.uint32).tobytes(), crc)

        for _ in range(2):
            crc = crc_of(w, crc)
            n_pass += 1
        while torch.max(w).item() >= 0.5:
            crc = crc_of(w, crc)
            n_pass += 1
            w = w / M_SQRT2
        ev = ev_of(t, (er0, er1))
        delta_e = torch.mean(ev[1:] - ev[:-1]).item()
        bins = [int(round(1.73998 / delta_e)), int(round(9.886 / delta_e))]
        spec = np.zeros(n_ch, dtype=np.float32)
        wins = []
        for e in elems:
            # replay opt.py:74-90 through the reference function on a ramp spectrum: the mean identifies the window
            ramp = np.arange(n_ch, dtype=np.float64)
            g = ropt.init_elem_amps(e, ramp, p["COHERENT_SCT_ENERGY"], p["ENERGY_OFFSET"], p["ENERGY_SLOPE"],
                                    compton_angle=p["COMPTON_ANGLE"])
            wins.append(float(g))
        rr = rutil.get_peak_ranges(elems, p["COHERENT_SCT_ENERGY"], p["COMPTON_ANGLE"], p["ENERGY_OFFSET"],
                                   p["ENERGY_SLOPE"], p["ENERGY_QUADRATIC"], (er0, er1))
        rcrc = zlib.crc32(np.array(list(rr.values()), dtype=np.int64).tobytes())
        recs.append([n_ch, er0, er1, n_pass, crc, bins[0], bins[1], rcrc] + wins +
                    [p[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET",
                                    "FWHM_FANOPRIME", "COHERENT_SCT_ENERGY", "COMPTON_ANGLE")])
    out["rt_records"] = np.array(recs, dtype=np.float64)
    out["rt_elems"] = np.array(elems)


def gen_refine(out, n_iter=3000):
    """BASELINE config 4 analogue: per-spectrum fit with ENERGY_OFFSET and FWHM_OFFSET free
    (fit_spec(fitting_params=[...], tune_params=True), opt.py:201,119), run long enough to converge,
    plus model-Jacobian entries from torch.autograd for the analytic-Jacobian check."""
    torch.set_num_threads(8)
    rng = np.random.default_rng(17)
    p = params_a()
    truth = dict(p, ENERGY_OFFSET=p["ENERGY_OFFSET"] + 0.004, FWHM_OFFSET=p["FWHM_OFFSET"] * 1.05)
    spec, amps = synth_spectrum(truth, (0, 2047), ELEMS20, rng, 2.0, 3.2, flat=0.0)
    names = list(ELEMS20) + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
    tensors, spec_fit, bkg, trace = ropt.fit_spec(
        spec, (0, 2047), elements_to_fit=list(ELEMS20), fitting_params=["ENERGY_OFFSET", "FWHM_OFFSET"],
        init_param_vals=p, tune_params=True, init_amp=True, use_snip=False, loss="mse", optimizer="adam",
        n_iter=n_iter, progress_bar=False)
    out["r_spec"] = spec
    out["r_names"] = np.array(names)
    out["r_param_names"] = np.array(list(p.keys()))
    out["r_param_vals"] = np.array([p[k] for k in p], dtype=np.float64)
    out["r_truth_offset_width"] = np.array([truth["ENERGY_OFFSET"], truth["FWHM_OFFSET"]])
    out["r_fit_offset_width"] = np.array([tensors["ENERGY_OFFSET"].item(), tensors["FWHM_OFFSET"].item()])
    out["r_logamps"] = np.array([tensors[n].item() for n in names], dtype=np.float64)
    out["r_truth_logamps"] = np.array([amps[n] for n in names], dtype=np.float64)
    out["r_spec_fit"] = spec_fit
    out["r_trace"] = np.array(trace, dtype=np.float64)
    print("refine final loss", trace[-1], "offset/width", out["r_fit_offset_width"], flush=True)
    # Jacobian entries of model_spec: all log10 A = 3, defaults with E0 = 12
    t = tens(p, {n: 3.0 for n in names})
    for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "FWHM_OFFSET", "FWHM_FANOPRIME", "Fe", "Zn", "COMPTON_AMPLITUDE"):
        t[k].requires_grad_(True)
    m = rmap.model_spec(t, (0, 2047), list(ELEMS20))
    chans = [100, 300, 534, 535, 600, 720, 900, 1000, 1003]
    jac = []
    for c in chans:
        grads = torch.autograd.grad(m[c], [t[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "FWHM_OFFSET",
                                                          "FWHM_FANOPRIME", "Fe", "Zn", "COMPTON_AMPLITUDE")],
                                    retain_graph=True, allow_unused=True)
        jac.append([0.0 if g is None else g.item() for g in grads])
    out["j_channels"] = np.array(chans)
    out["j_values"] = np.array(jac)
    out["j_model"] = m.detach().numpy()


def gen_objectives(out, n_iter=6000):
    """fit_spec(tune_params=False) under the other objectives: loss="l1", the l1_lambda penalty, and a channel
    subset (`indices`).  Ten elements are in the spectrum, twenty are fitted, so the penalty has something to
    switch off; two channels carry outliers so that L1 and MSE disagree."""
    torch.set_num_threads(8)
    rng = np.random.default_rng(11)
    p = params_a()
    er = (0, 2047)
    spec, _ = synth_spectrum(p, er, ELEMS10, rng, 1.5, 3.0, flat=0.0)
    spec = spec.copy()
    spec[500] += 5000.0
    spec[900] += 3000.0
    idx = np.r_[100:480, 520:880, 920:1400]
    runs = [("l1", dict(loss="l1"), ELEMS10, None), ("mse_pen", dict(loss="mse", l1_lambda=1.5e-6), ELEMS20, None),
            ("l1_pen", dict(loss="l1", l1_lambda=1e-4), ELEMS20, None), ("mse_idx", dict(loss="mse"), ELEMS10, idx)]
    out["runs"] = np.array([r[0] for r in runs])
    out["spec"] = spec
    out["er"] = np.array(er)
    out["param_names"] = np.array(list(p.keys()))
    out["param_vals"] = np.array([p[k] for k in p], dtype=np.float64)
    for tag, kw, elems, indices in runs:
        tensors, spec_fit, bkg, trace = ropt.fit_spec(
            spec, er, elements_to_fit=list(elems), init_param_vals=p, tune_params=False, init_amp=True,
            use_snip=False, optimizer="adam", n_iter=n_iter, progress_bar=False, indices=indices, **kw)
        names = list(elems) + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
        out[f"{tag}_names"] = np.array(names)
        out[f"{tag}_kwargs"] = np.array(repr(kw))
        out[f"{tag}_indices"] = np.array([] if indices is None else indices, dtype=np.int64)
        out[f"{tag}_logamps"] = np.array([tensors[n].item() for n in names], dtype=np.float64)
        out[f"{tag}_spec_fit"] = spec_fit
        out[f"{tag}_trace"] = np.array(trace, dtype=np.float64)
        print(tag, "final loss", trace[-1], "min", min(trace), flush=True)


OVERRIDE_ELEMS = ["Fe", "Gd_L", "Pt_L", "Sn_L", "I_L", "Si_Si"]
OVERRIDE_EXTRA = ("BRANCHING_RATIO_ADJUSTMENT_K: Pb_L, 1., 0.5, 2.0, 1.\n"
                  "BRANCHING_FAMILY_ADJUSTMENT_L: Au_L, 0.5, 0.7, 1.1\n"
                  "BRANCHING_RATIO_ADJUSTMENT_K: Fe, 1., 1., 1.2, 1.\n"
                  "BRANCHING_RATIO_ADJUSTMENT_L: Nope_L, 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1., 1.\n"
                  "BRANCHING_RATIO_ADJUSTMENT_L: W_L, 1., 0.9, 0.8\n")


def consts_arrays(consts):
    names = list(consts.keys())
    num = np.array([[[s.energy, s.ratio, s.Ge_mu, s.Si_mu, s.mu_fraction, s.width_multi] for s in consts[k]]
                    for k in names], dtype=np.float64)
    kinds = np.array([[s.type for s in consts[k]] for k in names], dtype=np.int64)
    ptypes = np.array([[s.ptype for s in consts[k]] for k in names])
    pile = np.array([[s.is_pileup for s in consts[k]] for k in names])
    slot_names = np.array([[s.name for s in consts[k]] for k in names])
    return {"names": np.array(names), "num": num, "type": kinds, "ptype": ptypes, "is_pileup": pile,
            "slot_names": slot_names}


def gen_override(out):
    import json
    import re
    import tempfile

    info = rconst.MapsElements().get_element_info()
    tmp = Path(tempfile.mkdtemp())
    cases = {
        "A": dict(param_values=None, elements=None, detector_material="Si", escape_factor=0.0),
        "B": dict(param_values={"ENERGY_SLOPE": 0.0123, "CAL_OFFSET_[E_OFFSET]": 7, "COMPTON_STEP": 0.25,
                                "KB_F_TAIL_LINEAR": 3e-4, "NOT_A_TAG": 1.5, "COHERENT_SCT_ENERGY": 12.0},
                  elements=["Fe", "Cu", "Pb_L", "Au_M", "Si_Si", "COMPTON_AMPLITUDE"], detector_material="Ge",
                  escape_factor=0.3),
        "D": dict(param_values={"SNIP_WIDTH": 0.75}, elements=["Ca"], detector_material="Si", escape_factor=0.07),
    }
    texts = {}
    for tag, kw in cases.items():
        path = tmp / f"{tag}.txt"
        rio.write_override_params_file(str(path), **kw)
        texts[tag] = path.read_text()
        out[f"{tag}_kwargs"] = np.array(json.dumps(kw))
    # C: case B edited by hand -- pile-ups listed, more adjustment rows (incl. unknown names, short rows)
    texts["C"] = texts["B"].replace("ELEMENTS_WITH_PILEUP: \n", "ELEMENTS_WITH_PILEUP: Si_Si, Fe_Fe, Ca_K\n").replace(
        "SI_ESCAPE_FACTOR:", OVERRIDE_EXTRA + "SI_ESCAPE_FACTOR:")
    assert texts["C"] != texts["B"] and OVERRIDE_EXTRA in texts["C"]
    t = tens(params_a(), {e: 0.0 for e in default_fitting_elems})
    ev = ev_of(t, (0, 2047))
    for tag, text in texts.items():
        path = tmp / f"{tag}.txt"
        path.write_text(text)
        out[f"{tag}_text"] = np.array(re.sub(r"DATE: .*", "DATE: <date>", text))
        out[f"{tag}_parsed"] = np.array(json.dumps(rio.parse_override_params_file(str(path))))
        consts = rconst.read_constants(str(path), info)
        for k, v in consts_arrays(consts).items():
            out[f"{tag}_consts_{k}"] = v
        present = [e for e in OVERRIDE_ELEMS if e in consts]
        out[f"{tag}_spectra_elems"] = np.array(present)
        out[f"{tag}_spectra"] = np.stack([rmap.model_elem_spec(t, e, ev, True, True, e_consts=consts).numpy()
                                          for e in present])
    out["spectra_ev"] = ev.numpy()
    # rows + names form, with pile-ups passed by the caller
    rows, names = rconst.read_constants(str(tmp / "A.txt"), info, pileups=["Si_Cl"], return_dict=False)
    out["A_rows_names"] = np.array([str(n) for n in names])
    out["A_rows_ratio"] = np.array([[s.ratio for s in r] for r in rows])
    # the two tail tags raise upstream (attribute access on an object-array slice)
    for tag_line in ("TAIL_FRACTION_ADJUST_SI: 1.5\n", "TAIL_WIDTH_ADJUST_SI: 2.0\n"):
        path = tmp / "tail.txt"
        path.write_text(texts["A"] + "\n" + tag_line)
        try:
            rconst.read_constants(str(path), info)
            raised = "none"
        except Exception as exc:  # noqa: BLE001
            raised = type(exc).__name__
        out["raises_" + tag_line.split(":")[0]] = np.array(raised)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-fit", action="store_true")
    ap.add_argument("--quick-fit", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    jobs = {"model": gen_model, "snip": gen_snip, "host": gen_host, "tables": gen_random_tables,
            "override": gen_override}
    for name, fn in jobs.items():
        if args.only and args.only != name:
            continue
        out = {}
        fn(out)
        np.savez_compressed(HERE / f"{name}_golden.npz", **out)
        print("wrote", name, (HERE / f"{name}_golden.npz").stat().st_size, flush=True)
    if args.only == "objectives":
        out = {}
        gen_objectives(out)
        np.savez_compressed(HERE / "objectives_golden.npz", **out)
        print("wrote objectives", (HERE / "objectives_golden.npz").stat().st_size, flush=True)
        return
    if args.only == "jacobian":
        out = {}
        gen_jacobian(out)
        np.savez_compressed(HERE / "jacobian_golden.npz", **out)
        print("wrote jacobian", (HERE / "jacobian_golden.npz").stat().st_size, flush=True)
        return
    if args.only == "fit_c3":
        out = {}
        gen_fit_c3(out)
        np.savez_compressed(HERE / "fit_c3_golden.npz", **out)
        print("wrote fit_c3", (HERE / "fit_c3_golden.npz").stat().st_size, flush=True)
        return
    if args.only == "refine":
        out = {}
        gen_refine(out)
        np.savez_compressed(HERE / "refine_golden.npz", **out)
        print("wrote refine", (HERE / "refine_golden.npz").stat().st_size, flush=True)
        return
    if not args.skip_fit and args.only in ("", "fit"):
        out = {}
        gen_fit(out, args.quick_fit)
        np.savez_compressed(HERE / "fit_golden.npz", **out)
        print("wrote fit", (HERE / "fit_golden.npz").stat().st_size, flush=True)


if __name__ == "__main__":
    main()
