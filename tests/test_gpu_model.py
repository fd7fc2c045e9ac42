"""K1 (mt_model_terms / mt_model_sum) on the GPU vs the oracle and the reference goldens.
Tolerance: 1e-5 of the column maximum (north_star: model spectra within 1e-5)."""
import numpy as np
import pytest
import torch

from conftest import golden_params

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def mods():
    from mapstorch_b200 import map as mmap, tables, default
    from oracle import maps_oracle as mo

    return mmap, tables, default, mo


def tens(p):
    return {k: torch.tensor(float(v)) for k, v in p.items()}


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_model_elem_spec_all_flag_combinations(model_golden, mods, tag):
    mmap, tables, d, mo = mods
    g = model_golden
    p = golden_params(g, tag)
    elems = [str(e) for e in g["model_elems"]]
    ev = torch.from_numpy(g[f"{tag}_ev"])
    worst = 0.0
    for amp in (0.0, 2.3):
        pp = tens(dict(p, COMPTON_AMPLITUDE=amp, COHERENT_SCT_AMPLITUDE=amp, **{e: amp for e in elems}))
        for us in (0, 1):
            for ut in (0, 1):
                key = f"{tag}_amp{amp}_s{us}_t{ut}"
                if key + "_elems" not in g:
                    continue
                for i, e in enumerate(elems):
                    out = mmap.model_elem_spec(pp, e, ev, bool(us), bool(ut), device="cpu").numpy()
                    ref = g[key + "_elems"][i]
                    scale = np.abs(ref).max()
                    if scale == 0:
                        assert not out.any()
                    else:
                        worst = max(worst, np.abs(out - ref).max() / scale)
                rc = g[key + "_compton"]
                out = mmap.compton_peak(pp, ev, pp["ENERGY_SLOPE"], bool(us), bool(ut)).numpy()
                worst = max(worst, np.abs(out - rc).max() / rc.max())
        re_ = g[f"{tag}_amp{amp}_elastic"]
        worst = max(worst, np.abs(mmap.elastic_peak(pp, ev, pp["ENERGY_SLOPE"]).numpy() - re_).max() / re_.max())
    assert worst <= TOL, worst


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("eset", ["e10", "e30"])
def test_model_spec_with_escape(model_golden, mods, tag, eset):
    mmap, tables, d, mo = mods
    g = model_golden
    p = golden_params(g, tag)
    elems = [str(e) for e in g["elems10" if eset == "e10" else "elems30"]]
    names = elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
    pp = tens(dict(p, **dict(zip(names, g[f"{tag}_{eset}_logamps"]))))
    er = tuple(int(v) for v in g[f"{tag}_er"])
    for suffix, kw in (("spec", {}), ("spec_si", dict(detector_type="Si", escape_factor=0.05)),
                       ("spec_ge", dict(detector_type="Ge", escape_factor=0.02))):
        ref = g[f"{tag}_{eset}_{suffix}"]
        out = mmap.model_spec(pp, er, elems, device="cpu", **kw).numpy()
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= TOL * ref.max(), (tag, eset, suffix)
    basis, bnames = mmap.element_basis(pp, er, elems, detector_type="Si", escape_factor=0.05)
    amps = torch.tensor([10.0 ** float(pp[n]) for n in bnames], dtype=torch.float64, device="cuda")
    recon = (amps[None, :] @ basis.double()).reshape(-1).cpu().numpy()
    ref = g[f"{tag}_{eset}_spec_si"]
    assert np.abs(recon - ref).max() <= TOL * ref.max()


def test_kernel_matches_oracle_term_evaluator_closely(model_golden, mods):
    """Same tables, two evaluators: CUDA expf/erfcf vs NumPy -- a few ulp apart."""
    mmap, tables, d, mo = mods
    p = golden_params(model_golden, "b")
    names = tables.fitted_components([e for e in d.default_fitting_elems], d.default_energy_consts)
    cols = tables.compile_terms(p, names, d.default_energy_consts, d.default_elem_info, True, True)
    ev = tables.energy_axis(p, (0, 2047))
    out = mmap.run_terms(ev, cols).cpu().numpy()
    terms, starts = tables.pack_columns(cols)
    ref = mo.eval_terms(terms, starts, ev.numpy())
    scale = np.abs(ref).max(axis=1)
    live = scale > 0
    assert live.sum() > 60
    assert (np.abs(out - ref).max(axis=1)[live] / scale[live]).max() <= 2e-6
    assert not out[~live].any()


def test_shape_helpers(mods):
    mmap, tables, d, mo = mods
    delta = torch.linspace(-1.0, 1.0, 1001)
    gain, sigma, gamma = torch.tensor(0.012), torch.tensor(0.07), torch.tensor(2.2)
    d32 = delta.numpy()
    assert np.abs(mmap.peak(gain, sigma, delta).numpy() - mo.peak(np.float32(0.012), np.float32(0.07), d32)).max() <= 1e-6 * 0.07
    st = mo.step(np.float32(0.012), np.float32(0.07), d32, 6.4)
    assert np.abs(mmap.step(gain, sigma, delta, 6.4).numpy() - st).max() <= 2e-6 * st.max()
    tl = mo.tail(np.float32(0.012), np.float32(0.07), d32, np.float32(2.2))
    assert np.abs(mmap.tail(gain, sigma, delta, gamma).numpy() - tl).max() <= 2e-6 * tl.max()
    spec = torch.arange(100, dtype=torch.float32)
    ev = torch.arange(100, dtype=torch.float32) * 0.5
    esc = mmap.escape_peak(spec, ev, 0.1, "cpu", "Si")
    assert torch.equal(esc, torch.from_numpy(mo.escape_peak(spec.numpy(), ev.numpy(), 0.1, "Si")))


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_element_spectra_with_override_file_constants(tmp_path, override_golden, model_golden, mods, tag):
    """e_consts edited by a MAPS override file (branching ratios, Ge absorption column, pile-ups from
    the file) flow through K1: spectra vs the reference evaluated with ITS read_constants output."""
    from mapstorch_b200 import constant

    mmap, tables, d, mo = mods
    g = override_golden
    path = tmp_path / "o.txt"
    path.write_text(str(g[f"{tag}_text"]))
    consts = constant.read_constants(str(path), d.default_elem_info)
    p = golden_params(model_golden, "a")
    elems = [str(e) for e in g[f"{tag}_spectra_elems"]]
    pp = tens(dict(p, **{e: 0.0 for e in elems}))
    ev = torch.from_numpy(g["spectra_ev"])
    for i, e in enumerate(elems):
        got = mmap.model_elem_spec(pp, e, ev, True, True, e_consts=consts).numpy()
        ref = g[f"{tag}_spectra"][i]
        assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), (tag, e)
    # and the edit is visible: Fe K-beta is 1.2x the default table's
    if tag == "A":
        base = mmap.model_elem_spec(pp, "Fe", ev, True, True).numpy()
        kb = int(np.argmin(np.abs(g["spectra_ev"] - 7.058)))
        assert got is not None and abs(mmap.model_elem_spec(pp, "Fe", ev, True, True, e_consts=consts).numpy()[kb]
                                       / base[kb] - 1.2) < 0.02


def test_device_energy_axis_is_bit_identical_to_the_host_axis():
    """mt_energy_axis (what model_spec / element_basis evaluate the lines on) against tables.energy_axis, the reference's
    torch CPU op sequence (map.py:271-281)."""
    import torch

    from mapstorch_b200 import map as mmap, tables

    rng = np.random.default_rng(11)
    for trial in range(100):
        lo = int(rng.integers(0, 500))
        hi = lo + int(rng.integers(1, 4096))
        p = {"ENERGY_OFFSET": float(rng.normal(0, 0.1)), "ENERGY_SLOPE": float(rng.uniform(0.002, 0.04)),
             "ENERGY_QUADRATIC": float(rng.normal(0, 1e-6)) if trial % 4 else 0.0}
        if trial % 7 == 0:
            p = {k: torch.tensor(v) for k, v in p.items()}
        want = tables.energy_axis(p, (lo, hi))
        got = mmap._energy_axis(p, (lo, hi)).cpu()
        assert torch.equal(want, got), trial
