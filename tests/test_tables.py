"""Host-side table compiler (mapstorch_b200.tables / util / opt helpers) against the reference
goldens: index tables are bit-exact, compiled term tables reproduce the reference columns."""
import numpy as np
import pytest
import torch

from conftest import golden_params
from mapstorch_b200 import default as d, opt, tables, util
from oracle import maps_oracle as mo


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_energy_axis_bit_exact(model_golden, tag):
    p = golden_params(model_golden, tag)
    ev = tables.energy_axis(p, tuple(model_golden[f"{tag}_er"]))
    assert ev.dtype == torch.float32 and np.array_equal(ev.numpy(), model_golden[f"{tag}_ev"])
    tens = {k: torch.tensor(float(v)) for k, v in p.items()}
    assert np.array_equal(tables.energy_axis(tens, tuple(model_golden[f"{tag}_er"])).numpy(), model_golden[f"{tag}_ev"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_compiled_terms_reproduce_reference_columns(model_golden, tag):
    g = model_golden
    p = golden_params(g, tag)
    elems = [str(e) for e in g["model_elems"]]
    ev = g[f"{tag}_ev"]
    comps = elems + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]
    for amp in (0.0, 2.3):
        pp = dict(p, **{c: amp for c in comps})
        for us in (0, 1):
            for ut in (0, 1):
                key = f"{tag}_amp{amp}_s{us}_t{ut}"
                if key + "_elems" not in g:
                    continue
                cols = tables.compile_terms(pp, comps, d.default_energy_consts, d.default_elem_info, bool(us), bool(ut),
                                            log_amps=pp)
                terms, starts = tables.pack_columns(cols)
                out = mo.eval_terms(terms, starts, ev)
                refs = list(g[key + "_elems"]) + [g[key + "_compton"], g[f"{tag}_amp{amp}_elastic"]]
                for name, o, r in zip(comps, out, refs):
                    scale = np.abs(r).max()
                    if scale == 0:
                        assert not o.any(), (key, name)
                    else:
                        assert np.abs(o - r).max() <= 1e-6 * scale, (key, name)


def test_term_counts_for_benchmark_set(model_golden):
    # the 30-element benchmark set: 136 lines with a non-zero ratio, of which the 12 keV gate opens 115
    p = golden_params(model_golden, "c")
    elems = [str(e) for e in model_golden["elems30"]]
    names = tables.fitted_components(elems, d.default_energy_consts)
    assert names[-2:] == ["COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE"] and len(names) == 32
    cols = tables.compile_terms(p, names, d.default_energy_consts, d.default_elem_info)
    n_lines = sum(1 for n in names[:-2] for s in d.default_energy_consts[n] if s.ratio != 0 and s.energy > 0)
    assert n_lines == 136
    n_open = sum(len(c) for c in cols[:-2])  # lines whose absorption edge lies below 12 keV
    assert n_open == 115
    terms, starts = tables.pack_columns(cols)
    assert terms.dtype.itemsize == 32 and starts[-1] == n_open + 2


@pytest.mark.parametrize("tag", ["c2048", "c1401", "c4096"])
def test_snip_tables_bit_exact(snip_golden, tag):
    g = snip_golden
    er, pr, lohi = tuple(g[f"{tag}_er"]), g[f"{tag}_params"], g[f"{tag}_lohi"]
    n_ch = g[f"{tag}_spec"].shape[-1]
    for args in ([torch.tensor(float(v)) for v in pr], [float(np.float32(v)) for v in pr]):
        tab = tables.snip_tables(n_ch, er[0], *args)
        assert tab.dtype == np.uint32 and tab.shape == (lohi.shape[0], n_ch)
        assert np.array_equal(tab & 0xFFFF, lohi[:, 0]) and np.array_equal(tab >> 16, lohi[:, 1])
    # last pass is the asymmetric lo = i-1 / hi = i window the truncation produces
    last = tab[-1]
    i = np.arange(n_ch)
    assert np.all((i - (last & 0xFFFF)) <= 1) and np.all((last >> 16) == i)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_escape_bins_and_peak_ranges(host_golden, model_golden, tag):
    p = golden_params(model_golden, tag)
    er = tuple(int(v) for v in model_golden[f"{tag}_er"])
    ev = tables.energy_axis(p, er)
    want = host_golden[f"escape_bins_{tag}"]
    for det, b in zip(("Si", "Ge"), want):
        assert tables.escape_bins(ev, det) == (int(b) if b < ev.numel() else 0)
    rr = util.get_peak_ranges(list(d.default_fitting_elems), p["COHERENT_SCT_ENERGY"], p["COMPTON_ANGLE"],
                              p["ENERGY_OFFSET"], p["ENERGY_SLOPE"], p["ENERGY_QUADRATIC"], er)
    keys = [str(k) for k in host_golden[f"ranges_{tag}_keys"]]
    assert list(rr.keys()) == keys
    assert np.array_equal(np.array(list(rr.values())), host_golden[f"ranges_{tag}_vals"])
    idx = util.peak_range_indices(rr, er[1] - er[0] + 1)
    assert idx.size and idx.max() <= er[1] - er[0]


def test_init_elem_amps_matches_reference(host_golden):
    g = host_golden
    dp = d.default_param_vals
    names = [str(n) for n in g["init_names"]]
    got = np.array([float(opt.init_elem_amps(e, g["init_spec"], 12.0, dp["ENERGY_OFFSET"], dp["ENERGY_SLOPE"],
                                             compton_angle=dp["COMPTON_ANGLE"])) for e in names])
    assert np.array_equal(got, g["init_amps_spec"])
    vol = np.stack([np.asarray(opt.init_elem_amps(e, g["init_vol"], 12.0, dp["ENERGY_OFFSET"], dp["ENERGY_SLOPE"],
                                                  compton_angle=dp["COMPTON_ANGLE"])) * np.ones((3, 4)) for e in names])
    assert np.array_equal(vol, g["init_amps_vol"])
    assert opt.init_elem_amps("NotAnElement", g["init_spec"], 12.0, 0.0, 0.012) == 0.0


def test_create_tensors_mirrors_reference_layout():
    tensors, groups = opt.create_tensors(["Fe", "Cu"], fixed_param_vals={"ENERGY_SLOPE": 0.0123})
    assert set(tensors) == set(d.default_param_vals) | {"Fe", "Cu"}
    assert not tensors["ENERGY_SLOPE"].requires_grad and tensors["ENERGY_OFFSET"].requires_grad
    assert tensors["Fe"].requires_grad and float(tensors["Fe"]) == 0.0
    assert len(groups) == 11 + 2 and groups[-1]["lr"] == 1e-2
    t2, g2 = opt.create_tensors(["Fe"], tune_params=False)
    assert [k for k, v in t2.items() if v.requires_grad] == ["Fe"] and len(g2) == 1


def test_digit_planes_hold_30_bits_in_fp16_exact_integers():
    rng = np.random.default_rng(0)
    w = rng.normal(size=(12, 512)) * np.exp(rng.uniform(-12, 3, size=(12, 512)))
    w[3] = 0.0
    scale, planes = opt.digit_planes(w)
    assert planes.shape == (3, 12, 512) and np.abs(planes).max() <= 512
    assert np.array_equal(planes, np.rint(planes)) and np.array_equal(planes.astype(np.float16).astype(np.float64), planes)
    rebuilt = scale[:, None].astype(np.float64) * sum(opt.PLANE_RATIO ** d * planes[d] for d in range(3))
    peak = np.abs(w).max(axis=1, keepdims=True)
    peak[peak == 0] = 1
    assert (np.abs(rebuilt - w) / peak).max() <= 2.0 ** -29
    assert np.all(np.log2(scale) == np.rint(np.log2(scale)))  # power-of-two row scales: exact


def test_index_tables_under_random_calibrations():
    """40 randomly perturbed calibrations (tables_golden.npz, computed by the reference): SNIP gather tables
    (CRC32 over all passes), pass counts, escape bins, amplitude-guess windows and peak ranges are exact --
    for the product's table compiler AND for the oracle restatement."""
    import zlib
    from conftest import GOLDEN
    from mapstorch_b200.default import default_param_vals

    g = np.load(GOLDEN / "tables_golden.npz")
    elems = [str(e) for e in g["rt_elems"]]
    for rec in g["rt_records"]:
        n_ch, er0, er1, n_pass, crc, bin_si, bin_ge, rcrc = (int(v) for v in rec[:8])
        wins = rec[8: 8 + len(elems)]
        off, slope, quad, snipw, fwhm, fano, e0, angle = rec[8 + len(elems):]
        n = er1 - er0 + 1
        args = [torch.tensor(float(v)) for v in (off, slope, quad, snipw, fwhm, fano)]
        tab = tables.snip_tables(n, er0, *args)
        assert tab.shape[0] == n_pass and zlib.crc32(tab.tobytes()) == crc
        passes = mo.snip_pass_tables(n, er0, off, slope, quad, snipw, fwhm, fano)
        assert len(passes) == n_pass
        assert zlib.crc32(np.stack([(lo | (hi << 16)).astype(np.uint32) for lo, hi in passes]).tobytes()) == crc
        p = dict(default_param_vals, ENERGY_OFFSET=off, ENERGY_SLOPE=slope, ENERGY_QUADRATIC=quad)
        ev = tables.energy_axis(p, (er0, er1))
        for det, want in (("Si", bin_si), ("Ge", bin_ge)):
            assert tables.escape_bins(ev, det) == (want if want < n else 0)
        ramp = np.arange(n_ch, dtype=np.float64)
        got = [float(opt.init_elem_amps(e, ramp, e0, off, slope, compton_angle=angle)) for e in elems]
        assert np.array_equal(np.array(got), wins)
        rr = util.get_peak_ranges(elems, e0, angle, off, slope, quad, (er0, er1))
        assert zlib.crc32(np.array(list(rr.values()), dtype=np.int64).tobytes()) == rcrc


def test_snip_tables_numpy_builder_equals_the_torch_op_sequence():
    """tables.snip_tables forms the SNIP window indices with NumPy; tables.snip_tables_torch is the same computation
    written with the reference's own torch CPU ops (bkg.py:57-62, 83-112).  They must agree bit for bit -- for Python
    float parameters, 0-dim tensor parameters and mixtures (the promotion rules differ in where double arithmetic
    happens), every window size the kernels serve, and widths that make the pass count vary."""
    import torch

    from mapstorch_b200 import tables

    rng = np.random.default_rng(7)
    for trial in range(300):
        n_ch = int(rng.choice([8, 512, 1401, 2048, 4096]))
        ch0 = int(rng.integers(0, 100))
        vals = [rng.uniform(-0.05, 0.05), rng.uniform(0.004, 0.02), rng.uniform(-5e-8, 5e-8), rng.uniform(0.1, 1.5),
                rng.uniform(0.0, 0.25), rng.uniform(0, 3e-4)]
        mode = trial % 3
        args = [torch.tensor(float(v)) if (mode == 1 or (mode == 2 and rng.random() < 0.5)) else float(v) for v in vals]
        ref = tables.snip_tables_torch(n_ch, ch0, *args)
        got = tables.snip_tables(n_ch, ch0, *args)
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref), (trial, n_ch, ch0, vals)


def test_vectorised_term_compiler_is_bit_identical_to_the_scalar_route():
    """tables.compile_packed (array arithmetic over all lines, used by model_spec / element_basis) against
    pack_columns(compile_terms(...)) (the reference's scalar op sequence, pinned above against the goldens): 200 random
    calibrations, random component subsets, incident energies exactly on absorption edges, with and without steps."""
    from mapstorch_b200 import default as d, tables

    rng = np.random.default_rng(0)
    ec, ei = d.default_energy_consts, d.default_elem_info
    for trial in range(200):
        pp = dict(d.default_param_vals)
        for k in pp:
            if isinstance(pp[k], float) and rng.random() < 0.7:
                pp[k] = float(np.float32(pp[k] * (1 + 0.2 * rng.standard_normal())))
        pp["COHERENT_SCT_ENERGY"] = float(rng.uniform(1.5, 30.0))
        if trial % 5 == 0:
            info = ei[int(rng.integers(10, 60))]
            pp["COHERENT_SCT_ENERGY"] = float(np.float32(list(info.bindingE.values())[0]))
        elems = d.default_fitting_elems if trial % 2 == 0 else list(
            rng.choice(d.default_fitting_elems, size=int(rng.integers(1, 40)), replace=False))
        comps = tables.fitted_components(elems, ec)
        log_amps = {c: float(rng.uniform(-2, 3)) for c in comps} if trial % 3 == 0 else None
        for use_step in (True, False):
            want_t, want_s = tables.pack_columns(tables.compile_terms(pp, comps, ec, ei, use_step, False, log_amps))
            got_t, got_s = tables.compile_packed(pp, comps, ec, ei, use_step, False, log_amps)
            assert np.array_equal(want_s, got_s)
            assert want_t.tobytes() == got_t.tobytes()
    # column orders other than model_spec's (scatter peaks first / in between / alone)
    for comps in (["COMPTON_AMPLITUDE", "Fe", "COHERENT_SCT_AMPLITUDE", "Cu"], ["Fe", "Cu"], ["COHERENT_SCT_AMPLITUDE"],
                  ["COMPTON_AMPLITUDE", "Zn"]):
        want_t, want_s = tables.pack_columns(tables.compile_terms(pp, comps, ec, ei, True, False))
        got_t, got_s = tables.compile_packed(pp, comps, ec, ei, True, False)
        assert np.array_equal(want_s, got_s) and want_t.tobytes() == got_t.tobytes()
    # tail terms take the scalar route
    want_t, want_s = tables.pack_columns(tables.compile_terms(pp, comps, ec, ei, True, True))
    got_t, got_s = tables.compile_packed(pp, comps, ec, ei, True, True)
    assert np.array_equal(want_s, got_s) and want_t.tobytes() == got_t.tobytes()


def test_snip_pass_count_follows_from_the_largest_width():
    """bkg.snip_tables_device builds the window table on the device and needs only the pass count from the host:
    tables.snip_pass_count(max width) must equal the number of rows of the full host table (every pass divides all
    widths by the same constant and rounding is monotone)."""
    import torch

    from mapstorch_b200 import tables

    rng = np.random.default_rng(3)
    for trial in range(200):
        n_ch = int(rng.integers(8, 4096))
        ch0 = int(rng.integers(0, 200))
        six = [torch.tensor(float(v)) for v in (rng.normal(0, 0.05), rng.uniform(0.003, 0.03), rng.normal(0, 1e-7),
                                                 rng.uniform(0.05, 4.0), rng.uniform(0.02, 0.4), rng.uniform(1e-6, 2e-3))]
        full = tables.snip_tables(n_ch, ch0, *six)
        width = tables.snip_widths(n_ch, ch0, *six)
        assert tables.snip_pass_count(width.max()) == full.shape[0], trial
        assert np.array_equal(full, tables.snip_tables_torch(n_ch, ch0, *six)) or trial % 20  # spot check vs the torch ops
