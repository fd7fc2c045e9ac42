"""Host tables of the per-pixel refinement kernel (mapstorch_b200.refine.refine_tables): pure host logic,
no device.  The kernel trusts these invariants (mt_refine validates only sizes and the permutation)."""
import numpy as np

from mapstorch_b200 import refine, synth, tables
from mapstorch_b200.default import default_energy_consts


def _tables(elems, free=("ENERGY_OFFSET", "FWHM_OFFSET"), n_ch=2048, **kw):
    p = synth.benchmark_params(n_ch)
    names = tables.fitted_components(elems, default_energy_consts)
    return names, refine.refine_tables(p, names, (0, n_ch - 1), free, **kw)


def test_tables_cover_every_line_and_component():
    names, (cfg, lines, chan_lines, comp_start, comp_lines, window, comp_order) = _tables(synth.ELEMS20)
    n_lines = cfg.n_lines
    assert cfg.n_comp == len(names) and cfg.n_ch == 2048 and cfg.n_theta == 2
    energies = [lines[i].energy for i in range(n_lines)]
    assert energies == sorted(energies)                       # the channel ranges rely on energy order
    assert sorted(comp_lines.tolist()) == list(range(n_lines))  # every line belongs to exactly one component
    assert comp_start[0] == 0 and comp_start[-1] == n_lines and np.all(np.diff(comp_start) >= 1)
    for e in range(cfg.n_comp):
        for j in range(comp_start[e], comp_start[e + 1]):
            assert lines[int(comp_lines[j])].comp == e
    lo, hi = window & 0xFFFF, window >> 16
    assert np.all(lo < hi) and np.all(hi <= 2048)
    # a channel's line range is the union of the windows that contain it
    first, last = chan_lines & 0xFFFF, chan_lines >> 16
    for c in (0, 300, 533, 1000, 1900):
        inside = [i for i in range(n_lines) if lo[i] <= c < hi[i]]
        if inside:
            assert first[c] == min(inside) and last[c] == max(inside) + 1
        else:
            assert first[c] == last[c]


def test_component_schedule_is_a_balanced_permutation():
    names, (cfg, lines, chan_lines, comp_start, comp_lines, window, comp_order) = _tables(synth.ELEMS30, n_ch=4096)
    assert sorted(comp_order.tolist()) == list(range(len(names)))
    length = (window >> 16).astype(np.int64) - (window & 0xFFFF).astype(np.int64)
    cost = np.zeros(len(names))
    for i in range(cfg.n_lines):
        cost[lines[i].comp] += length[i] + 16
    per_warp = [cost[comp_order[w::refine.REFINE_WARPS]].sum() for w in range(refine.REFINE_WARPS)]
    natural = [cost[np.arange(len(names))[w::refine.REFINE_WARPS]].sum() for w in range(refine.REFINE_WARPS)]
    assert max(per_warp) <= max(natural) + 1e-9          # never worse than taking the components in order
    assert max(per_warp) <= 1.25 * np.mean(per_warp)     # and close to even


def test_windows_scale_with_k_sigma_and_bounds():
    _, small = _tables(synth.ELEMS10, k_sigma=4.0)
    _, large = _tables(synth.ELEMS10, k_sigma=7.0)
    w_small = ((small[5] >> 16) - (small[5] & 0xFFFF)).astype(int)
    w_large = ((large[5] >> 16) - (large[5] & 0xFFFF)).astype(int)
    assert np.all(w_large >= w_small) and w_large.sum() > 1.3 * w_small.sum()
    _, tight = _tables(synth.ELEMS10, bounds={"ENERGY_OFFSET": ("abs", 0.0), "FWHM_OFFSET": ("rel", 0.0)})
    w_tight = ((tight[5] >> 16) - (tight[5] & 0xFFFF)).astype(int)
    default = ((_tables(synth.ELEMS10)[1][5] >> 16) - (_tables(synth.ELEMS10)[1][5] & 0xFFFF)).astype(int)
    assert np.all(w_tight <= default) and w_tight.sum() < default.sum()
    cfg = tight[0]
    assert cfg.theta_lo[0] == cfg.theta0[0] == cfg.theta_hi[0]
