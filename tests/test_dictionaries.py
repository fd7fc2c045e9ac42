"""The parameter/element dictionaries (boundary data) against the reference dump."""
import json
from pathlib import Path

import torch

from mapstorch_b200 import constant, default

DOC = json.loads((Path(constant.__file__).parent / "data" / "xrf_tables.json").read_text())
FIELDS = ("name", "energy", "ratio", "Ge_mu", "Si_mu", "mu_fraction", "width_multi", "type", "ptype", "is_pileup")


def test_line_tables_match_reference_dump_bit_exactly():
    ref = DOC["reference_energy_consts"]
    assert list(default.default_energy_consts) == list(ref)
    assert len(ref) == 119
    for key, slots in ref.items():
        mine = default.default_energy_consts[key]
        assert len(mine) == len(slots) == 12
        for s, r in zip(mine, slots):
            assert [getattr(s, f) for f in FIELDS] == r, key


def test_dictionary_shapes():
    assert len(default.default_param_vals) == 30
    assert len(default.default_fitting_params) == 12
    assert len(default.default_fitting_elems) == 93
    assert len(default.default_K_lines) == 35 and len(default.default_L_lines) == 45
    assert len(default.default_M_lines) == 8
    assert set(default.default_learning_rates) == set(default.default_fitting_params) | {"default_lr"}
    assert default.default_param_vals["ENERGY_SLOPE"] == 0.011979999952018261
    assert default.param_name_map["CAL_SLOPE_[E_LINEAR]"] == "ENERGY_SLOPE"
    assert len(default.default_elem_info) == 100 and default.default_elem_info[25].name == "Fe"


def test_binding_gate_counts():
    # at E0 = 12 keV, 465 of the 630 default lines pass the gate
    e0 = torch.tensor(12.0)
    n_lines = n_open = 0
    for name in default.default_fitting_elems:
        if name not in default.default_energy_consts:
            continue
        for s in default.default_energy_consts[name]:
            if s.ratio == 0 or s.energy <= 0:
                continue
            n_lines += 1
            n_open += bool(s.check_binding_energy(default.default_elem_info, e0))
    assert (n_lines, n_open) == (630, 465)


def test_custom_pileup_and_gate_edges():
    extra = constant.define_pileup_constants(default.default_elem_info, ["Fe_Fe", "Ca_K_Fe"], return_dict=True)
    fe = next(e for e in default.default_elem_info if e.name == "Fe")
    assert extra["Fe_Fe"][0].energy == fe.xrf["ka1"] * 2 and extra["Fe_Fe"][3].energy == 0.0
    assert extra["Fe_Fe"][0].is_pileup and extra["Fe_Fe"][0].Si_mu > 0
    assert extra["Fe_Fe"][0].check_binding_energy(default.default_elem_info, 12.0)
    assert not extra["Fe_Fe"][0].check_binding_energy(default.default_elem_info, 7.0)
    pb = default.default_energy_consts["Pb_L"]
    assert not any(s.check_binding_energy(default.default_elem_info, 12.0) for s in pb if s.ratio)
