"""MAPS override-file round trip against files, parser output and constant tables
produced by the reference itself (tests/golden/make_golden.py::gen_override).

Text and integers must be identical; the float tables must be identical too, because the edits
are a handful of IEEE multiplications applied in the reference's order (constant.py:1815-1987)."""
import json
import re

import numpy as np
import pytest

from mapstorch_b200 import constant, io as mio
from mapstorch_b200.default import default_elem_info, default_energy_consts


def _write(tmp_path, golden, tag):
    path = tmp_path / f"{tag}.txt"
    path.write_text(str(golden[f"{tag}_text"]))
    return str(path)


@pytest.mark.parametrize("tag", ["A", "B", "D"])
def test_writer_emits_the_reference_file(tmp_path, override_golden, tag):
    kw = json.loads(str(override_golden[f"{tag}_kwargs"]))
    out = tmp_path / "maps_fit_parameters_override.txt"
    assert mio.write_override_params_file(str(out), **kw) == str(out)
    text = out.read_text()
    assert re.search(r"^DATE: \d{4}-\d\d-\d\d \d\d:\d\d:\d\d\.\d{6}$", text, re.M)
    assert re.sub(r"DATE: .*", "DATE: <date>", text) == str(override_golden[f"{tag}_text"])


def test_writer_rejects_what_the_reference_rejects(tmp_path):
    with pytest.raises(ValueError, match="Detector material"):
        mio.write_override_params_file(str(tmp_path / "x.txt"), detector_material="CdTe")
    with pytest.raises(ValueError, match="Unsupported elements"):
        mio.write_override_params_file(str(tmp_path / "x.txt"), elements=["Fe", "Xx"])


@pytest.mark.parametrize("tag", ["A", "B", "C", "D"])
def test_parser_returns_the_reference_dictionary(tmp_path, override_golden, tag):
    want = json.loads(str(override_golden[f"{tag}_parsed"]))
    got = mio.parse_override_params_file(_write(tmp_path, override_golden, tag))
    assert list(got.keys()) == list(want.keys())
    assert got == want


def test_parser_round_trip_and_missing_file(tmp_path):
    path = tmp_path / "o.txt"
    vals = {"ENERGY_OFFSET": -0.0123, "ENERGY_SLOPE": 0.0101, "COMPTON_F_STEP": 0.5, "SNIP_WIDTH": 0.8}
    mio.write_override_params_file(str(path), vals, ["Fe", "Zn"], "Ge", 0.125)
    back = mio.parse_override_params_file(str(path))
    for k, v in vals.items():
        assert back[k] == v
    assert back["elements_to_fit"] == ["Fe", "Zn"] and back["detector_material"] == "Ge"
    assert back["escape_factor"] == 0.125
    with pytest.warns(UserWarning, match="Error parsing override parameter file"):
        empty = mio.parse_override_params_file(str(tmp_path / "nope.txt"))
    assert empty == {"elements_to_fit": [], "elements_with_pileup": [], "detector_material": "Si",
                     "escape_factor": 0.0}


def _tables(consts):
    names = list(consts.keys())
    num = np.array([[[s.energy, s.ratio, s.Ge_mu, s.Si_mu, s.mu_fraction, s.width_multi] for s in consts[k]]
                    for k in names], dtype=np.float64)
    kinds = np.array([[s.type for s in consts[k]] for k in names], dtype=np.int64)
    ptypes = np.array([[s.ptype for s in consts[k]] for k in names])
    pile = np.array([[s.is_pileup for s in consts[k]] for k in names])
    slot_names = np.array([[s.name for s in consts[k]] for k in names])
    return names, num, kinds, ptypes, pile, slot_names


@pytest.mark.parametrize("tag", ["A", "B", "C", "D"])
def test_read_constants_tables_are_identical(tmp_path, override_golden, tag):
    g = override_golden
    consts = constant.read_constants(_write(tmp_path, g, tag), default_elem_info)
    names, num, kinds, ptypes, pile, slot_names = _tables(consts)
    assert names == [str(n) for n in g[f"{tag}_consts_names"]]
    assert np.array_equal(num, g[f"{tag}_consts_num"])
    assert np.array_equal(kinds, g[f"{tag}_consts_type"])
    assert np.array_equal(ptypes, g[f"{tag}_consts_ptype"])
    assert np.array_equal(pile, g[f"{tag}_consts_is_pileup"])
    assert np.array_equal(slot_names, g[f"{tag}_consts_slot_names"])


def test_read_constants_edits_what_the_file_says(tmp_path, override_golden):
    base = default_energy_consts
    si = constant.read_constants(_write(tmp_path, override_golden, "A"), default_elem_info)
    ge = constant.read_constants(_write(tmp_path, override_golden, "B"), default_elem_info)
    assert si["Fe"][2].ratio == base["Fe"][2].ratio * 1.2 and si["Fe"][0].ratio == 1.0
    assert si["Sn_L"][0].ratio == 1.0 and si["Sn_L"][2].ratio == 0.0 and si["Pt_L"][4].ratio == 0.0  # L1/L2 factors 0
    assert si["Cu"][2].ratio == base["Cu"][2].ratio
    assert all(s.mu_fraction == s.Si_mu for s in si["Fe"]) and all(s.mu_fraction == s.Ge_mu for s in ge["Fe"])
    assert ge["Fe"][0].Ge_mu != ge["Fe"][0].Si_mu
    # the default dictionaries are not touched by reading a file
    assert default_energy_consts["Fe"][2].ratio == base["Fe"][2].ratio


def test_read_constants_rows_form_and_tail_tags(tmp_path, override_golden):
    g = override_golden
    path = _write(tmp_path, g, "A")
    rows, names = constant.read_constants(path, default_elem_info, pileups=["Si_Cl"], return_dict=False)
    assert [str(n) for n in names] == [str(n) for n in g["A_rows_names"]]
    assert np.array_equal(np.array([[s.ratio for s in r] for r in rows]), g["A_rows_ratio"])
    for tag in ("TAIL_FRACTION_ADJUST_SI", "TAIL_WIDTH_ADJUST_SI"):
        assert str(g["raises_" + tag]) == "AttributeError"
        bad = tmp_path / "tail.txt"
        bad.write_text(str(g["A_text"]) + f"\n{tag}: 1.5\n")
        with pytest.raises(AttributeError):
            constant.read_constants(str(bad), default_elem_info)
