#!/usr/bin/env python
"""Export the physics tables the hot path needs into one compact JSON file.

Run ONCE in the build container, where the upstream reference is mounted at
/root/reference (it cannot travel to the GPU box):

    PYTHONPATH=/root/reference python tools/export_xrf_tables.py

It imports the reference's ``mapstorch.constant`` (constant.py:1223-1374 parses
``xrf_library.csv``; constant.py:768-933 reads ``henke.xdr``) and writes
``mapstorch_b200/data/xrf_tables.json`` holding

* per-Z element records (line energies, absolute line yields, shell yields,
  density, mass, binding energies, jump factors) -- the ``ElementInfo`` fields;
* the Henke f2 grids of the two detector materials (Si, Ge), which is all that
  ``calculate_mu_fractions`` (constant.py:1443-1524) ever looks at;
* the reference's own ``default_energy_consts`` dump (119 keys x 12 slots), kept
  ONLY as a cross-check: ``mapstorch_b200.constant.define_constants`` rebuilds
  the same tables from the first two items and the CPU tests compare.

Nothing here is source code of the reference; it is derived numeric data.
"""
import json
import sys
import warnings
from pathlib import Path

warnings.filterwarnings("ignore")
import numpy as np  # noqa: E402

sys.path.insert(0, "/root/reference")
from mapstorch import constant as rc  # noqa: E402
from mapstorch.default import default_energy_consts, default_elem_info  # noqa: E402

OUT = Path(__file__).resolve().parents[1] / "mapstorch_b200" / "data" / "xrf_tables.json"

LINES = ["ka1", "ka2", "kb1", "kb2", "la1", "la2", "lb1", "lb2", "lb3", "lb4",
         "lg1", "lg2", "lg3", "lg4", "ll", "ln", "ma1", "ma2", "mb", "mg"]


def main():
    elements = []
    for info in default_elem_info:
        elements.append({
            "z": int(info.z),
            "name": info.name,
            "xrf": [float(info.xrf.get(k, 0.0)) for k in LINES],
            "xrf_abs_yield": [float(info.xrf_abs_yield.get(k, 0.0)) for k in LINES],
            "yieldD": {k: float(v) for k, v in info.yieldD.items()},
            "density": float(info.density),
            "mass": float(info.mass),
            "bindingE": {k: float(v) for k, v in info.bindingE.items()},
            "jump": {k: float(v) for k, v in info.jump.items()},
        })
    henke = {}
    for name in ("Si", "Ge"):
        z_array, _ = rc.henkedata.compound(name)
        z = int(np.where(z_array == 1.0)[0][0]) + 1
        energies, _f1, f2, *_ = rc.henkedata.extra(ielement=z - 1)
        henke[name] = {"z": z, "energies_eV": [float(x) for x in energies],
                       "f2": [float(x) for x in f2]}
    check = {}
    for key, slots in default_energy_consts.items():
        check[key] = [[s.name, float(s.energy), float(s.ratio), float(s.Ge_mu), float(s.Si_mu),
                       float(s.mu_fraction), float(s.width_multi), int(s.type), s.ptype,
                       bool(s.is_pileup)] for s in slots]
    doc = {
        "about": "derived from xrf_library.csv v1.2 and henke.xdr through the reference loaders; "
                 "see tools/export_xrf_tables.py",
        "line_keys": LINES,
        "k_names": list(rc.kele), "l_names": list(rc.lele), "m_names": list(rc.mele),
        "elements": elements,
        "henke_f2": henke,
        "reference_energy_consts": check,
    }
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(doc, separators=(",", ":")))
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(elements), "elements,", len(check), "line sets")


if __name__ == "__main__":
    main()
