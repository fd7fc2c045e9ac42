#!/usr/bin/env python
"""Derive the layout of ``maps_fit_parameters_override.txt`` from the reference writer.

Run ONCE in the build container (the reference is mounted at /root/reference there):

    python tools/export_override_template.py

The override file is a fixed XRF-Maps text format: ~200 lines of tags, limits and beamline
process-variable names with a few dozen values filled in.  Instead of restating that text by hand,
this script RUNS the reference's ``write_override_params_file`` (io.py:192-459) with a recording
parameter dictionary and both detector materials, and turns what it wrote into a template:

  * every ``params.get(KEY, fallback)`` the writer performs becomes a ``<<KEY|fallback>>`` field,
  * the element list, the timestamp and the detector/escape lines become named fields.

The result is ``mapstorch_b200/data/override_template.txt`` (derived data, like xrf_tables.json);
``mapstorch_b200.io.write_override_params_file`` fills the fields.
"""
import sys
import types
from pathlib import Path

sys.path.insert(0, "/root/reference")
sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # not installed here; the writer never touches it

from mapstorch import io as rio  # noqa: E402

OUT = Path(__file__).resolve().parents[1] / "mapstorch_b200" / "data" / "override_template.txt"


class Recorder(dict):
    """dict whose get() answers with a field token instead of a value"""

    def copy(self):
        return Recorder(self)

    def get(self, key, default=None):
        return f"<<{key}|{default!r}>>"


def run(detector, escape, tmp):
    rio.write_override_params_file(tmp, None, None, detector, escape)
    return Path(tmp).read_text().split("\n")


def main():
    tmp = "/tmp/_override_probe.txt"
    real = rio.default_param_vals
    rio.default_param_vals = Recorder(real)
    try:
        si = run("Si", 0.123456, tmp)
        ge = run("Ge", 0.123456, tmp)
    finally:
        rio.default_param_vals = real
    assert len(si) == len(ge)
    out = []
    for a, b in zip(si, ge):
        tag = a.split(":", 1)[0]
        if tag == "DATE":
            out.append("DATE: <<@DATE>>")
        elif tag == "ELEMENTS_TO_FIT":
            out.append("ELEMENTS_TO_FIT: <<@ELEMENTS>>")
        elif a != b:
            assert tag == b.split(":", 1)[0] and tag in (
                "DETECTOR_MATERIAL", "SI_ESCAPE_FACTOR", "GE_ESCAPE_FACTOR", "SI_ESCAPE_ENABLE", "GE_ESCAPE_ENABLE"), tag
            out.append(f"{tag}: <<@{tag}>>")
        else:
            out.append(a)
    OUT.write_text("\n".join(out))
    n_fields = sum(line.count("<<") for line in out)
    print(f"wrote {OUT} ({len(out)} lines, {n_fields} fields)")


if __name__ == "__main__":
    main()
