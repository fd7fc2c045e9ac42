"""Physics tables behind the fitting dictionaries (host side, float64, one-off).

Mirror of the reference interface ``mapstorch/constant.py`` for the pieces the
hot path consumes -- ``MapsElements.get_element_info`` (constant.py:1223-1374),
``EnergyStruct`` + ``check_binding_energy`` (constant.py:1377-1440),
``calculate_mu_fractions`` (constant.py:1443-1524), ``define_pileup_constants``
(constant.py:1527-1621) and ``define_constants`` (constant.py:1624-1812) --
re-built table-driven on top of ``data/xrf_tables.json`` (derived numeric data,
see ``tools/export_xrf_tables.py``).  The outputs feed the line-table compiler
in :mod:`mapstorch_b200.tables`; nothing here runs per pixel.
"""
import json
import math
from pathlib import Path

import numpy as np

AVOGADRO = 6.02204531e23
HC_ANGSTROMS = 12398.52
RE = 2.817938070e-13  # classical electron radius, cm
M_PI = math.pi
M_SQRT2 = math.sqrt(2)
SQRT_2XPI = math.sqrt(2 * math.pi)
ENERGY_RES_OFFSET = 150.0
ENERGY_RES_SQRT = 12.0

_DATA_PATH = Path(__file__).resolve().parent / "data" / "xrf_tables.json"
_DOC = None


def _doc():
    global _DOC
    if _DOC is None:
        with open(_DATA_PATH, "r") as fh:
            _DOC = json.load(fh)
    return _DOC


def _names(which):
    return list(_doc()[which])


# element-name tables in the order the reference lays out its parameter rows
kele = _names("k_names")
lele = _names("l_names")
mele = _names("m_names")
kele_pos = np.arange(len(kele)) + 1
lele_pos = np.arange(len(lele)) + np.amax(kele_pos) + 1
mele_pos = np.arange(len(mele)) + np.amax(lele_pos) + 1

# (slot order, line key, EnergyStruct.type, ptype) per line family
_K_SLOTS = (("ka1", 1, "Ka1"), ("ka2", 1, "Ka2"), ("kb1", 2, "Kb1"), ("kb2", 2, "Kb2"))
_L_SLOTS = tuple((k, 3, k[0].upper() + k[1:]) for k in
                 ("la1", "la2", "lb1", "lb2", "lb3", "lb4", "lg1", "lg2", "lg3", "lg4", "ll", "ln"))
_M_SLOTS = (("ma1", 7, "Ma1"), ("ma2", 7, "Ma2"), ("mb", 7, "Mb"), ("mg", 7, "Mg"))
N_SLOTS = 12

_L1_GATED = ("Lb3", "Lb4", "Lg2", "Lg3", "Lg4")
_L2_GATED = ("Lb1", "Lg1", "Ln")
_L3_GATED = ("La1", "La2", "Lb2", "Ll")


class ElementInfo:
    """Per-Z record; same attribute names as the reference (constant.py:1116-1217)."""

    def __init__(self):
        self.z = 0
        self.name = ""
        self.xrf = {}
        self.xrf_abs_yield = {}
        self.yieldD = {"k": 0.0, "l1": 0.0, "l2": 0.0, "l3": 0.0, "m": 0.0}
        self.density = 1.0
        self.mass = 1.0
        self.bindingE = {}
        self.jump = {}

    def __repr__(self):
        return f"Element {self.name} with Z={self.z} and mass={self.mass} g/mol"


class MapsElements:
    def get_element_info(self, filepath=None):
        """List of 100 ``ElementInfo`` (index = Z-1).  ``filepath`` may point at another
        JSON export with the same schema; the CSV itself is not shipped."""
        doc = _doc() if filepath is None else json.loads(Path(filepath).read_text())
        keys = doc["line_keys"]
        out = []
        for rec in doc["elements"]:
            info = ElementInfo()
            info.z, info.name = rec["z"], rec["name"]
            info.xrf = dict(zip(keys, rec["xrf"]))
            info.xrf_abs_yield = dict(zip(keys, rec["xrf_abs_yield"]))
            info.yieldD = dict(rec["yieldD"])
            info.density, info.mass = rec["density"], rec["mass"]
            info.bindingE = dict(rec["bindingE"])
            info.jump = dict(rec["jump"])
            out.append(info)
        return out


class EnergyStruct:
    """One emission-line slot of an element (constant.py:1377-1388)."""

    __slots__ = ("name", "energy", "ratio", "Ge_mu", "Si_mu", "mu_fraction", "width_multi",
                 "type", "ptype", "is_pileup")

    def __init__(self):
        self.name = ""
        self.energy = 0.0
        self.ratio = 0.0
        self.Ge_mu = 0.0
        self.Si_mu = 0.0
        self.mu_fraction = 0.0
        self.width_multi = 1.0
        self.type = 0
        self.ptype = ""
        self.is_pileup = False

    def binding_edge(self, info_elements):
        """Absorption-edge energies (keV) that must lie below the incident energy for this
        line to be excited; ``()`` means "always on", ``None`` means "never".
        Same decision tree as constant.py:1390-1437, returned as data so that the
        line-table compiler can apply the gate for any incident energy."""
        if self.type == 0:
            return None
        by_name = {e.name: e for e in info_elements}
        if self.is_pileup:
            parts = self.name.split("_")
            if len(parts) != 2 or parts[0] not in by_name or parts[1] not in by_name:
                return ()
            k1 = by_name[parts[0]].bindingE.get("K", 0.0)
            k2 = by_name[parts[1]].bindingE.get("K", 0.0)
            if k1 <= 0.0 or k2 <= 0.0:
                return ()
            return (k1, k2)
        info = by_name.get(self.name.split("_")[0])
        if info is None:
            return ()
        if self.ptype.startswith("K"):
            return (info.bindingE["K"],)
        if self.ptype.startswith("L"):
            if self.ptype in _L1_GATED:
                return (info.bindingE["L1"],)
            if self.ptype in _L2_GATED:
                return (info.bindingE["L2"],)
            if self.ptype in _L3_GATED:
                return (info.bindingE["L3"],)
            print(f"Error: Unknown L shell type {self.ptype}")
            return None
        if self.ptype.startswith("M"):
            return (info.bindingE["M1"],)
        return None

    def check_binding_energy(self, info_elements, coherent_sct_energy):
        edges = self.binding_edge(info_elements)
        if edges is None:
            return False
        return all(bool(edge < coherent_sct_energy) for edge in edges)

    def __repr__(self):
        return (f"{self.name} {self.ptype} {self.energy} {self.ratio} "
                f"{self.mu_fraction} {self.width_multi}")


def _blank_rows(n):
    rows = np.empty((n, N_SLOTS), dtype=object)
    for i in range(n):
        for j in range(N_SLOTS):
            rows[i, j] = EnergyStruct()
    return rows


def _detector_mu(energies_kev, detector, info_elements):
    """Linear attenuation-style mu of the detector material at each line energy from
    log-log interpolated Henke f2 (arithmetic order of constant.py:1484-1524)."""
    table = _doc()["henke_f2"][detector]
    density = {"Ge": 5.323, "Si": 2.33}[detector]
    grid = np.asarray(table["energies_eV"], dtype=np.float64)
    f2_grid = np.asarray(table["f2"], dtype=np.float64)
    mass = next((e.mass for e in info_elements if e.name == detector), None)
    if mass is None:
        raise ValueError(f"Could not find detector element {detector} in element info")
    molecules_per_cc = density * AVOGADRO / mass
    out = []
    for e_kev in energies_kev:
        energy = e_kev * 1000.0
        if energy == 0.0:
            out.append(0.0)
            continue
        wavelength = HC_ANGSTROMS / energy
        constant = RE * (1.0e-16 * wavelength * wavelength) * molecules_per_cc / (2.0 * np.pi)
        above = np.nonzero(grid > energy)[0]
        below = np.nonzero(grid < energy)[0]
        hi = above[0] if above.size else 0
        lo = below[-1] if below.size else len(grid) - 1
        ln_lo, ln_hi = np.log(grid[lo]), np.log(grid[hi])
        frac = (np.log(energy) - ln_lo) / (ln_hi - ln_lo)
        lf_lo, lf_hi = np.log(np.abs(f2_grid[lo])), np.log(np.abs(f2_grid[hi]))
        f2 = np.exp(lf_lo + frac * (lf_hi - lf_lo))
        beta = constant * f2
        out.append((energy * 4.0 * np.pi * beta) / (density * 1.239852) * 10000.0)
    return out


def calculate_mu_fractions(pars, info_elements, calc_range):
    rows = range(calc_range) if isinstance(calc_range, (int, np.integer)) else range(*calc_range)
    for i in rows:
        energies = [pars[i, j].energy for j in range(N_SLOTS)]
        ge = _detector_mu(energies, "Ge", info_elements)
        si = _detector_mu(energies, "Si", info_elements)
        for j in range(N_SLOTS):
            if energies[j] * 1000.0 == 0.0:
                continue
            pars[i, j].Ge_mu = ge[j]
            pars[i, j].Si_mu = si[j]
            # silicon detector by default; a Ge detector is an override-file switch upstream
            pars[i, j].mu_fraction = si[j]


def _split_pileup(name):
    parts = name.split("_")
    shells = ("L", "M")
    if len(parts) == 2:
        return parts[0], parts[1]
    if len(parts) == 3:
        if parts[1] in shells:
            return parts[0], parts[2]
        if parts[2] in shells:
            return parts[0], parts[1]
    if len(parts) == 4 and parts[1] in shells and parts[3] in shells:
        return parts[0], parts[2]
    return None


def define_pileup_constants(info_elements, pileups, return_dict=False):
    by_name = {}
    for e in info_elements:
        by_name.setdefault(e.name, e)
    rows = _blank_rows(len(pileups))
    names = np.array([" " * 32] * len(pileups))
    for i, pileup in enumerate(pileups):
        names[i] = pileup
        pair = _split_pileup(pileup)
        if pair is None:
            print(f"Pileup {pileup} skipped. Currently only A_B pileups are supported.")
            continue
        a, b = by_name.get(pair[0]), by_name.get(pair[1])
        if not pair[0] or not pair[1] or a is None or b is None:
            continue
        for slot in rows[i]:
            slot.name = pileup
        kb_over_ka_a = a.xrf_abs_yield["kb1"] / a.xrf_abs_yield["ka1"]
        kb_over_ka_b = b.xrf_abs_yield["kb1"] / b.xrf_abs_yield["ka1"]
        combos = [(a.xrf["ka1"] + b.xrf["ka1"], 1.0, 1),
                  (a.xrf["ka1"] + b.xrf["kb1"], kb_over_ka_b, 2),
                  (a.xrf["kb1"] + b.xrf["kb1"], kb_over_ka_a * kb_over_ka_b, 2)]
        if pair[0] != pair[1]:
            combos.append((a.xrf["kb1"] + b.xrf["ka1"], kb_over_ka_a, 2))
        for slot, (energy, ratio, kind) in zip(rows[i], combos):
            slot.energy, slot.ratio, slot.type, slot.is_pileup = energy, ratio, kind, True
    calculate_mu_fractions(rows, info_elements, len(pileups))
    if return_dict:
        return {rows[i, 0].name: rows[i] for i in range(rows.shape[0])}
    return rows, names


def _fill_family(row, label, info, slots, m_family=False):
    for slot in row:
        slot.name = label
    ref_yield = info.xrf_abs_yield[slots[0][0]]
    for slot, (key, kind, ptype) in zip(row, slots):
        slot.energy = info.xrf[key]
        slot.type, slot.ptype = kind, ptype
        line_yield = info.xrf_abs_yield[key]
        if key == slots[0][0]:
            slot.ratio = 1.0
        elif m_family:
            slot.ratio = line_yield / ref_yield if (ref_yield != 0.0 and line_yield != 0.0) else 0.0
        else:
            slot.ratio = line_yield / ref_yield


def define_constants(info_elements, pileups=None, return_dict=True):
    by_name = {}
    for e in info_elements:
        by_name.setdefault(e.name, e)
    n_rows = int(np.amax(mele_pos) - np.amin(kele_pos) + 1)
    rows = _blank_rows(n_rows)
    names = []
    r = 0
    for family, slots, is_m in ((kele, _K_SLOTS, False), (lele, _L_SLOTS, False), (mele, _M_SLOTS, True)):
        for label in family:
            base = label if slots is _K_SLOTS else label[:-2]
            if base in by_name:
                _fill_family(rows[r], label, by_name[base], slots, m_family=is_m)
            names.append(label)
            r += 1
    calculate_mu_fractions(rows, info_elements, n_rows)
    names = np.array(names)
    if pileups is not None:
        p_rows, p_names = define_pileup_constants(info_elements, pileups)
        rows = np.vstack((rows, p_rows))
        names = np.append(names, p_names)
    if return_dict:
        return {rows[i, 0].name: rows[i] for i in range(rows.shape[0])}
    return rows, names


def _literature_ratios(row, info, slots):
    """reset the first len(slots) ratios of a row to yield / yield(first line)"""
    first = info.xrf_abs_yield[slots[0][0]]
    row[0].ratio = 1.0
    for slot, (key, _, _) in zip(row[1:], slots[1:]):
        slot.ratio = info.xrf_abs_yield[key] / first


# which L sub-shell scales which slot (order of _L_SLOTS): factor index 0 = L1, 1 = L2, 2 = L3
_L_SUBSHELL_OF_SLOT = (2, 2, 1, 2, 0, 0, 1, 0, 0, 0, 2, 1)


def read_constants(filename, info_elements, pileups=None, return_dict=True):
    """``define_constants`` plus the edits a MAPS override file asks for (constant.py:1815-1987).

    Tags are matched un-stripped at the start of a line (indented lines are comments):
    ``ELEMENTS_WITH_PILEUP`` appends pile-up rows, ``DETECTOR_MATERIAL`` selects the Si (1) or Ge (other)
    absorption column as ``mu_fraction``, ``BRANCHING_FAMILY_ADJUSTMENT_L`` (name + 3 factors: L1, L2, L3)
    resets an L row to the literature ratios and scales them per sub-shell,
    ``BRANCHING_RATIO_ADJUSTMENT_L`` (name + 12 factors) and ``BRANCHING_RATIO_ADJUSTMENT_K`` (name + 4)
    multiply the ratios slot by slot.  The two ``TAIL_*_ADJUST_SI`` tags raise AttributeError upstream
    (attribute access on an object-array slice); they do here too rather than silently doing something
    the reference never did."""
    rows, names = define_constants(info_elements, pileups, False)
    el_names = [e.name for e in info_elements]

    def find_row(label):
        hits = np.where(names == label)[0]
        return int(hits[0]) if hits.size else None

    def base_element(label):
        # the reference cuts two characters off whatever name it is given (meant for the "_L" suffix)
        return el_names.index(label[:-2]) if label[:-2] in el_names else -1

    with open(str(filename), "rt") as fh:
        for line in fh:
            if ":" not in line:
                continue
            pieces = line.split(":")
            tag, value = pieces[0], "".join(pieces[1:])
            fields = [x.strip() for x in value.split(",")]
            if tag == "ELEMENTS_WITH_PILEUP":
                p_rows, p_names = define_pileup_constants(info_elements, fields)
                rows = np.vstack((rows, p_rows))
                names = np.append(names, p_names)
            elif tag == "DETECTOR_MATERIAL":
                column = "Si_mu" if int(value) == 1 else "Ge_mu"
                for slot in rows.flat:
                    slot.mu_fraction = getattr(slot, column)
            elif tag == "BRANCHING_FAMILY_ADJUSTMENT_L":
                ii = find_row(fields[0])
                if ii is not None and len(fields) == 4:
                    factors = [float(v) for v in fields[1:4]]
                    j = base_element(fields[0])
                    if j > 0:
                        _literature_ratios(rows[ii], info_elements[j], _L_SLOTS)
                        for slot, shell in zip(rows[ii], _L_SUBSHELL_OF_SLOT):
                            slot.ratio *= factors[shell]
            elif tag == "BRANCHING_RATIO_ADJUSTMENT_L":
                ii = find_row(fields[0])
                if ii is not None and len(fields) >= 13:
                    for jj in range(12):
                        rows[ii, jj].ratio *= float(fields[jj + 1])
            elif tag == "BRANCHING_RATIO_ADJUSTMENT_K":
                ii = find_row(fields[0])
                if ii is not None and len(fields) == 5:
                    j = base_element(fields[0])
                    if j > 0:
                        _literature_ratios(rows[ii], info_elements[j], _K_SLOTS)
                    for jj in range(4):
                        rows[ii, jj].ratio *= float(fields[jj + 1])
            elif tag in ("TAIL_FRACTION_ADJUST_SI", "TAIL_WIDTH_ADJUST_SI"):
                float(value)
                if find_row("Si") is not None:
                    attr = "mu_fraction" if tag == "TAIL_FRACTION_ADJUST_SI" else "width_multi"
                    raise AttributeError(f"'numpy.ndarray' object has no attribute '{attr}'")
    if return_dict:
        return {rows[i, 0].name: rows[i] for i in range(rows.shape[0])}
    return rows, names
