"""Peak-window helpers of the reference's ``mapstorch/util.py`` (host side, float64, bit-exact ints).

``get_channel_from_energy`` (util.py:60-64), ``get_peak_centers`` (util.py:67-84) and
``get_peak_ranges`` (util.py:87-108) build the channel windows that callers turn into the
``indices`` masks of ``fit_spec`` / ``fit_map``.  ``estimate_and_update_params`` (util.py:189-229)
seeds ENERGY_SLOPE / COMPTON_ANGLE from the scatter peaks of an integrated spectrum.  The
periodic-table widget (UI) is out of scope.
"""
import math

import numpy as np

from mapstorch_b200.constant import M_PI, ENERGY_RES_OFFSET, ENERGY_RES_SQRT
from mapstorch_b200.default import default_fitting_elems, default_energy_consts, default_param_vals


def get_channel_from_energy(energy, ev):
    """Index of the first channel whose energy is >= `energy` (last channel if none)."""
    hits = np.nonzero(np.asarray(ev) >= energy)[0]
    return int(hits[0]) if hits.size else len(ev) - 1


def _compton_energy(e0, angle_deg):
    """Energy of the Compton peak for incident energy e0 (keV) scattered by angle_deg (same float64 expression as
    util.py:74-78, so the window edges derived from it are bit-identical)."""
    return e0 / (1 + (e0 / 511) * (1 - math.cos(angle_deg * 2 * M_PI / 360.0)))


def get_peak_centers(elements, coherent_sct_energy, compton_angle, e_consts=default_energy_consts):
    """{peak name: centre energy}: the two scatter peaks under their component names, element lines as "<El>_<slot>"
    for every slot with a positive energy, in the order of `elements` (util.py:67-84)."""
    centers = {}
    for name in elements:
        if name == "COHERENT_SCT_AMPLITUDE":
            centers[name] = coherent_sct_energy
        elif name == "COMPTON_AMPLITUDE":
            centers[name] = _compton_energy(coherent_sct_energy, compton_angle)
        elif name in default_fitting_elems:
            centers.update({f"{name}_{slot}": line.energy for slot, line in enumerate(e_consts[name]) if line.energy > 0})
    return centers


def get_peak_ranges(elements, coherent_sct_energy, compton_angle, energy_offset, energy_slope, energy_quadratic,
                    energy_range):
    ch = np.arange(energy_range[0], energy_range[1] + 1, dtype=float)
    ev = energy_offset + energy_slope * ch + energy_quadratic * (ch ** 2)
    ranges = {}
    for name, c in get_peak_centers(elements, coherent_sct_energy, compton_angle).items():
        width = math.sqrt(ENERGY_RES_OFFSET ** 2 + (c * ENERGY_RES_SQRT) ** 2) / 2000
        ranges[name] = (get_channel_from_energy(c - width, ev), get_channel_from_energy(c + width, ev))
    return ranges


def peak_range_indices(ranges, n_ch):
    """Sorted unique channel indices covered by the windows of get_peak_ranges (the `indices`
    argument of fit_spec / fit_map, built like apps/optimize_parameters.py:566-603)."""
    mask = np.zeros(n_ch, dtype=bool)
    for lo, hi in ranges.values():
        mask[lo: hi + 1] = True
    return np.nonzero(mask)[0]


def _smooth(values, window):
    arr = np.asarray(values, dtype=float)
    if arr.size <= 1 or window <= 1:
        return arr
    window = int(window) | 1
    pad = window // 2
    return np.convolve(np.pad(arr, (pad, pad), mode="edge"), np.ones(window) / window, mode="valid")[: arr.size]


def _scatter_peak_indices(values, window=3, min_distance=10, top_k=20, threshold_percentile=60, start_frac=0.4):
    """Heuristic (compton, elastic) peak indices: local maxima of the smoothed upper part of the
    spectrum above a percentile, non-maximum suppressed (util.py:111-186)."""
    arr = np.asarray(values, dtype=float)
    if arr.size < 3:
        return None, None
    sm = _smooth(arr, window)
    start = min(max(0, int(start_frac * sm.size)), sm.size - 1)
    threshold = float(np.percentile(sm[start:], threshold_percentile))
    first = max(1, min(start, sm.size - 2))
    cands = [i for i in range(first, sm.size - 1) if sm[i] >= threshold and sm[i] >= sm[i - 1] and sm[i] >= sm[i + 1]]
    cands.sort(key=lambda i: sm[i], reverse=True)
    chosen = []
    for i in cands:
        if all(abs(i - j) >= min_distance for j in chosen):
            chosen.append(i)
        if len(chosen) >= top_k:
            break
    if not chosen:
        return None, None
    elastic = max(chosen)
    left = [i for i in chosen if i < elastic]
    return (max(left, key=lambda i: sm[i]) if left else None), elastic


def _scatter_angle(e_compton, e_elastic):
    """Scattering angle (degrees) that puts the Compton peak of e_elastic at e_compton; None outside acos's domain."""
    try:
        return math.acos(1 - 511 * (1 / e_compton - 1 / e_elastic)) * 180 / math.pi
    except (ValueError, ZeroDivisionError):
        return None


def estimate_and_update_params(int_spec, energy_range, coherent_sct_energy, param_reference=None):
    """Updates for COHERENT_SCT_ENERGY / ENERGY_SLOPE / COMPTON_ANGLE inferred from the scatter peaks (util.py:189-229):
    the elastic peak fixes the gain (its channel carries the incident energy), the Compton peak's channel then gives the
    scattering angle.  Peaks that cannot be found fall back to fixed fractions of the spectrum length."""
    spectrum = np.asarray(int_spec, dtype=float).ravel()
    if coherent_sct_energy is None or coherent_sct_energy <= 0 or spectrum.size == 0:
        return {}
    first = max(0, min(int(energy_range[0]), spectrum.size - 1))
    last = min(spectrum.size, max(first + 1, int(energy_range[1]) + 1))
    if last - first < 3:
        return {}
    found = _scatter_peak_indices(spectrum[first:last])  # (compton, elastic), relative to the window
    fallback = (max(1, int((spectrum.size - 1) // 2)), max(1, int((spectrum.size - 1) // 1.9)))
    compton_ch, elastic_ch = (first + rel if rel is not None else default for rel, default in zip(found, fallback))
    slope = coherent_sct_energy / elastic_ch
    angle = _scatter_angle(slope * compton_ch, coherent_sct_energy)
    return {"COHERENT_SCT_ENERGY": coherent_sct_energy, "ENERGY_SLOPE": slope,
            "COMPTON_ANGLE": default_param_vals["COMPTON_ANGLE"] if angle is None else angle}
