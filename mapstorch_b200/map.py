"""MAPS forward model -- same functions as the reference's ``mapstorch/map.py``, run by kernel K1.

``peak/step/tail`` (map.py:57-69), ``elastic_peak`` (72-79), ``compton_peak`` (82-124),
``escape_peak`` (127-150), ``model_elem_spec`` (152-257) and ``model_spec`` (260-309) keep their
names, argument order and meaning.  The scalar parameter arithmetic happens once on the host
(:mod:`mapstorch_b200.tables`); the per-channel work is one launch of ``mt_model_terms``
(csrc/mt_model.cu) instead of ~10 ATen calls per emission line.  Results live on the CUDA
device and are copied to ``device`` if the caller asks for another one; there is no CPU
compute path.
"""
import ctypes

import numpy as np
import torch

from mapstorch_b200 import _native, tables
from mapstorch_b200.constant import M_PI, SQRT_2XPI, M_SQRT2  # noqa: F401  (re-exported like the reference)
from mapstorch_b200.default import default_fitting_elems, default_energy_consts, default_elem_info

f32 = np.float32


def _cuda_f32(t):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.float32))
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


def _to_device(t, device):
    return t if device is None or torch.device(device).type == "cuda" else t.to(device)


_EV_CACHE = {}


def _energy_axis(params, energy_range):
    """Energy axis on the device (mt_energy_axis: the reference's fp32 operation sequence, bit-identical to
    tables.energy_axis), kept for the last few calibrations: the trial points of fit_spec(tune_params=True) move one
    parameter at a time, most of them not the axis.  No host copy, no host-to-device transfer, no synchronisation."""
    cal = tuple(tables.scalar32(params[k]) for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC"))
    key = (int(energy_range[0]), int(energy_range[1]), torch.cuda.current_device()) + tuple(float(v) for v in cal)
    ev_dev = _EV_CACHE.get(key)
    if ev_dev is None:
        if len(_EV_CACHE) >= 16:
            _EV_CACHE.clear()
        n_ch = key[1] - key[0] + 1
        ev_dev = _EV_CACHE[key] = torch.empty(n_ch, dtype=torch.float32, device="cuda")
        _native.call("mt_energy_axis", key[0], n_ch, *cal, ev_dev.data_ptr(), _native.stream_ptr())
    return ev_dev


def run_terms(ev, columns, out=None):
    """Evaluate term columns on the GPU: returns a CUDA tensor [n_cols, len(ev)]."""
    terms, starts = tables.pack_columns(columns)
    return run_packed(ev, terms, starts, out)


def run_packed(ev, terms, starts, out=None):
    """run_terms on an already packed table (tables.pack_columns / tables.compile_packed)."""
    _native.require_device()
    ev_dev = ev if isinstance(ev, torch.Tensor) and ev.is_cuda and ev.dtype == torch.float32 and ev.dim() == 1 \
        and ev.is_contiguous() else _cuda_f32(ev).reshape(-1)
    n_cols, n_ch = len(starts) - 1, ev_dev.numel()
    if out is None:
        out = torch.empty((n_cols, n_ch), dtype=torch.float32, device="cuda")
    _native.call("mt_model_terms", ev_dev.data_ptr(), n_ch, terms.ctypes.data_as(ctypes.c_void_p),
                 starts.ctypes.data_as(ctypes.c_void_p), n_cols, out.data_ptr(), out.stride(0),
                 _native.stream_ptr())
    return out


def _single(ev, term, device=None):
    shape = ev.shape if isinstance(ev, torch.Tensor) else np.shape(ev)
    return _to_device(run_terms(ev, [[term]])[0].reshape(shape), device)


def _dev_of(t):
    return t.device if isinstance(t, torch.Tensor) else None


def peak(gain, sigma, delta_energy):
    """Gaussian line shape, map.py:57-58."""
    g, s = tables.scalar32(gain), tables.scalar32(sigma)
    return _single(delta_energy, (_native.MT_TERM_GAUSS, 1.0, 0.0, s, g / (s * f32(SQRT_2XPI)), 0.0, 0.0, 1.0),
                   _dev_of(delta_energy))


def step(gain, sigma, delta_energy, peak_E):
    """Step (incomplete charge collection) shape, map.py:61-62."""
    g, s, pe = tables.scalar32(gain), tables.scalar32(sigma), tables.scalar32(peak_E)
    return _single(delta_energy, (_native.MT_TERM_STEP, 1.0, 0.0, f32(M_SQRT2) * s, g / f32(2.0) / pe, 0.0, 0.0, 1.0),
                   _dev_of(delta_energy))


def tail(gain, sigma, delta_energy, gamma):
    """Exponential tail shape, map.py:65-69."""
    term = tables._tail(f32(0.0), tables.scalar32(sigma), tables.scalar32(gain), tables.scalar32(gamma), f32(1.0))
    return _single(delta_energy, term, _dev_of(delta_energy))


def elastic_peak(params, ev, gain):
    """Coherent scatter peak, map.py:72-79."""
    amp = tables.pow10(params["COHERENT_SCT_AMPLITUDE"])
    return _to_device(run_terms(ev, [tables.elastic_terms(params, tables.scalar32(gain), amp)])[0], _dev_of(ev))


def compton_peak(params, ev, gain, use_step=True, use_tail=False):
    """Compton scatter peak, map.py:82-124."""
    amp = tables.pow10(params["COMPTON_AMPLITUDE"])
    cols = [tables.compton_terms(params, tables.scalar32(gain), amp, use_step, use_tail)]
    return _to_device(run_terms(ev, cols)[0], _dev_of(ev))


def escape_peak(spectra, ev, escape_factor, device, detector_type="Si"):
    """Detector escape peak of a spectrum: shifted, scaled copy (map.py:127-150)."""
    _native.require_device()
    ev_cpu = ev.detach().to("cpu", torch.float32) if isinstance(ev, torch.Tensor) else torch.as_tensor(ev, dtype=torch.float32)
    spec = _cuda_f32(spectra)
    out = torch.zeros_like(spec)
    bins = tables.escape_bins(ev_cpu, detector_type)
    if bins > 0:
        out[: spec.numel() - bins] = spec[bins:] * float(escape_factor)
    return _to_device(out, device)


def model_elem_spec(params, element_to_fit, ev, use_step=True, use_tail=True, device="cpu",
                    e_consts=default_energy_consts, elem_info=default_elem_info):
    """Spectrum of one element (all its gated lines), map.py:152-257."""
    amp = tables.pow10(params[element_to_fit])
    cols = [tables.element_terms(params, e_consts[element_to_fit], elem_info, amp, use_step, use_tail)]
    return _to_device(run_terms(ev, cols)[0], device)


def model_spec(params, energy_range, elements_to_fit=default_fitting_elems, device="cpu",
               e_consts=default_energy_consts, elem_info=default_elem_info, detector_type="Si",
               escape_factor=0.0):
    """Full model spectrum over energy_range (inclusive channel pair), map.py:260-309."""
    _native.require_device()
    ev_dev = _energy_axis(params, energy_range)
    names = tables.fitted_components(elements_to_fit, e_consts)
    per_col = run_packed(ev_dev, *tables.compile_packed(params, names, e_consts, elem_info, True, False, log_amps=params))
    out = torch.empty(ev_dev.numel(), dtype=torch.float32, device="cuda")
    bins = tables.escape_bins(tables.energy_axis(params, energy_range), detector_type) if escape_factor > 0.0 else 0
    _native.call("mt_model_sum", per_col.data_ptr(), per_col.stride(0), per_col.shape[0], ev_dev.numel(), bins,
                 float(escape_factor), out.data_ptr(), _native.stream_ptr())
    return _to_device(out, device)


def element_basis(params, energy_range, elements_to_fit, e_consts=default_energy_consts,
                  elem_info=default_elem_info, detector_type="Si", escape_factor=0.0, use_step=True,
                  use_tail=False):
    """Unit-amplitude model columns B[E, C] (CUDA, fp32) and their names, in model_spec order.

    The model is linear in the amplitudes 10**a, so model_spec == sum_e 10**a_e * B[e];
    with an escape factor each column carries its own shifted copy (the shift-add is linear too)."""
    _native.require_device()
    ev_dev = _energy_axis(params, energy_range)
    names = tables.fitted_components(elements_to_fit, e_consts)
    basis = run_packed(ev_dev, *tables.compile_packed(params, names, e_consts, elem_info, use_step, use_tail, log_amps=None))
    if escape_factor > 0.0:
        bins = tables.escape_bins(tables.energy_axis(params, energy_range), detector_type)
        if bins > 0:
            shifted = torch.zeros_like(basis)
            shifted[:, : basis.shape[1] - bins] = basis[:, bins:] * float(escape_factor)
            basis = basis + shifted
    return basis, names
