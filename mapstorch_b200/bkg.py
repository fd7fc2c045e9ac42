"""SNIP background -- same entry points as the reference's ``mapstorch/bkg.py``, run by kernel K2.

``snip_bkg`` (bkg.py:69-114) keeps its signature; the whole chain (boxcar, log1p, all clipping
passes, expm1) is ONE launch of ``mt_snip`` (csrc/mt_snip.cu) with the spectra held on chip,
instead of ~10 ATen ops and a host sync per pass.  The per-pass gather indices depend only on
the calibration parameters and are computed once on the host (tables.snip_tables) with the
reference's own fp32 op sequence, so they are bit-identical.
"""
import numpy as np
import torch

from mapstorch_b200 import _native, tables
from mapstorch_b200.constant import M_SQRT2  # noqa: F401


def snip_tables_device(n_ch, er, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime, aux=True):
    """(gather table [n_pass, n_ch] int32 on the device, n_pass).  For windows of up to 4096 channels the table
    also carries its pre-scaled form (mt_snip_scale_tables) as attribute ``_mt_scaled`` and the tables of the
    contiguous-run kernel as ``_mt_runs``; snip_launch picks them up.  aux=False skips both (three launches and a
    stream synchronisation that only pay off over many spectra): the plain table drives the strided kernel."""
    six = (e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime)
    if not aux and n_ch <= 65535 and all(isinstance(v, (torch.Tensor, np.floating)) for v in six):
        # every operand enters the reference's arithmetic as an fp32 value: the table is built on the device with the same
        # rounded operations (mt_snip_tables_f32); the host only needs the pass count, i.e. the largest width
        width = tables.snip_widths(n_ch, er[0], *six)
        n_pass = tables.snip_pass_count(width.max()) if n_ch else 2
        vals = [np.float32(v.detach().cpu().to(torch.float32).item()) if isinstance(v, torch.Tensor) else np.float32(v)
                for v in six]
        lohi = torch.empty((n_pass, n_ch), dtype=torch.int32, device="cuda")
        _native.call("mt_snip_tables_f32", int(n_ch), int(er[0]), *vals, n_pass, lohi.data_ptr(), _native.stream_ptr())
        return lohi, n_pass
    tab = tables.snip_tables(n_ch, er[0], *six)
    lohi = torch.from_numpy(tab.view(np.int32)).to("cuda")
    if not aux:
        return lohi, tab.shape[0]
    pitch = _native.load().mt_snip_scaled_pitch(int(n_ch))
    if pitch > 0 and tab.shape[0] > 0:
        scaled = torch.empty((tab.shape[0] + 1, pitch), dtype=torch.int32, device="cuda")
        _native.call("mt_snip_scale_tables", lohi.data_ptr(), tab.shape[0], int(n_ch), scaled.data_ptr(), pitch,
                     _native.stream_ptr())
        lohi._mt_scaled = scaled
    pitch = _native.load().mt_snip_runs_pitch(int(n_ch))
    if pitch > 0 and tab.shape[0] > 0:
        # tables of the contiguous-run kernel (csrc/mt_snip_runs.cu): swizzled gather words + one descriptor per run
        runs = torch.empty((tab.shape[0], pitch), dtype=torch.int32, device="cuda")
        _native.call("mt_snip_runs_tables", lohi.data_ptr(), tab.shape[0], int(n_ch), runs.data_ptr(), pitch,
                     _native.stream_ptr())
        lohi._mt_runs = runs
    return lohi, tab.shape[0]


def _runs_launch(spec2d, ch0, n_ch, lohi_dev, n_pass, boxcar_size, out, out_mode, overflow, stream):
    """Launch the contiguous-run kernel where it applies (aligned rows and crops, n_ch % 8 == 0, boxcar 5); its results
    are bit-identical to the strided kernel's.  Returns False when the caller has to use the strided kernel."""
    runs = getattr(lohi_dev, "_mt_runs", None)
    if runs is None or runs.shape[0] != int(n_pass):
        return False
    if not _native.load().mt_snip_runs_eligible(spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), spec2d.stride(0),
                                                int(ch0), int(n_ch), int(boxcar_size), out.data_ptr(), out.stride(0),
                                                int(out_mode)):
        return False
    _native.call("mt_snip_runs", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), spec2d.shape[0], spec2d.stride(0),
                 int(ch0), int(n_ch), runs.data_ptr(), runs.stride(0), int(n_pass), int(boxcar_size), out.data_ptr(),
                 out.stride(0), int(out_mode), overflow.data_ptr() if overflow is not None else None,
                 _native.stream_ptr(stream))
    return True


def snip_launch(spec2d, ch0, n_ch, lohi_dev, n_pass, boxcar_size, out, out_mode, stream=None):
    """spec2d: CUDA [P, ld] fp32/fp16 (row-contiguous); out: CUDA fp32 [P, ld_out]."""
    if _runs_launch(spec2d, ch0, n_ch, lohi_dev, n_pass, boxcar_size, out, out_mode, None, stream):
        return out
    scaled = getattr(lohi_dev, "_mt_scaled", None)
    if scaled is not None and scaled.shape[0] == int(n_pass) + 1:
        _native.call("mt_snip_scaled", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), spec2d.shape[0],
                     spec2d.stride(0), int(ch0), int(n_ch), scaled.data_ptr(), scaled.stride(0), int(n_pass),
                     int(boxcar_size), out.data_ptr(), out.stride(0), int(out_mode), _native.stream_ptr(stream))
        return out
    _native.call("mt_snip", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), spec2d.shape[0], spec2d.stride(0),
                 int(ch0), int(n_ch), lohi_dev.data_ptr(), int(n_pass), int(boxcar_size), out.data_ptr(),
                 out.stride(0), int(out_mode), _native.stream_ptr(stream))
    return out


def snip_residual_f16(spec2d, ch0, n_ch, lohi_dev, n_pass, boxcar_size, out_half, overflow, stream=None):
    """The residual y - bkg as two fp16 planes [P, 2*n_ch] (hi | remainder) for the fp16 tensor-core contraction;
    overflow (int32 [1]) is raised when a residual does not fit fp16."""
    if _runs_launch(spec2d, ch0, n_ch, lohi_dev, n_pass, boxcar_size, out_half, _native.MT_SNIP_RESIDUAL_HILO_F16, overflow,
                    stream):
        return out_half
    scaled = getattr(lohi_dev, "_mt_scaled", None)
    use_scaled = scaled is not None and scaled.shape[0] == int(n_pass) + 1
    tab = scaled if use_scaled else lohi_dev
    _native.call("mt_snip_residual_f16", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), spec2d.shape[0],
                 spec2d.stride(0), int(ch0), int(n_ch), tab.data_ptr(), tab.stride(0) if use_scaled else int(n_ch),
                 int(use_scaled), int(n_pass), int(boxcar_size), out_half.data_ptr(), out_half.stride(0),
                 overflow.data_ptr(), _native.stream_ptr(stream))
    return out_half


def snip_op(background, current_width, max_of_xmin, min_of_xmax, device):
    """One clipping pass (bkg.py:53-66) on an already log-compressed background."""
    _native.require_device()
    bg = background.detach().to("cuda", torch.float32)
    n_ch = bg.shape[-1]
    idx = torch.arange(n_ch, dtype=torch.float32)
    w = current_width.detach().cpu().to(torch.float32) if isinstance(current_width, torch.Tensor) else current_width
    lo = torch.clamp(idx - w, max_of_xmin, min_of_xmax).to(torch.int64).to("cuda")
    hi = torch.clamp(idx + w, max_of_xmin, min_of_xmax).to(torch.int64).to("cuda")
    lo, hi = lo.expand(bg.shape), hi.expand(bg.shape)
    avg = (torch.gather(bg, -1, lo) + torch.gather(bg, -1, hi)) / 2
    out = torch.minimum(avg, bg)
    return out if torch.device(device).type == "cuda" else out.to(device)


def snip_bkg(spec, er, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime, boxcar_size=5,
             device="cpu"):
    """SNIP background of spec[..., C] in counts; `er` = (first, last) global channel of the crop."""
    _native.require_device()
    if not isinstance(spec, torch.Tensor):
        spec = torch.as_tensor(np.asarray(spec))
    if spec.dtype not in (torch.float32, torch.float16):
        spec = spec.to(torch.float32)
    spec_dev = spec.detach().to("cuda")
    n_ch = spec_dev.shape[-1]
    flat = spec_dev.reshape(-1, n_ch).contiguous()
    lohi, n_pass = snip_tables_device(n_ch, er, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime,
                                      aux=flat.shape[0] >= 64)
    out = torch.empty((flat.shape[0], n_ch), dtype=torch.float32, device="cuda")
    snip_launch(flat, 0, n_ch, lohi, n_pass, boxcar_size, out, _native.MT_SNIP_BACKGROUND)
    out = out.reshape(spec_dev.shape)
    return out if torch.device(device).type == "cuda" else out.to(device)
