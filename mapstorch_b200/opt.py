"""Fitting entry points -- the reference's ``mapstorch/opt.py`` surface plus the per-pixel ``fit_map``.

``init_elem_amps`` (opt.py:63-97), ``create_tensors`` (opt.py:100-162) and ``fit_spec``
(opt.py:165-327) keep their signatures.  ``fit_map`` is the entry point the reference lacks
(SURVEY.md 0.2): it fits every pixel of a spectral volume.  For fixed global parameters the
model is linear in the amplitudes 10**a and the Gram matrix is pixel independent, so the
reference's per-spectrum Adam loop is replaced by

    K1 basis B[E,C]  ->  host (fp64): G = B B^T, H = G^-1, W = H B, hi/lo operand split
    [K2 SNIP residual]  ->  K3 tensor-core contraction X = Y W^T  ->  batched non-negative fix-up

which is what the reference's fit converges to.  No CPU compute path exists.
"""
import ctypes

import numpy as np
import torch

from mapstorch_b200 import _native, bkg as _bkg, map as _map, tables
from mapstorch_b200.default import (
    default_fitting_elems,
    default_fitting_params,
    default_param_vals,
    default_learning_rates,
    default_energy_consts,
    default_elem_info,
)

DIGIT_BITS = 10    # bits per operand plane of the pseudo-inverse (|digit| <= 512: exact in fp16 / tf32)
N_PLANES = 3       # 30 bits in total
MAX_GROUP = 80     # components per tensor-core tile: N = N_PLANES * e_pad <= 256
PLANE_RATIO = 2.0 ** -DIGIT_BITS
SNIP_CHUNK_PX = 131072  # pixels per scratch slab of the SNIP residual / split operand (2 GB of fp16 planes at 4096 channels): fewer, fuller launches of K2 and K3


def init_elem_amps(elem, spec, e_sct, e_offset, e_slope, e_consts=default_energy_consts, verbose=False, *,
                   compton_angle=None):
    """log10 amplitude guess from the mean counts in a window around the element's first line
    (opt.py:63-97).  Host-side NumPy like the reference; vectorises over leading dims of `spec`."""
    try:
        lo, hi = tables.amp_window(elem, spec.shape[-1], e_sct, e_offset, e_slope, e_consts, compton_angle)
        size = (hi + 1) - lo
        mean = np.sum(spec[..., lo: hi + 1], axis=-1) / size
        return np.log10(np.maximum(mean * 8.0 + 0.01, 1.0))
    except Exception:
        if verbose:
            print("Cannot initialize {} amplitude".format(elem))
        return 0.0


def create_tensors(elems_to_fit=default_fitting_elems, fitting_params=default_fitting_params,
                   init_param_vals=default_param_vals, fixed_param_vals={}, map_shape=None, tune_params=True,
                   init_amp=False, spec=None, device="cpu", verbose=False):
    """Parameter / amplitude tensor dict and per-parameter optimiser groups (opt.py:100-162)."""
    if init_amp:
        assert spec is not None, "Initial amplitude initialization requires the spectrum to be provided."
    tensors, opt_configs = {}, []
    for name, default in default_param_vals.items():
        start = init_param_vals.get(name, default)
        tunable = name in fitting_params and tune_params and name not in fixed_param_vals
        value = float(start) if tunable else fixed_param_vals.get(name, start)
        tensors[name] = torch.tensor(value, requires_grad=tunable, device=device)
        if tunable:
            opt_configs.append({"params": tensors[name],
                                "lr": default_learning_rates.get(name, default_learning_rates["default_lr"])})
    for elem in elems_to_fit:
        if init_amp:
            angle = tensors["COMPTON_ANGLE"].item() if "COMPTON_ANGLE" in tensors else None
            start = init_elem_amps(elem, spec, tensors["COHERENT_SCT_ENERGY"].item(), tensors["ENERGY_OFFSET"].item(),
                                   tensors["ENERGY_SLOPE"].item(), verbose=verbose, compton_angle=angle)
        else:
            start = np.zeros(map_shape) if map_shape is not None else 0.0
        tensors[elem] = torch.tensor(start, requires_grad=True, device=device)
        opt_configs.append({"params": tensors[elem],
                            "lr": default_learning_rates.get(elem, default_learning_rates["default_lr"])})
    return tensors, opt_configs


# ---- operand packing ---------------------------------------------------------------------------
def digit_planes(w64, n_planes=N_PLANES, bits=DIGIT_BITS):
    """W (float64 [E, C]) -> (colscale[E] fp32, digits int [n_planes, E, C]) with
    W ~= colscale * sum_d 2**(-bits*d) * digits[d]  and  |digits| <= 2**(bits-1).

    Balanced base-2**bits digits of W scaled row-wise by a power of two: each plane is exactly
    representable in fp16 and tf32, products with integer counts are integers, and the truncation
    error is <= 2**(-bits*n_planes) of the row maximum."""
    peak = np.max(np.abs(w64), axis=1)
    peak = np.where(peak > 0, peak, 1.0)
    scale = np.exp2(np.ceil(np.log2(peak)) - (bits - 1))
    rest = w64 / scale[:, None]
    planes = []
    for _ in range(n_planes):
        digit = np.rint(rest)
        planes.append(digit)
        rest = (rest - digit) * (2.0 ** bits)
    return scale.astype(np.float32), np.stack(planes)


def _shift_peers(peers, n_px, px_stride=1):
    """The same peer destinations, advanced by n_px pixels (px_stride fp32 elements per pixel)."""
    if peers is None or n_px == 0:
        return peers
    moved = _native.MtPeers()
    moved.n_peers, moved.n_owned = peers.n_peers, peers.n_owned
    step = 4 * n_px * px_stride
    for k in range(peers.n_peers):
        moved.peer[k] = peers.peer[k] + step
    moved.multicast = peers.multicast + step if peers.multicast else None
    for e in range(peers.n_owned):
        moved.owner_row[e] = peers.owner_row[e] + 4 * n_px if peers.owner_row[e] else None
    return moved


class MapFitPlan:
    """Everything that depends only on the global parameters, prepared once per fit.

    Attributes: names (component order = model_spec order), basis (CUDA fp32 [E, C]), gram/hinv
    (fp64), er (inclusive channel pair)."""

    def __init__(self, energy_range, elements_to_fit=default_fitting_elems, param_vals=default_param_vals,
                 indices=None, use_snip=True, e_consts=default_energy_consts, elem_info=default_elem_info,
                 detector_type="Si", escape_factor=0.0, boxcar_size=5):
        _native.require_device()
        self.er = (int(energy_range[0]), int(energy_range[1]))
        self.n_ch = self.er[1] - self.er[0] + 1
        self.params = dict(default_param_vals)
        self.params.update({k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in param_vals.items()
                            if k in default_param_vals})
        self.use_snip = bool(use_snip)
        self.boxcar_size = int(boxcar_size)
        self.detector_type, self.escape_factor = detector_type, float(escape_factor)
        self.e_consts, self.elem_info = e_consts, elem_info  # the refinement builds its line tables from them
        elements = list(dict.fromkeys(list(elements_to_fit)))
        self.basis, self.names = _map.element_basis(self.params, self.er, elements, e_consts, elem_info,
                                                    detector_type, escape_factor)
        b64 = self.basis.double().cpu().numpy()
        self.chan_mask = None  # float 0/1 per channel of the window when `indices` restricts the fit
        if indices is not None:
            mask = np.zeros(self.n_ch, dtype=bool)
            mask[np.asarray(indices)] = True
            b64 = b64 * mask[None, :]
            self.chan_mask = torch.from_numpy(mask.astype(np.float32)).cuda()
        # components without any open line (or fully masked) cannot be fitted: amplitude 0
        self.live = np.nonzero(np.sum(b64 * b64, axis=1) > 0)[0]
        bl = b64[self.live]
        gram = bl @ bl.T
        hinv = np.linalg.inv(gram)
        hinv = 0.5 * (hinv + hinv.T)
        self.gram, self.hinv = gram, hinv
        self.pinv = hinv @ bl  # [E_live, C]: x = pinv @ (y - bkg)
        self.n_live = len(self.live)
        # N = N_PLANES * e_pad must be a multiple of 16 and <= 256: one tensor-core tile holds 80 components;
        # longer lists (the 93 default components at high incident energy) are contracted in groups
        self.groups = [(g0, min(g0 + MAX_GROUP, self.n_live)) for g0 in range(0, self.n_live, MAX_GROUP)]
        self.e_pad = max(16, -(-min(self.n_live, MAX_GROUP) // 16) * 16)
        if self.n_live > 104:
            raise _native.NativeError(-3, f"{self.n_live} fitted components exceed the 104 the non-negative solver holds")
        self.gram_dev = torch.from_numpy(gram).cuda()
        self.hinv_dev = torch.from_numpy(hinv).cuda()
        self._packs = {}
        self._snip = None
        self._neg = None
        self._last_engine = None
        self._last_wide = False
        self._last_snip_half = False
        self.snip_planes = "f16"  # residual operand planes of the SNIP path: "f16" (4 B per channel) or "tf32" (8 B)
        self._side = None
        self.pipeline = False
        self.epilogue_fixup = True  # K3 epilogue fixes pixels whose zero set has <= 2 components itself
        self._h32 = None
        self._pixel_major = False
        self.timers = None  # set to a list to collect (kind, start, end, n_px) CUDA events around K2 / K3 launches
        if self.use_snip:
            p = self.params
            self._snip = _bkg.snip_tables_device(
                self.n_ch, self.er, *(torch.tensor(float(p[k])) for k in (
                    "ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME")))

    # -- operand packs, cached per (dtype, layout) ------------------------------------------------
    def _pack(self, kind, n_cols, col0, repeat, group=0):
        """(wpack [N_PLANES*e_pad, n_cols], colscale, e_pad) of one component group: digit planes of its
        pinv rows at columns [col0, col0+n_ch) (`repeat` copies back to back), in the volume's operand dtype."""
        key = (kind, n_cols, col0, repeat, group)
        if key not in self._packs:
            g0, g1 = self.groups[group]
            n, e_pad = g1 - g0, max(16, -(-(g1 - g0) // 16) * 16)
            scale, planes = digit_planes(self.pinv[g0:g1])
            dt = np.float16 if kind == "f16" else np.float32
            pack = np.zeros((N_PLANES * e_pad, n_cols), dtype=dt)
            for r in range(repeat):
                c = col0 + r * self.n_ch
                for d in range(N_PLANES):
                    pack[d * e_pad: d * e_pad + n, c: c + self.n_ch] = planes[d]
            self._packs[key] = (torch.from_numpy(pack).cuda(), torch.from_numpy(scale).cuda(), e_pad)
        return self._packs[key]

    def _pinv_f32(self, n_cols, col0):
        key = ("simt", n_cols, col0)
        if key not in self._packs:
            w = np.zeros((self.n_live, n_cols), dtype=np.float32)
            w[:, col0: col0 + self.n_ch] = self.pinv.astype(np.float32)
            self._packs[key] = torch.from_numpy(w).cuda()
        return self._packs[key]

    def _hinv32(self):
        if self._h32 is None:
            self._h32 = torch.from_numpy(self.hinv.astype(np.float32)).cuda()
        return self._h32

    def _neg_buffers(self, n_px, n_counters=1):
        """(pixel list, counters) the K3 epilogue fills; grown on demand, reused across calls."""
        if self._neg is None or self._neg[0].numel() < n_px or self._neg[1].numel() < n_counters:
            self._neg = (torch.empty(n_px, dtype=torch.int32, device="cuda"),
                         torch.zeros(max(n_counters, 64), dtype=torch.int32, device="cuda"))
        return self._neg[0], self._neg[1][:n_counters]

    # -- the fit -----------------------------------------------------------------------------------
    def _x_args(self, x):
        """(pointer, leading dimension, layout code) of an amplitude array in either layout."""
        return x.data_ptr(), x.stride(0), (_native.MT_X_PIXEL_MAJOR if self._pixel_major else _native.MT_X_ELEMENT_MAJOR)

    def _x_chunk(self, x, p0, n):
        return x[p0: p0 + n] if self._pixel_major else x[:, p0: p0 + n]

    def _timed(self, kind, n_px, stream, fn):
        """Run fn(); with self.timers set, bracket it with CUDA events on the launching stream."""
        if self.timers is None:
            return fn()
        s = stream if stream is not None else torch.cuda.current_stream()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(s)
        fn()
        t1.record(s)
        self.timers.append((kind, t0, t1, n_px))

    def _contract(self, spec2d, x, col0, repeat, k_begin, k_end, engine, stream, neg=None, px_offset=0, peers=None):
        """x[e, p] = sum_k spec2d[p, k] * pinv-pack[e, k] for the live components."""
        self._timed("contract", spec2d.shape[0], stream,
                    lambda: self._contract_untimed(spec2d, x, col0, repeat, k_begin, k_end, engine, stream, neg,
                                                   px_offset, peers))

    def _contract_untimed(self, spec2d, x, col0, repeat, k_begin, k_end, engine, stream, neg=None, px_offset=0,
                          peers=None):
        n_px, n_ch_total, ld = spec2d.shape[0], spec2d.shape[1], spec2d.stride(0)
        es = spec2d.element_size()
        tma_ok = (ld * es) % 16 == 0 and spec2d.data_ptr() % 16 == 0
        if engine == "tensor" and not tma_ok:
            engine = "simt"
        self._last_engine = engine
        if engine == "tensor":
            kind = "f16" if spec2d.dtype == torch.float16 else "tf32"
            if len(self.groups) > 1 and (self._pixel_major or peers is not None):
                raise NotImplementedError("more than 80 components: element-major output without peer stores only")
            for gi, (g0, g1) in enumerate(self.groups):
                pack, scale, e_pad = self._pack(kind, -(-n_ch_total // 8) * 8, col0, repeat, gi)
                xg = x if len(self.groups) == 1 else x[g0:g1]
                _native.call("mt_fit_contract", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), n_px, ld,
                             n_ch_total, int(k_begin), int(k_end), pack.data_ptr(), pack.stride(0), e_pad,
                             g1 - g0, N_PLANES, scale.data_ptr(), float(PLANE_RATIO), *self._x_args(xg),
                             int(px_offset), neg[0].data_ptr() if neg is not None else None,
                             neg[1].data_ptr() if neg is not None else None,
                             self._hinv32().data_ptr() if (neg is not None and self.epilogue_fixup and self.e_pad <= 32
                                                           and not self._pixel_major) else None,
                             ctypes.byref(peers) if peers is not None else None, _native.stream_ptr(stream))
        elif engine == "simt":
            if peers is not None:
                raise _native.NativeError(-3, "peer-memory destinations need the tensor-core path")
            if repeat != 1:
                raise _native.NativeError(-3, "the CUDA-core contraction expects an unsplit residual")
            w = self._pinv_f32(n_ch_total, col0)
            _native.call("mt_fit_contract_simt", spec2d.data_ptr(), _native.dtype_code(spec2d.dtype), n_px, ld,
                         int(k_begin), int(k_end), w.data_ptr(), w.stride(0), self.n_live, *self._x_args(x),
                         _native.stream_ptr(stream))
        else:
            raise ValueError(f"unknown engine {engine!r}")

    def _operand_inexact(self, spec2d, flag, stream=None):
        """flag (int32 [1], device) |= 1 if the fitted window of an fp32 volume holds values a tf32 operand
        cannot carry exactly (more than 11 significant bits: counts >= 2048, normalised / corrected data)."""
        _native.call("mt_operand_inexact", spec2d.data_ptr(), spec2d.shape[0], spec2d.stride(0), self.er[0], self.n_ch,
                     flag.data_ptr(), _native.stream_ptr(stream))

    def _penalty_coef(self, l1_lambda):
        """coef[e] of mt_penalty_shift: l1_lambda / (2 n_elements) * (G^-1 1)[e] for the live components;
        n_elements counts every fitted name like the reference's len(elements) (opt.py:377-381)."""
        coef = (float(l1_lambda) / (2.0 * max(len(self.names), 1))) * self.hinv.sum(axis=1)
        return torch.from_numpy(coef.astype(np.float32)).cuda()

    def _shift(self, operand, col0, repeat, x, coef, stream):
        _native.call("mt_penalty_shift", operand.data_ptr(), _native.dtype_code(operand.dtype), operand.shape[0],
                     operand.stride(0), int(col0), self.n_ch, int(repeat),
                     self.chan_mask.data_ptr() if self.chan_mask is not None else None, coef.data_ptr(),
                     *self._x_args(x), self.n_live, _native.stream_ptr(stream))

    def _fit_serial(self, spec2d, x, nonneg, engine, stream, n_fixed, exact_counts, peers, coef=None, snip_overflow=None):
        n_px = spec2d.shape[0]
        # the tensor-core epilogue compacts the pixels that need the non-negative fix-up
        neg = None
        if nonneg and engine == "tensor" and n_px and len(self.groups) == 1 and coef is None:
            neg = self._neg_buffers(n_px)
            neg[1].zero_()
        wide = False
        if not self.use_snip and engine == "tensor" and spec2d.dtype == torch.float32 and exact_counts is not True:
            if exact_counts == "split":
                wide = True
            else:
                flag = torch.zeros(1, dtype=torch.int32, device="cuda")
                self._operand_inexact(spec2d, flag, stream)
                wide = bool(flag.item())
        self._last_wide = wide
        if wide:
            # exact for any magnitude: stream the volume as a tf32-exact high part plus remainder
            width = 2 * self.n_ch
            pitch = -(-width // 4) * 4
            chunk = min(SNIP_CHUNK_PX, n_px)
            scratch = torch.empty((chunk, pitch), dtype=torch.float32, device="cuda")[:, :width]
            for p0 in range(0, n_px, chunk):
                n = min(chunk, n_px - p0)
                _native.call("mt_split_hilo", spec2d[p0: p0 + n].data_ptr(), n, spec2d.stride(0), self.er[0], self.n_ch,
                             scratch.data_ptr(), scratch.stride(0), _native.stream_ptr(stream))
                self._contract(scratch[:n], self._x_chunk(x, p0, n), 0, 2, 0, width, engine, stream, neg, p0,
                               _shift_peers(peers, p0, x.stride(0) if self._pixel_major else 1))
                if coef is not None:
                    self._shift(scratch[:n], 0, 2, self._x_chunk(x, p0, n), coef, stream)
        elif not self.use_snip:
            self._contract(spec2d, x, self.er[0], 1, self.er[0], self.er[1] + 1, engine, stream, neg, 0, peers)
            if coef is not None:
                self._shift(spec2d, self.er[0], 1, x, coef, stream)
        else:
            lohi, n_pass = self._snip
            hilo = engine == "tensor"
            # tensor-core path: the residual crosses HBM as two fp16 planes (4 bytes per channel), unless it does not
            # fit fp16 (|y - bkg| > 65504: the kernel raises a flag) -- then as a tf32-exact high part + fp32 remainder
            half = hilo and self.snip_planes == "f16" and exact_counts != "split"
            width = 2 * self.n_ch if hilo else self.n_ch
            chunk = min(SNIP_CHUNK_PX, n_px)
            starts = list(range(0, n_px, chunk))
            flags = None
            if half:
                flags = snip_overflow if snip_overflow is not None else torch.zeros(1, dtype=torch.int32, device="cuda")
            self._last_snip_half = half

            def run_chunks(which, as_half):
                es_pitch = 8 if as_half else 4  # 16-byte row pitch so that TMA can address the scratch rows
                pitch = -(-width // es_pitch) * es_pitch
                scratch = torch.empty((chunk, pitch), dtype=torch.float16 if as_half else torch.float32, device="cuda")[:, :width]
                mode = _native.MT_SNIP_RESIDUAL_HILO if hilo else _native.MT_SNIP_RESIDUAL
                for p0 in which:
                    n = min(chunk, n_px - p0)
                    if as_half:
                        self._timed("snip", n, stream,
                                    lambda: _bkg.snip_residual_f16(spec2d[p0: p0 + n], self.er[0], self.n_ch, lohi, n_pass,
                                                                   self.boxcar_size, scratch[:n], flags, stream))
                    else:
                        self._timed("snip", n, stream,
                                    lambda: _bkg.snip_launch(spec2d[p0: p0 + n], self.er[0], self.n_ch, lohi, n_pass,
                                                             self.boxcar_size, scratch[:n], mode, stream))
                    self._contract(scratch[:n], self._x_chunk(x, p0, n), 0, 2 if hilo else 1, 0, width, engine, stream, neg, p0,
                                   _shift_peers(peers, p0, x.stride(0) if self._pixel_major else 1))
                    if coef is not None:
                        self._shift(scratch[:n], 0, 2 if hilo else 1, self._x_chunk(x, p0, n), coef, stream)

            run_chunks(starts, half)
            # fp16 volumes cannot overflow; exact_counts=True (graph replays, streamed slabs) defers the test to the caller
            if half and snip_overflow is None and spec2d.dtype != torch.float16 and exact_counts is not True:
                if bool(flags.item()):
                    if neg is not None:
                        neg[1].zero_()
                    self._last_snip_half = False
                    run_chunks(starts, False)
        if nonneg and n_px:
            self._nonneg_tail(x, n_px, neg, n_fixed, peers, stream)

    def _nonneg_tail(self, x, n_px, neg, n_fixed, peers, stream):
        """Batched non-negative solve after a contraction: driven by the pixel list the K3 epilogue compacted (neg),
        or a scan of all pixels."""
        if self._last_engine == "tensor" and neg is not None:
            _native.call("mt_nnls_list", *self._x_args(x), neg[0].data_ptr(), neg[1].data_ptr(),
                         int(self.epilogue_fixup and self.e_pad <= 32), self.n_live, self.gram_dev.data_ptr(), self.hinv_dev.data_ptr(),
                         ctypes.byref(peers) if peers is not None else None, _native.stream_ptr(stream))
            if n_fixed is not None:
                n_fixed.copy_(neg[1])
        else:
            _native.call("mt_nnls_fixup", *self._x_args(x), n_px, self.n_live, self.gram_dev.data_ptr(),
                         self.hinv_dev.data_ptr(), n_fixed.data_ptr() if n_fixed is not None else None,
                         ctypes.byref(peers) if peers is not None else None, _native.stream_ptr(stream))

    def fit_operand_hilo(self, q2d, x, nonneg=True, stream=None):
        """x [n_live, P] <- (non-negative) least squares of a target given as operand planes q2d [P, 2 * n_ch] fp32
        (tf32-exact high part | remainder, the layout of MT_SNIP_RESIDUAL_HILO): contraction on the tensor cores +
        fix-up.  The x-update of the per-pixel loss="l1" iteration (robust.fit_map_lad)."""
        n_px = q2d.shape[0]
        self._pixel_major = False
        neg = None
        if nonneg and n_px and len(self.groups) == 1:
            neg = self._neg_buffers(n_px)
            neg[1].zero_()
        self._contract(q2d, x, 0, 2, 0, 2 * self.n_ch, "tensor", stream, neg, 0, None)
        if nonneg and n_px:
            self._nonneg_tail(x, n_px, neg, None, None, stream)
        return x

    def _fit_pipelined(self, spec2d, x, n_fixed, peers, rounds_per_chunk=4):
        """Contraction and non-negative solve overlapped: the volume is cut into chunks of whole
        waves of the persistent kernel (148 CTAs x 256 pixels), chunk k's fix-up runs on a side
        stream while chunk k+1 streams from HBM (the solve is latency bound, the contraction
        bandwidth bound, and the contraction leaves room on every SM for the solve's CTAs)."""
        n_px = spec2d.shape[0]
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        sms = torch.cuda.get_device_properties(spec2d.device).multi_processor_count
        chunk = sms * 256 * rounds_per_chunk
        starts = list(range(0, n_px, chunk))
        lists, counts = self._neg_buffers(n_px, len(starts))
        counts.zero_()
        self._last_wide = False
        for c, p0 in enumerate(starts):
            n = min(chunk, n_px - p0)
            self._contract(spec2d[p0: p0 + n], self._x_chunk(x, p0, n), self.er[0], 1, self.er[0], self.er[1] + 1, "tensor",
                           None, (lists[p0: p0 + n], counts[c: c + 1]), p0, _shift_peers(peers, p0, x.stride(0) if self._pixel_major else 1))
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                _native.call("mt_nnls_list", *self._x_args(x), lists[p0: p0 + n].data_ptr(),
                             counts[c: c + 1].data_ptr(), int(self.epilogue_fixup and self.e_pad <= 32), self.n_live, self.gram_dev.data_ptr(),
                             self.hinv_dev.data_ptr(), ctypes.byref(peers) if peers is not None else None,
                             _native.stream_ptr(side))
        joined = torch.cuda.Event()
        joined.record(side)
        main.wait_event(joined)
        if n_fixed is not None:
            n_fixed.copy_(counts.sum().reshape(1))

    def fit(self, spec2d, out=None, nonneg=True, engine="tensor", stream=None, n_fixed=None, exact_counts=None,
            peers=None, layout="ep", l1_lambda=0.0, snip_overflow=None):
        """spec2d: CUDA [P, C_full] fp32/fp16 with contiguous rows.  Returns x [E, P] fp32 (CUDA),
        rows in `self.names` order (components without an open line stay 0).

        exact_counts: the tensor-core operands (fp16 / tf32) hold 11 significant bits, i.e. integer
        counts below 2048 exactly.  fp16 volumes satisfy that by construction and the SNIP path
        splits its residual exactly; for fp32 volumes without SNIP, None (default) tests every value of
        the window for tf32 representability on the device (mt_operand_inexact: one extra read of the
        volume and a host sync) and routes volumes that fail -- large counts, but also normalised or
        dead-time corrected data -- through an exact hi/lo split; True asserts that the volume holds
        integer counts below 2048 and skips the test (other data would be truncated to 11 bits).

        peers: optional _native.MtPeers whose pointers address, on the other GPUs of the box, the
        element of their map buffers that corresponds to out[0, 0]; the kernels then store every
        amplitude there as well (dist.PeerMaps) -- the all-gather of the maps fused into the fit.

        layout: "ep" (default) returns x [E, P], one contiguous map per component; "pe" returns
        x [P, e_pad], one 16-byte aligned record per pixel (columns >= E are zero) -- the layout
        whose stores are 128-bit wide, used for the peer-memory gather.

        snip_overflow: optional int32 [1] device flag.  On the SNIP path the residual is handed to the contraction as
        two fp16 planes; a residual beyond the fp16 range (> 65504 counts) raises the flag.  Without this argument
        fit() reads the flag itself (one host sync per call for fp32 volumes) and repeats the fit with tf32 planes;
        with it the caller does (fit_host does so once per volume); exact_counts="split" forces tf32 planes.

        l1_lambda > 0 adds the reference's amplitude-sparsity penalty (opt.py:372-396) to the MSE objective:
        the unconstrained amplitudes are shifted per pixel (mt_penalty_shift, one more read of the spectra)
        before the non-negative solve."""
        if l1_lambda > 0 and (not nonneg or peers is not None):
            raise ValueError("the l1_lambda penalty needs the non-negative solve and local output")
        if spec2d.device.type != "cuda" or spec2d.dim() != 2 or spec2d.stride(1) != 1:
            raise ValueError("spec2d must be a CUDA [P, C] tensor with contiguous rows")
        if spec2d.shape[1] <= self.er[1]:
            raise ValueError(f"spectra have {spec2d.shape[1]} channels, energy_range ends at {self.er[1]}")
        n_px, n_all = spec2d.shape[0], len(self.names)
        dense = self.n_live == n_all
        self._pixel_major = layout == "pe"
        if self._pixel_major:
            if not dense:
                raise NotImplementedError("pixel-major output with components that have no open line")
            if out is None:
                out = torch.zeros((n_px, self.e_pad), dtype=torch.float32, device="cuda")
            if out.shape[0] != n_px or out.stride(0) < self.e_pad or out.stride(0) % 4 or out.stride(1) != 1:
                raise ValueError("pixel-major out must be [P, >= e_pad] with a row pitch that is a multiple of 4")
        elif out is None:
            out = torch.zeros((n_all, n_px), dtype=torch.float32, device="cuda")
        if peers is not None and not dense:
            raise NotImplementedError("peer-memory destinations with components that have no open line")
        x = out if dense else torch.empty((self.n_live, n_px), dtype=torch.float32, device="cuda")
        # measured on B200 (c5 shard): chunking costs the contraction more (0.77 vs 0.72 ms) than the
        # overlap hides (0.09 ms), so the pipelined variant is opt-in
        pipelined = (self.pipeline and nonneg and engine == "tensor" and n_px > 0 and not self.use_snip and stream is None
                     and (spec2d.dtype == torch.float16 or exact_counts is True)
                     and (spec2d.stride(0) * spec2d.element_size()) % 16 == 0 and spec2d.data_ptr() % 16 == 0)
        if pipelined and not l1_lambda > 0:
            self._fit_pipelined(spec2d, x, n_fixed, peers)
        else:
            self._fit_serial(spec2d, x, nonneg, engine, stream, n_fixed, exact_counts, peers,
                             self._penalty_coef(l1_lambda) if l1_lambda > 0 else None, snip_overflow)
        if not dense:
            out.zero_()  # components without an open line (caller-provided buffers included)
            out[torch.as_tensor(self.live, device="cuda")] = x
        return out


def _fit_slabs(self, n_px, n_ch, dtype, load, nonneg=True, engine="tensor", chunk_px=16384, l1_lambda=0.0, after_slab=None,
               int_spec=None):
    """Stream a volume that is NOT resident on the device through it in slabs of pixels: `load(p0, n, dst)` brings
    pixels [p0, p0 + n) into the device buffer dst [n, n_ch] (energy last) on the current stream -- a host-to-device
    copy, plus the transposition of a channel-first slab.  Slab k+1 is loaded on a second stream while slab k is
    being fitted.  Returns x [E, P] on the device.

    after_slab(dev_slab, p0, n, x_slab), if given, runs on the compute stream right after the fit of each
    slab while the slab is still resident (per-pixel refinement uses it), so the volume crosses PCIe once.
    int_spec: optional fp64 [n_ch] device accumulator; the integrated spectrum (io.py:159) is summed slab by slab
    as a by-product of the pass (mt_int_spec)."""
    out = torch.zeros((len(self.names), n_px), dtype=torch.float32, device="cuda")
    chunk = max(1, min(int(chunk_px), n_px))
    compute = torch.cuda.current_stream()
    copier = torch.cuda.Stream()
    bufs = [torch.empty((chunk, n_ch), dtype=dtype, device="cuda") for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    # fp32 volumes without SNIP: the operand-exactness check (counts < 2048) must not stall the pipeline, so
    # every slab is fitted optimistically, its maximum lands in a device flag, and the (rare) slabs that turn
    # out to hold larger counts are redone through the exact split path after ONE synchronisation at the end
    check = dtype == torch.float32 and not self.use_snip and engine == "tensor"
    # SNIP path, fp32 volumes: the fp16 residual planes overflow beyond 65504 counts; same deferred treatment
    snip_check = dtype == torch.float32 and self.use_snip and engine == "tensor" and self.snip_planes == "f16"
    starts = list(range(0, n_px, chunk))
    too_big = torch.zeros(len(starts), dtype=torch.int32, device="cuda") if (check or snip_check) else None
    for i, p0 in enumerate(starts):
        b, n = i & 1, min(chunk, n_px - p0)
        with torch.cuda.stream(copier):
            if i >= 2:
                copier.wait_event(consumed[b])
            load(p0, n, bufs[b][:n])
            copied[b].record(copier)
        compute.wait_event(copied[b])
        if int_spec is not None:
            _native.call("mt_int_spec", bufs[b].data_ptr(), _native.dtype_code(dtype), n, bufs[b].stride(0), n_ch,
                         int_spec.data_ptr(), _native.stream_ptr())
        if check:
            self._operand_inexact(bufs[b][:n], too_big[i: i + 1])
        self.fit(bufs[b][:n], out=out[:, p0: p0 + n], nonneg=nonneg, engine=engine,
                 exact_counts=True if check else None, l1_lambda=l1_lambda,
                 snip_overflow=too_big[i: i + 1] if snip_check else None)
        if after_slab is not None:
            after_slab(bufs[b][:n], p0, n, out[:, p0: p0 + n])
        consumed[b].record(compute)
    if check or snip_check:
        for i in torch.nonzero(too_big).flatten().tolist():
            p0 = starts[i]
            n = min(chunk, n_px - p0)
            load(p0, n, bufs[0][:n])
            self.fit(bufs[0][:n], out=out[:, p0: p0 + n], nonneg=nonneg, engine=engine, exact_counts="split",
                     l1_lambda=l1_lambda)
            if after_slab is not None:
                after_slab(bufs[0][:n], p0, n, out[:, p0: p0 + n])
    for buf in bufs:
        buf.record_stream(compute)
    return out


def _fit_host(self, host2d, nonneg=True, engine="tensor", chunk_px=16384, l1_lambda=0.0, after_slab=None, int_spec=None):
    """Stream a HOST volume [P, C] (energy last) through the device in row slabs (pinned host memory makes the
    copies asynchronous); see fit_slabs."""
    n_px, n_ch = host2d.shape

    def load(p0, n, dst):
        dst.copy_(host2d[p0: p0 + n], non_blocking=True)

    return self.fit_slabs(n_px, n_ch, host2d.dtype, load, nonneg, engine, chunk_px, l1_lambda, after_slab, int_spec)


class ChannelFirstLoader:
    """load() of fit_slabs for volumes stored channel-first, [C, H, W] (raw XRF-Maps MAPS/Spectra/mca_arr): a slab of
    whole rows [C, rows, W] is staged in pinned memory (that is where a memory-mapped file is actually read), copied
    to the device and transposed there to [rows * W, C] (mt_transpose_cf) -- the reference does this on the host with
    np.moveaxis over the whole volume (io.py:154-156).  `source[c, r0:r1, :]` may be a NumPy array / memmap, an
    h5py dataset or a torch tensor (host or CUDA)."""

    def __init__(self, source, n_rows, width, n_ch, dtype, rows_per_slab):
        self.source, self.n_rows, self.width, self.n_ch, self.dtype = source, n_rows, width, n_ch, dtype
        self.rows = max(1, int(rows_per_slab))
        self.on_device = isinstance(source, torch.Tensor) and source.is_cuda
        self._stage, self._pinned, self._turn = None, None, 0
        self._staged = [None, None]

    @property
    def chunk_px(self):
        return self.rows * self.width

    def __call__(self, p0, n, dst):
        r0, rows = p0 // self.width, n // self.width
        assert p0 % self.width == 0 and n % self.width == 0, "channel-first slabs are whole rows"
        if self.on_device:
            src = self.source[:, r0: r0 + rows, :]
            flat, ld = src, self.source.stride(0)
            ptr = src.data_ptr()
        else:
            if self._stage is None:
                self._stage = [torch.empty((self.n_ch, self.rows * self.width), dtype=self.dtype, device="cuda") for _ in range(2)]
                self._pinned = [torch.empty((self.n_ch, self.rows * self.width), dtype=self.dtype, pin_memory=True)
                                for _ in range(2)]
            t = self._turn
            self._turn ^= 1
            if self._staged[t] is not None:
                self._staged[t].synchronize()  # the pinned buffer's previous copy has left the host
            pinned = self._pinned[t][:, :n].reshape(self.n_ch, rows, self.width)
            block = self.source[:, r0: r0 + rows, :]
            pinned.copy_(block if isinstance(block, torch.Tensor) else torch.from_numpy(np.array(block)))
            stage = self._stage[t][:, :n]
            stage.copy_(self._pinned[t][:, :n], non_blocking=True)
            self._staged[t] = torch.cuda.Event()
            self._staged[t].record()
            ptr, ld = stage.data_ptr(), self._stage[t].stride(0)
        _native.call("mt_transpose_cf", ptr, _native.dtype_code(self.dtype), self.n_ch, n, ld, dst.data_ptr(),
                     dst.stride(0), _native.stream_ptr())


class RowLoader:
    """load() of fit_slabs for energy-last volumes [H, W, C] that are read lazily (np.memmap, h5py dataset): each slab
    of whole rows is read into pinned memory -- the file is touched here, slab by slab, never as a whole -- and copied
    to the device asynchronously."""

    def __init__(self, source, n_rows, width, n_ch, dtype, rows_per_slab):
        self.source, self.n_rows, self.width, self.n_ch, self.dtype = source, n_rows, width, n_ch, dtype
        self.rows = max(1, int(rows_per_slab))
        self._pinned, self._turn, self._staged = None, 0, [None, None]

    @property
    def chunk_px(self):
        return self.rows * self.width

    def __call__(self, p0, n, dst):
        r0, rows = p0 // self.width, n // self.width
        assert p0 % self.width == 0 and n % self.width == 0, "file slabs are whole rows"
        if self._pinned is None:
            self._pinned = [torch.empty((self.rows * self.width, self.n_ch), dtype=self.dtype, pin_memory=True) for _ in range(2)]
        t = self._turn
        self._turn ^= 1
        if self._staged[t] is not None:
            self._staged[t].synchronize()
        block = np.asarray(self.source[r0: r0 + rows]).reshape(n, self.n_ch)
        self._pinned[t][:n].copy_(torch.from_numpy(np.array(block)))  # np.array: a writable copy of the mapped pages
        dst.copy_(self._pinned[t][:n], non_blocking=True)
        self._staged[t] = torch.cuda.Event()
        self._staged[t].record()


def _make_graph(self, spec2d, out, nonneg=True, engine="tensor"):
    """Capture one fit of a resident volume into a CUDA graph (the launch-bound inner loop of a
    repeated fit: two kernel launches per step replayed without host work).  Returns the
    torch.cuda.CUDAGraph; call .replay()."""
    saved, self.timers = self.timers, None
    self.fit(spec2d, out=out, nonneg=nonneg, engine=engine)  # warm caches; checks the count range once
    # no host sync inside the capture: the warm run above has settled which operand form the volume needs
    exact = "split" if (self._last_wide or (self.use_snip and engine == "tensor" and not self._last_snip_half)) else True
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        self.fit(spec2d, out=out, nonneg=nonneg, engine=engine, exact_counts=exact)
    self.timers = saved
    return graph


def _to_host(self, x):
    """Device -> pinned host copy of a result.  Every call returns a tensor the caller owns (plans are cached and
    reused across fit_map calls, so a shared staging buffer would let the next fit overwrite earlier results);
    torch's caching host allocator recycles the pinned block once the caller drops it."""
    host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    host.copy_(x, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host


MapFitPlan.fit_slabs = _fit_slabs
MapFitPlan.fit_host = _fit_host
MapFitPlan.make_graph = _make_graph
MapFitPlan.to_host = _to_host


_PLAN_CACHE = {}


def _cached_plan(energy_range, elements_to_fit, params, indices, use_snip, e_consts, elem_info, detector_type,
                 escape_factor):
    """Plans depend only on the global parameters; keep the last few (a dataset is usually fitted slab
    by slab, or map after map, with the same calibration)."""
    try:
        key = (tuple(int(v) for v in energy_range), tuple(elements_to_fit),
               tuple(sorted((k, float(v)) for k, v in params.items() if k in default_param_vals)),
               None if indices is None else np.asarray(indices).tobytes(), bool(use_snip), id(e_consts), id(elem_info),
               detector_type, float(escape_factor), torch.cuda.current_device())
    except (TypeError, ValueError):
        key = None
    if key is not None and key in _PLAN_CACHE:
        return _PLAN_CACHE[key]
    plan = MapFitPlan(energy_range, elements_to_fit, params, indices, use_snip, e_consts, elem_info, detector_type,
                      escape_factor)
    if key is not None:
        if len(_PLAN_CACHE) >= 8:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = plan
    return plan


def fit_map(spec_vol, energy_range, elements_to_fit=default_fitting_elems, init_param_vals=default_param_vals,
            fixed_param_vals={}, indices=None, use_snip=True, e_consts=default_energy_consts,
            elem_info=default_elem_info, detector_type="Si", escape_factor=0.0, nonneg=True, as_log10=False,
            device="cuda", engine="tensor", plan=None, return_plan=False, chunk_px=16384, refine=None,
            refine_iters=12, refine_bounds=None, loss="mse", l1_lambda=0.0, group=None, gather="all", energy_dim=2,
            rows_per_slab=None, int_spec=False):
    """Fit every pixel of spec_vol[..., C] with the global parameters held fixed.

    Equivalent to calling the reference's ``fit_spec(spec_vol[i, j], energy_range, elements_to_fit,
    init_param_vals=..., fixed_param_vals=..., tune_params=False, use_snip=..., indices=...)`` on each
    pixel and running it to convergence.  Returns ``{name: amplitude map [...]}`` (linear counts, or
    log10 like the reference's tensors if as_log10) on `device`; the two scatter components are
    always fitted, as in fit_spec (opt.py:194-200).

    Any float spectra are accepted, like the reference.  Integer counts below 2048 (every fp16 volume) run the
    single-pass tensor-core contraction with exact accumulation; fp32 volumes holding anything else (larger counts,
    normalised or corrected data) are detected on the device and contracted as an exact two-operand split.

    refine: optional tuple of calibration parameters (subset of ENERGY_OFFSET, ENERGY_SLOPE, FWHM_OFFSET,
    FWHM_FANOPRIME) to free PER PIXEL together with the amplitudes, like fit_spec(fitting_params=refine,
    tune_params=True) per pixel; their maps, the final "loss" and "iterations" are added to the result.

    l1_lambda: the reference's sparsity penalty on the amplitudes (opt.py:372-396), per pixel.
    loss="l1": the sum of absolute residuals instead of squares (opt.py:232-233, 371-375) for every pixel: ADMM whose
    x-update is the ordinary tensor-core map fit of a modified target (robust.fit_map_lad, csrc/mt_lad.cu); the
    per-pixel objective comes back as "loss".

    energy_dim: axis of spec_vol that holds the channels.  2 (default) is the layout create_dataset stores
    (io.py:154-156); 0 is the channel-first layout of raw XRF-Maps files (MAPS/Spectra/mca_arr, [C, H, W]): row slabs
    are then transposed on the device (mt_transpose_cf) instead of the reference's np.moveaxis over the whole volume on
    the host.  spec_vol may also be a lazily read array (np.memmap, h5py dataset, io.open_dataset(...).volume): it is
    streamed from the file slab by slab (rows_per_slab rows at a time), never loaded as a whole.

    int_spec=True adds "int_spec" to the result: the integrated spectrum over all pixels (io.py:159), summed on the
    device as a by-product of the streaming pass (fp64 accumulators; fp32 [C] result).

    group: a torch.distributed process group (one rank per GPU).  spec_vol is then THIS rank's block of rows of a
    row-sharded grid (dist.shard_rows) and the returned maps cover the whole grid: every component on every rank
    (gather="all"), each rank's own components ("owner": component e on rank e % world) or everything on rank 0
    ("root").  The exchange is fused into the fit kernels (peer stores over NVLink), see dist.ShardedMapFit."""
    if loss not in ("mse", "l1"):
        raise ValueError("Loss function not supported")
    if loss == "l1" and (refine or int_spec):
        raise NotImplementedError('fit_map(loss="l1") is the per-pixel L1 fit (no refine / int_spec)')
    _native.require_device()
    if group is not None:
        # multi-GPU form: spec_vol is this rank's block of rows; maps of the whole grid come back (dist.py)
        if as_log10 or plan is not None or return_plan or engine != "tensor" or energy_dim != 2 or int_spec:
            raise NotImplementedError("fit_map(group=...): no as_log10 / plan / return_plan / engine / energy_dim / int_spec")
        from mapstorch_b200.dist import fit_map_sharded

        if not isinstance(spec_vol, torch.Tensor):
            spec_vol = torch.from_numpy(np.asarray(spec_vol))
        maps = fit_map_sharded(spec_vol, energy_range, elements_to_fit, group=group, gather=gather,
                               init_param_vals=init_param_vals, fixed_param_vals=fixed_param_vals, indices=indices,
                               use_snip=use_snip, e_consts=e_consts, elem_info=elem_info, detector_type=detector_type,
                               escape_factor=escape_factor, nonneg=nonneg, l1_lambda=l1_lambda, loss=loss, refine=refine,
                               refine_iters=refine_iters, refine_bounds=refine_bounds)
        return maps if torch.device(device).type == "cuda" else {k: v.cpu() for k, v in maps.items()}
    if energy_dim not in (0, 1, 2):
        raise ValueError("energy_dim must be 0, 1, or 2")
    lazy = not isinstance(spec_vol, (torch.Tensor, np.ndarray)) or isinstance(spec_vol, np.memmap)
    if lazy and not (hasattr(spec_vol, "shape") and hasattr(spec_vol, "dtype") and len(spec_vol.shape) == 3):
        lazy = False
    loader = None
    if lazy or energy_dim == 0:
        # streamed sources: lazily read files (either layout) and channel-first volumes wherever they live
        if energy_dim == 1:
            raise NotImplementedError("energy_dim=1 for lazily read volumes; convert with io.create_dataset first")
        if isinstance(spec_vol, np.ndarray) and not isinstance(spec_vol, np.memmap):
            if spec_vol.dtype not in (np.float32, np.float16):
                spec_vol = spec_vol.astype(np.float32)
            spec_vol = torch.from_numpy(spec_vol)
        np_dtype = np.dtype(spec_vol.dtype) if not isinstance(spec_vol, torch.Tensor) else None
        if isinstance(spec_vol, torch.Tensor):
            if spec_vol.dtype not in (torch.float32, torch.float16):
                spec_vol = spec_vol.to(torch.float32)
            t_dtype = spec_vol.dtype
        else:
            t_dtype = torch.float16 if np_dtype == np.float16 else torch.float32
            if np_dtype not in (np.float16, np.float32):
                raise NotImplementedError(f"lazily read volumes must be stored as float16 or float32, not {np_dtype}")
        shape = tuple(int(v) for v in spec_vol.shape)
        if energy_dim == 0:
            n_ch_all, n_rows, width = shape
        else:
            n_rows, width, n_ch_all = shape
        lead = (n_rows, width)
        rows = rows_per_slab or max(1, int(chunk_px) // max(width, 1))
        cls = ChannelFirstLoader if energy_dim == 0 else RowLoader
        loader = cls(spec_vol, n_rows, width, n_ch_all, t_dtype, min(rows, n_rows))
        flat = None
    else:
        if not isinstance(spec_vol, torch.Tensor):
            arr = np.asarray(spec_vol)
            if arr.dtype not in (np.float32, np.float16):
                arr = arr.astype(np.float32)
            spec_vol = torch.from_numpy(arr)
        if spec_vol.dtype not in (torch.float32, torch.float16):
            spec_vol = spec_vol.to(torch.float32)
        if energy_dim == 1:
            spec_vol = spec_vol.movedim(1, 2)
        lead = spec_vol.shape[:-1]
        flat = spec_vol.reshape(-1, spec_vol.shape[-1])
        if flat.stride(1) != 1:
            flat = flat.contiguous()
    if plan is None:
        params = dict(init_param_vals)
        params.update(fixed_param_vals)
        plan = _cached_plan(energy_range, elements_to_fit, params, indices, use_snip, e_consts, elem_info,
                            detector_type, escape_factor)
    extra = {}
    refine_slab = None
    if refine:
        # per-pixel nonlinear refinement of peak-shape parameters + amplitudes (kernel K4), slab by slab so that
        # neither a host volume nor the SNIP background has to be resident as a whole
        from mapstorch_b200.refine import refine_map

        n_px = flat.shape[0] if flat is not None else lead[0] * lead[1]
        theta = torch.empty((len(refine), n_px), dtype=torch.float32, device="cuda")
        loss_px = torch.empty(n_px, dtype=torch.float32, device="cuda")
        iters = torch.empty(n_px, dtype=torch.int32, device="cuda")

        def refine_slab(dev, p0, n, x_slab):
            bkg = None
            if plan.use_snip:  # same tables as the fit (plan._snip), one kernel, no host work per slab
                bkg = torch.empty((n, plan.n_ch), dtype=torch.float32, device="cuda")
                _bkg.snip_launch(dev, plan.er[0], plan.n_ch, plan._snip[0], plan._snip[1], plan.boxcar_size, bkg,
                                 _native.MT_SNIP_BACKGROUND)
            t, l, it = refine_map(plan, dev, x_slab, free=tuple(refine), bounds=refine_bounds, max_iter=refine_iters,
                                  bkg=bkg)
            theta[:, p0: p0 + n], loss_px[p0: p0 + n], iters[p0: p0 + n] = t, l, it

        extra = {name: theta[i] for i, name in enumerate(refine)}
        extra["loss"], extra["iterations"] = loss_px, iters
    if loss == "l1":
        # sum of absolute residuals per pixel (opt.py:232-233): ADMM around the same contraction, robust.fit_map_lad
        from mapstorch_b200 import robust

        x, objective = robust.fit_map_lad(plan, flat, nonneg=nonneg, l1_lambda=l1_lambda, loader=loader,
                                          n_px=lead[0] * lead[1] if loader is not None else None)
        if as_log10:
            x = torch.log10(x)
        on_dev = torch.device(device).type == "cuda"
        maps = {name: (x[i] if on_dev else x[i].cpu()).reshape(lead) for i, name in enumerate(plan.names)}
        maps["loss"] = (objective if on_dev else objective.cpu()).reshape(lead)
        return (maps, plan) if return_plan else maps
    acc = None
    if int_spec:
        n_all = loader.n_ch if loader is not None else flat.shape[1]
        acc = torch.zeros(n_all, dtype=torch.float64, device="cuda")
    if loader is not None:
        x = plan.fit_slabs(lead[0] * lead[1], loader.n_ch, loader.dtype, loader, nonneg=nonneg, engine=engine,
                           chunk_px=loader.chunk_px, l1_lambda=l1_lambda, after_slab=refine_slab, int_spec=acc)
    elif flat.device.type == "cuda":
        if acc is not None:
            _native.call("mt_int_spec", flat.data_ptr(), _native.dtype_code(flat.dtype), flat.shape[0], flat.stride(0),
                         flat.shape[1], acc.data_ptr(), _native.stream_ptr())
        x = plan.fit(flat, nonneg=nonneg, engine=engine, l1_lambda=l1_lambda)
        if refine_slab is not None:
            for p0 in range(0, flat.shape[0], SNIP_CHUNK_PX):
                p1 = min(p0 + SNIP_CHUNK_PX, flat.shape[0])
                refine_slab(flat[p0:p1], p0, p1 - p0, x[:, p0:p1])
    else:
        # the slab is fitted and refined while it is resident: the volume crosses PCIe once, and the next slab's
        # copy overlaps both kernels
        x = plan.fit_host(flat, nonneg=nonneg, engine=engine, chunk_px=chunk_px, l1_lambda=l1_lambda,
                          after_slab=refine_slab, int_spec=acc)
    if as_log10:
        x = torch.log10(x)
    if torch.device(device).type != "cuda":
        x = plan.to_host(x)
    maps = {name: x[i].reshape(lead) for i, name in enumerate(plan.names)}
    for name, val in extra.items():
        val = val if torch.device(device).type == "cuda" else val.cpu()
        maps[name] = val.reshape(lead)
    if acc is not None:
        maps["int_spec"] = acc.float() if torch.device(device).type == "cuda" else acc.float().cpu()
    return (maps, plan) if return_plan else maps


def fit_dataset(file_path, energy_range, elements_to_fit=None, spec_vol_key=None, energy_dim=None, rows_per_slab=None,
                int_spec=True, device="cpu", **fit_map_kw):
    """Fit every pixel of a dataset FILE (io.open_dataset: HDF5 via h5py, .npy, .npz) without loading it: row slabs are
    read into pinned memory, copied to the device while the previous slab is being fitted, transposed there if the
    file is channel-first, and the integrated spectrum is accumulated on the way.  elements_to_fit defaults to the
    channel names stored in the file.  Returns fit_map's dictionary (+ "int_spec")."""
    from mapstorch_b200 import io as _io

    with _io.open_dataset(file_path, spec_vol_key=spec_vol_key, energy_dim=energy_dim) as handle:
        if elements_to_fit is None:
            elements_to_fit = [e for e in (handle.elems or default_fitting_elems)
                               if e not in ("COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE")]
        vol = handle.volume
        if not isinstance(vol, np.memmap) and isinstance(vol, np.ndarray):
            vol = torch.from_numpy(vol if vol.dtype in (np.float32, np.float16) else vol.astype(np.float32))
        return fit_map(vol, energy_range, elements_to_fit, energy_dim=handle.energy_dim, rows_per_slab=rows_per_slab,
                       int_spec=int_spec, device=device, **fit_map_kw)


def fit_spec(int_spec, energy_range, elements_to_fit=default_fitting_elems, fitting_params=default_fitting_params,
             init_param_vals=default_param_vals, fixed_param_vals={}, indices=None, tune_params=True, init_amp=True,
             use_snip=True, loss="mse", optimizer="adam", n_iter=500, l1_lambda=0.0, progress_bar=True,
             device="cuda", status_updator=None, verbose=False, use_finite_diff=False, finite_diff_epsilon=1e-8,
             e_consts=default_energy_consts, elem_info=default_elem_info, detector_type="Si", escape_factor=0.0):
    """Fit one spectrum (opt.py:165-327).  Returns (tensors, spec_fit, bkg, loss_trace) like the
    reference; amplitudes are log10 in the tensor dict.

    With ``tune_params=False`` the problem is linear in the amplitudes and is solved directly on the
    GPU -- non-negative least squares for the MSE loss (what the reference's Adam loop converges to),
    a shifted right-hand side for the ``l1_lambda`` penalty and iteratively re-weighted solves for
    ``loss="l1"`` (robust.py); the iteration controls (optimizer, n_iter, init_amp) are accepted for
    source compatibility and ignored.  ``loss_trace`` holds the objective in the reference's units."""
    assert loss in ["mse", "l1"], "Loss function not supported"
    assert optimizer in ["adam", "adamw"], "Optimizer not supported"

    def finish(result):
        # the reference ticks status_updator once per Adam iteration (opt.py:319-320); callers size their progress
        # bars with n_iter, so the direct solve reports the same number of ticks when it is done
        if status_updator is not None:
            for _ in range(int(n_iter)):
                status_updator.update()
        return result

    if tune_params:
        from mapstorch_b200.refine import fit_spec_global

        return finish(fit_spec_global(int_spec, energy_range, elements_to_fit, fitting_params, init_param_vals,
                                      fixed_param_vals, indices, use_snip, n_iter, e_consts, elem_info, detector_type,
                                      escape_factor, device, verbose, loss=loss, l1_lambda=l1_lambda))
    _native.require_device()
    spec = torch.as_tensor(np.asarray(int_spec, dtype=np.float32) if not isinstance(int_spec, torch.Tensor)
                           else int_spec.detach().float().cpu().numpy())
    elements = list(dict.fromkeys(list(elements_to_fit)))
    maps, plan = fit_map(spec.reshape(1, -1), energy_range, elements, init_param_vals, fixed_param_vals, indices,
                         use_snip, e_consts, elem_info, detector_type, escape_factor, return_plan=True)
    tensors, _ = create_tensors(plan.names, [], init_param_vals, fixed_param_vals, tune_params=False, device=device)
    amps = torch.stack([maps[n].reshape(()) for n in plan.names])
    cropped = spec[plan.er[0]: plan.er[1] + 1].cuda()
    if use_snip:
        p = plan.params
        bkg = _bkg.snip_bkg(cropped, plan.er, *(torch.tensor(float(p[k])) for k in (
            "ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME")),
            device="cuda")
    else:
        bkg = torch.zeros_like(cropped)
    idx = None if indices is None else torch.as_tensor(np.asarray(indices), device="cuda")
    loss_trace = []
    if loss != "mse" or l1_lambda > 0:
        # other objectives (opt.py:230-236, 372-396): re-weighted / shifted non-negative solves, robust.py
        from mapstorch_b200 import robust

        bm = plan.basis.double() if idx is None else plan.basis.double()[:, idx]
        tm = (cropped - bkg).double() if idx is None else (cropped - bkg).double()[idx]
        c = robust.penalty_constant(tm, loss, l1_lambda, len(plan.names))
        amps = robust.solve_amplitudes(bm, tm, loss, c, loss_trace).float()
    for n, a in zip(plan.names, amps):
        tensors[n] = torch.log10(torch.clamp(a, min=1e-30)).to(device).requires_grad_(True)
    spec_fit = (amps[None, :] @ plan.basis).reshape(-1)
    if not loss_trace:
        resid = spec_fit + bkg - cropped
        if idx is not None:
            resid = resid[idx]
        loss_trace = [float(torch.sum(resid * resid).item())]
    return finish((tensors, spec_fit.cpu().numpy(), bkg.cpu().numpy(), loss_trace))
