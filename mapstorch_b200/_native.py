"""ctypes binding of ``libmapstorch_b200.so`` (the C ABI in ``include/mapstorch_b200.h``).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
called without an sm_100 device, a :class:`NativeError` is raised.
"""
import ctypes
import os
from pathlib import Path

MT_OK = 0
MT_ERR_INVALID, MT_ERR_CUDA, MT_ERR_UNSUPPORTED, MT_ERR_NO_DEVICE = -1, -2, -3, -4  # include/mapstorch_b200.h
MT_DTYPE_F32, MT_DTYPE_F16 = 0, 1
MT_TERM_GAUSS, MT_TERM_STEP, MT_TERM_TAIL = 0, 1, 2
MT_SNIP_BACKGROUND, MT_SNIP_RESIDUAL, MT_SNIP_RESIDUAL_HILO, MT_SNIP_RESIDUAL_HILO_F16 = 0, 1, 2, 3
MT_X_ELEMENT_MAJOR, MT_X_PIXEL_MAJOR = 0, 1

_LIB_NAME = "libmapstorch_b200.so"
_lib = None


class NativeError(RuntimeError):
    """A C-ABI call returned a non-zero status (or the library could not be loaded)."""

    def __init__(self, code, message):
        super().__init__(f"mapstorch_b200 native error {code}: {message}")
        self.code = code


class MtTerm(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("sign", ctypes.c_float), ("e", ctypes.c_float),
                ("p0", ctypes.c_float), ("p1", ctypes.c_float), ("p2", ctypes.c_float),
                ("p3", ctypes.c_float), ("amp", ctypes.c_float)]


MT_MAX_THETA, MT_MAX_THETA_REFINE = 8, 4
MT_SPEC_MAX_COMP, MT_SPEC_MAX_CH = 96, 8192
MT_LINE_ELEMENT, MT_LINE_ELASTIC, MT_LINE_COMPTON = 0, 1, 2


class MtRefineLine(ctypes.Structure):
    _fields_ = [("energy", ctypes.c_float), ("factor", ctypes.c_float), ("comp", ctypes.c_int32),
                ("sigma_scale", ctypes.c_float), ("e296", ctypes.c_float), ("step", ctypes.c_float),
                ("kind", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class MtRefineConfig(ctypes.Structure):
    _fields_ = [("n_ch", ctypes.c_int32), ("ch0", ctypes.c_int32), ("n_comp", ctypes.c_int32),
                ("n_lines", ctypes.c_int32), ("n_theta", ctypes.c_int32), ("max_iter", ctypes.c_int32),
                ("theta_id", ctypes.c_int32 * MT_MAX_THETA), ("theta0", ctypes.c_float * MT_MAX_THETA),
                ("theta_lo", ctypes.c_float * MT_MAX_THETA), ("theta_hi", ctypes.c_float * MT_MAX_THETA),
                ("energy_offset", ctypes.c_float), ("energy_slope", ctypes.c_float),
                ("energy_quad", ctypes.c_float), ("fwhm_offset", ctypes.c_float), ("fwhm_fanoprime", ctypes.c_float),
                ("coherent_sct_energy", ctypes.c_float), ("compton_angle", ctypes.c_float),
                ("compton_fwhm_corr", ctypes.c_float), ("tol", ctypes.c_float)]


class MtPeers(ctypes.Structure):
    _fields_ = [("n_peers", ctypes.c_int32), ("n_owned", ctypes.c_int32), ("peer", ctypes.c_void_p * 8),
                ("multicast", ctypes.c_void_p), ("owner_row", ctypes.c_void_p * 32)]


class MtPeerSync(ctypes.Structure):
    _fields_ = [("n_ranks", ctypes.c_int32), ("rank", ctypes.c_int32), ("flags", ctypes.c_void_p * 8)]


# the calibration parameters with a non-zero gradient on model_spec's path (include/mapstorch_b200.h MT_THETA_*)
THETA_IDS = {"ENERGY_OFFSET": 0, "ENERGY_SLOPE": 1, "FWHM_OFFSET": 2, "FWHM_FANOPRIME": 3, "ENERGY_QUADRATIC": 4,
             "COHERENT_SCT_ENERGY": 5, "COMPTON_ANGLE": 6, "COMPTON_FWHM_CORR": 7}

_vp, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double

# name -> argument types; every function returns int except mt_last_error
SIGNATURES = {
    "mt_abi_version": [],
    "mt_device_check": [],
    "mt_energy_axis": [_i32, _i32, _f32, _f32, _f32, _vp, _vp],
    "mt_model_terms": [_vp, _i32, _vp, _vp, _i32, _vp, _i64, _vp],
    "mt_model_sum": [_vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp],
    "mt_snip": [_vp, _i32, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _vp, _i64, _i32, _vp],
    "mt_snip_tables_f32": [_i32, _i32, _f32, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _vp],
    "mt_snip_scaled_pitch": [_i32],
    "mt_snip_scale_tables": [_vp, _i32, _i32, _vp, _i32, _vp],
    "mt_snip_scaled": [_vp, _i32, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _i32, _vp],
    "mt_snip_runs_pitch": [_i32],
    "mt_snip_runs_tables": [_vp, _i32, _i32, _vp, _i32, _vp],
    "mt_snip_runs_eligible": [_vp, _i32, _i64, _i32, _i32, _i32, _vp, _i64, _i32],
    "mt_snip_runs": [_vp, _i32, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _i32, _vp, _vp],
    "mt_snip_residual_f16": [_vp, _i32, _i64, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp],
    "mt_fit_contract": [_vp, _i32, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _f32, _vp,
                        _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    "mt_nnls_list": [_vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp],
    "mt_fit_contract_simt": [_vp, _i32, _i64, _i64, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp],
    "mt_nnls_fixup": [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp],
    "mt_penalty_shift": [_vp, _i32, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "mt_split_hilo": [_vp, _i64, _i64, _i32, _i32, _vp, _i64, _vp],
    "mt_operand_inexact": [_vp, _i64, _i64, _i32, _i32, _vp, _vp],
    "mt_nnls_singular_count": [_vp, _i32],
    "mt_peer_signal": [_vp, _vp, _vp],
    "mt_peer_wait": [_vp, _vp, _i32, _vp, _vp],
    "mt_transpose_cf": [_vp, _i32, _i32, _i64, _i64, _vp, _i64, _vp],
    "mt_int_spec": [_vp, _i32, _i64, _i64, _i32, _vp, _vp],
    "mt_lad_update": [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _vp],
    "mt_model_jacobian": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "mt_refine": [_vp, _i32, _i64, _i64, _vp, _i64, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64,
                  _vp, _vp, _vp],
    "mt_spec_solve": [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _vp, _i32, _vp, _i64, _f64, _i32, _i32, _vp, _vp, _vp, _vp,
                      _vp, _i32, _vp],
}


def library_path():
    override = os.environ.get("MAPSTORCH_B200_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / _LIB_NAME


def load():
    """Load (once) and return the ctypes handle of the native library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise NativeError(-100, f"{path} not found; build it with `make -C mapstorch_b200/csrc` "
                                "(or python -c 'import __graft_entry__ as g; g.build()'). "
                                "There is no CPU fallback.")
    lib = ctypes.CDLL(str(path))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise NativeError(-101, f"{path} does not export {name}")
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.mt_last_error.argtypes = []
    lib.mt_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def check(status):
    if status != MT_OK:
        msg = load().mt_last_error()
        raise NativeError(status, msg.decode("utf-8", "replace") if msg else "unknown error")


# kernels launched per successful call of each entry point (bench.py reports the total)
KERNELS_PER_CALL = {"mt_model_terms": 1, "mt_model_sum": 1, "mt_snip": 1, "mt_fit_contract": 1,
                    "mt_fit_contract_simt": 1, "mt_nnls_fixup": 1, "mt_nnls_list": 2, "mt_refine": 1,
                    "mt_penalty_shift": 1, "mt_snip_scaled": 1, "mt_snip_scale_tables": 1, "mt_split_hilo": 1, "mt_operand_inexact": 1, "mt_peer_signal": 1, "mt_lad_update": 1, "mt_transpose_cf": 1, "mt_int_spec": 1, "mt_model_jacobian": 1, "mt_snip_residual_f16": 1, "mt_snip_runs": 1, "mt_snip_runs_tables": 2,
                    "mt_peer_wait": 1, "mt_spec_solve": 1, "mt_snip_tables_f32": 1, "mt_energy_axis": 1}
launch_count = 0


def call(name, *args):
    global launch_count
    check(getattr(load(), name)(*args))
    launch_count += KERNELS_PER_CALL.get(name, 0)


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream) as a void*."""
    import torch

    if stream is None:
        # the raw handle of the current stream without building a torch.cuda.Stream object (a third of the cost; this
        # runs three times per trial point of fit_spec(tune_params=True))
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    return ctypes.c_void_p(stream.cuda_stream)


def dtype_code(torch_dtype):
    import torch

    if torch_dtype == torch.float32:
        return MT_DTYPE_F32
    if torch_dtype == torch.float16:
        return MT_DTYPE_F16
    raise NativeError(MT_ERR_UNSUPPORTED, f"spectral volumes must be float32 or float16 on the device, got {torch_dtype}")


_checked_devices = set()


def require_device():
    """Raise unless the current CUDA device is an sm_100 part (no CPU fallback).  A device that passed once is not asked
    again (its compute capability cannot change); the C entry points repeat the check on every call anyway."""
    import torch

    if _checked_devices:
        dev = torch.cuda.current_device()
        if dev in _checked_devices:
            return
    if not torch.cuda.is_available():
        raise NativeError(-4, "no CUDA device is visible; mapstorch_b200 has no CPU path")
    call("mt_device_check")
    _checked_devices.add(torch.cuda.current_device())
