"""The C-ABI shared library: it loads, exports every symbol include/mapstorch_b200.h declares,
and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from mapstorch_b200 import _native

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "mapstorch_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = declared_symbols()
    assert "mt_fit_contract" in names and "mt_snip" in names and len(names) >= 9
    for name in names:
        assert hasattr(lib, name), name
        assert name in _native.SIGNATURES or name == "mt_last_error", f"{name} has no ctypes signature"
    assert lib.mt_abi_version() == 1


def test_term_struct_layout_matches_header():
    assert ctypes.sizeof(_native.MtTerm) == 32
    from mapstorch_b200.tables import TERM_DTYPE

    assert TERM_DTYPE.itemsize == 32
    assert [n for n, _ in _native.MtTerm._fields_] == list(TERM_DTYPE.names)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(_native.NativeError):
        _native.require_device()
    from mapstorch_b200 import bkg, map as mmap, opt
    from mapstorch_b200.default import default_param_vals

    with pytest.raises(_native.NativeError):
        mmap.model_spec(default_param_vals | {"Fe": 1.0, "COMPTON_AMPLITUDE": 0.0, "COHERENT_SCT_AMPLITUDE": 0.0},
                        (0, 2047), ["Fe"])
    with pytest.raises(_native.NativeError):
        bkg.snip_bkg(torch.zeros(4, 64), (0, 63), 0.0, 0.012, 0.0, 0.5, 0.08, 1e-4)
    with pytest.raises(_native.NativeError):
        opt.fit_map(torch.zeros(2, 2, 2048), (0, 2047), ["Fe"])
    lib = _native.load()
    assert lib.mt_device_check() == -4 and b"no CUDA device" in lib.mt_last_error()


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "mapstorch_b200").glob("*.py"):
        src = path.read_text()
        assert "import oracle" not in src and "from oracle" not in src, path
