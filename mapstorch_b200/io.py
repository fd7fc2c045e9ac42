"""Dataset I/O with the reference's signatures and dictionary keys (io.py:58-120, 123-189).

``read_dataset`` returns ``{"spec_vol", "elems", "int_spec"}``; ``create_dataset`` stores the
volume energy-axis-last under ``MAPS/mca_arr`` with ``MAPS/int_spec`` and ``MAPS/channel_names``.
HDF5 needs ``h5py``, which is imported lazily; paths ending in ``.npz`` use the same keys in a
NumPy archive so the hot path can be exercised where no HDF5 library exists.

``write_override_params_file`` / ``parse_override_params_file`` (io.py:192-459, 462-547) carry
fitted parameters and the element list to and from XRF-Maps' ``maps_fit_parameters_override.txt``.
The file layout is data (``data/override_template.txt``, derived from the reference writer by
``tools/export_override_template.py``); the writer only fills its fields.
"""
import re
import warnings
from datetime import datetime
from pathlib import Path

import numpy as np

from mapstorch_b200.default import default_fitting_elems, default_param_vals, param_name_map

_TEMPLATE_PATH = Path(__file__).resolve().parent / "data" / "override_template.txt"
_FIELD = re.compile(r"<<(@?[A-Z_0-9]+)(?:\|([^>]*))?>>")

_VOL_KEYS = ("MAPS/Spectra/mca_arr", "MAPS/mca_arr")
_ELEM_KEYS = ("MAPS/channel_names", "MAPS/XRF_Analyzed/Fitted/Channel_Names")
_INT_KEYS = ("MAPS/int_spec",)
_SCATTER = ("COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE")


def _open(path, mode):
    if str(path).endswith(".npz"):
        return None
    try:
        import h5py
    except ImportError as exc:  # pragma: no cover - depends on the image
        raise ImportError("reading/writing HDF5 datasets needs h5py; use a .npz path otherwise") from exc
    return h5py.File(path, mode)


def _lookup(store, explicit, fallbacks, what, dtype=None):
    keys = (explicit,) if explicit is not None else fallbacks
    for key in keys:
        try:
            data = store[key][:] if not isinstance(store, np.lib.npyio.NpzFile) else store[key]
            return data.astype(dtype) if dtype is not None else data
        except Exception:
            continue
    if explicit is not None:
        raise KeyError(f"Could not find {what} in the dataset with the given key {explicit}")
    warnings.warn(f"Could not find {what} in the dataset")
    return None


def read_dataset(file_path, spec_vol_key=None, fit_elem_key=None, int_spec_key=None, dtype=np.float32):
    handle = _open(file_path, "r")
    store = handle if handle is not None else np.load(file_path, allow_pickle=False)
    try:
        spec_vol = _lookup(store, spec_vol_key, _VOL_KEYS, "spectra volume", dtype)
        elems = _lookup(store, fit_elem_key, _ELEM_KEYS, "fitting elements")
        int_spec = _lookup(store, int_spec_key, _INT_KEYS, "integrated spectra", dtype)
    finally:
        store.close()
    if elems is not None:
        elems = [e.decode("utf-8") if isinstance(e, bytes) else str(e) for e in elems]
        for name in _SCATTER:
            if name not in elems:
                elems.append(name)
    return {"spec_vol": spec_vol, "elems": elems, "int_spec": int_spec}


class DatasetHandle:
    """A dataset file opened for streaming (``open_dataset``).

    ``volume`` is the spectral volume as a LAZILY read array -- an h5py dataset, a memory-mapped ``.npy`` file, or
    (``.npz``: NumPy archives cannot be mapped) the loaded member -- in the layout it is stored in: ``energy_dim`` = 2
    for files written by ``create_dataset`` (``MAPS/mca_arr``, [H, W, C]), 0 for raw XRF-Maps files
    (``MAPS/Spectra/mca_arr``, [C, H, W]).  ``opt.fit_map(handle.volume, ..., energy_dim=handle.energy_dim)`` or
    ``opt.fit_dataset(path, ...)`` stream it through the GPU slab by slab."""

    def __init__(self, volume, energy_dim, elems, int_spec, closer=None):
        self.volume, self.energy_dim, self.elems, self.int_spec, self._closer = volume, energy_dim, elems, int_spec, closer

    @property
    def shape(self):
        return tuple(int(v) for v in self.volume.shape)

    @property
    def map_shape(self):
        s = self.shape
        return (s[1], s[2]) if self.energy_dim == 0 else (s[0], s[1])

    @property
    def n_channels(self):
        return self.shape[0] if self.energy_dim == 0 else self.shape[2]

    def close(self):
        if self._closer is not None:
            self._closer()
            self._closer = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def open_dataset(file_path, spec_vol_key=None, fit_elem_key=None, int_spec_key=None, energy_dim=None):
    """Open a dataset for streaming: the keys and fallbacks of read_dataset (io.py:63-112), but the volume is not read.
    A bare ``.npy`` file is taken as the volume itself.  energy_dim: axis holding the channels; default 0 when the
    volume is found under ``MAPS/Spectra/mca_arr`` (the raw XRF-Maps layout), 2 otherwise."""
    path = str(file_path)
    if path.endswith(".npy"):
        vol = np.load(path, mmap_mode="r")
        return DatasetHandle(vol, 2 if energy_dim is None else int(energy_dim), None, None)
    handle = _open(path, "r")
    store = handle if handle is not None else np.load(path, allow_pickle=False)
    keys = (spec_vol_key,) if spec_vol_key is not None else _VOL_KEYS
    vol, found = None, None
    for key in keys:
        try:
            vol = store[key]
            found = key
            break
        except Exception:
            continue
    if vol is None:
        store.close()
        raise KeyError(f"Could not find spectra volume in the dataset (keys tried: {', '.join(map(str, keys))})")
    elems = _lookup(store, fit_elem_key, _ELEM_KEYS, "fitting elements") if (fit_elem_key or any(k in store for k in _ELEM_KEYS)) else None
    int_spec = _lookup(store, int_spec_key, _INT_KEYS, "integrated spectra") if (int_spec_key or any(k in store for k in _INT_KEYS)) else None
    if elems is not None:
        elems = [e.decode("utf-8") if isinstance(e, bytes) else str(e) for e in elems]
        for name in _SCATTER:
            if name not in elems:
                elems.append(name)
    if energy_dim is None:
        energy_dim = 0 if str(found).endswith("Spectra/mca_arr") else 2
    return DatasetHandle(vol, int(energy_dim), elems, int_spec, closer=store.close)


def create_dataset(spec_vol, energy_dim, output_path, fit_elems=None, dtype=np.float32, compression="gzip",
                   compression_opts=4):
    """Write a dataset readable by read_dataset; `energy_dim` says which axis holds the channels."""
    if energy_dim not in [0, 1, 2]:
        raise ValueError("energy_dim must be 0, 1, or 2")
    if isinstance(spec_vol, str):
        try:
            spec_vol = np.load(spec_vol)
        except Exception:
            raise ValueError("Could not load .npy file")
    elif not isinstance(spec_vol, np.ndarray) or spec_vol.ndim != 3:
        raise ValueError("spec_vol must be a 3D numpy array")
    if energy_dim != 2:
        spec_vol = np.moveaxis(spec_vol, energy_dim, 2)
    int_spec = np.sum(spec_vol, axis=(0, 1))
    names = None
    if fit_elems is not None:
        names = [e.encode("utf-8") for e in fit_elems]
        names += [n.encode("utf-8") for n in _SCATTER if n not in fit_elems]
    handle = _open(output_path, "w")
    if handle is None:
        payload = {"MAPS/mca_arr": spec_vol.astype(dtype), "MAPS/int_spec": int_spec.astype(dtype)}
        if names is not None:
            payload["MAPS/channel_names"] = np.array(names)
        with open(output_path, "wb") as fh:
            np.savez_compressed(fh, **payload) if compression else np.savez(fh, **payload)
        return output_path
    with handle as f:
        opts = dict(compression=compression, compression_opts=compression_opts) if compression == "gzip" else dict(
            compression=compression)
        f.create_dataset("MAPS/mca_arr", data=spec_vol.astype(dtype), **opts)
        f.create_dataset("MAPS/int_spec", data=int_spec.astype(dtype), **opts)
        if names is not None:
            f.create_dataset("MAPS/channel_names", data=names)
    return output_path


def _literal(text):
    """fallback value of a template field, as the Python literal it was written from"""
    try:
        return int(text)
    except ValueError:
        return float(text)


def write_override_params_file(output_path, param_values=None, elements=None, detector_material="Si",
                               escape_factor=0.0):
    """Write a maps_fit_parameters_override.txt (io.py:192-459).

    `param_values` may use either this package's names or the file's own tags (`param_name_map`);
    anything not given comes from `default_param_vals`.  Values are written with ``str()`` exactly as
    passed, like the reference's f-strings.  Unknown elements / detector materials raise ValueError."""
    if detector_material not in ["Si", "Ge"]:
        raise ValueError(f"Detector material {detector_material} not supported")
    is_si = detector_material == "Si"
    if elements is not None:
        unknown = [e for e in elements if e not in default_fitting_elems]
        if unknown:
            raise ValueError(f"Unsupported elements: {unknown}")
    else:
        elements = default_fitting_elems
    params = default_param_vals.copy()
    if param_values:
        params.update({param_name_map.get(k, k): v for k, v in param_values.items()})
    special = {
        "@DATE": datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f"),
        "@ELEMENTS": ", ".join(elements),
        "@DETECTOR_MATERIAL": 1 if is_si else 0,
        "@SI_ESCAPE_FACTOR": escape_factor if is_si else 0.0,
        "@GE_ESCAPE_FACTOR": 0.0 if is_si else escape_factor,
        "@SI_ESCAPE_ENABLE": 1 if is_si else 0,
        "@GE_ESCAPE_ENABLE": 0 if is_si else 1,
    }

    def fill(match):
        key, fallback = match.group(1), match.group(2)
        if key in special:
            return str(special[key])
        return str(params[key]) if key in params else str(_literal(fallback))

    text = _FIELD.sub(fill, _TEMPLATE_PATH.read_text())
    with open(output_path, "w") as fh:
        fh.write(text)
    return output_path


def parse_override_params_file(file_path):
    """Read a maps_fit_parameters_override.txt back (io.py:462-547).

    Returns the numeric parameters under this package's names plus ``elements_to_fit``,
    ``elements_with_pileup``, ``detector_material`` and the ``escape_factor`` of the enabled
    detector.  A file that cannot be read gives a warning and the defaults, like the reference."""
    tags = {k: k for k in default_param_vals.keys()} | param_name_map
    params = {}
    lists = {"ELEMENTS_TO_FIT": [], "ELEMENTS_WITH_PILEUP": []}
    material = "Si"
    enable = {"SI": 0, "GE": 0}
    factor = {"SI": 0.0, "GE": 0.0}
    try:
        with open(file_path, "r") as fh:
            for raw in fh.readlines():
                line = raw.strip()
                if not line or ":" not in line:
                    continue
                tag, value = (part.strip() for part in line.split(":", 1))
                if tag == "DETECTOR_MATERIAL":
                    material = "Ge" if int(value) == 0 else "Si"
                elif tag in lists:
                    lists[tag] = [e.strip() for e in value.split(",") if e.strip()]
                elif tag in ("SI_ESCAPE_FACTOR", "GE_ESCAPE_FACTOR"):
                    factor[tag[:2]] = float(value)
                elif tag in ("SI_ESCAPE_ENABLE", "GE_ESCAPE_ENABLE"):
                    enable[tag[:2]] = int(value)
                elif tag in tags:
                    try:
                        params[tags[tag]] = float(value)
                    except ValueError:
                        pass
    except Exception as exc:
        warnings.warn(f"Error parsing override parameter file: {exc}")
    params["elements_to_fit"] = lists["ELEMENTS_TO_FIT"]
    params["elements_with_pileup"] = lists["ELEMENTS_WITH_PILEUP"]
    params["detector_material"] = material
    key = material.upper()
    params["escape_factor"] = factor[key] if enable[key] == 1 else 0.0
    return params
