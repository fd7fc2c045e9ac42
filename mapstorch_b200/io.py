"""Dataset I/O with the reference's signatures and dictionary keys (io.py:58-120, 123-189).

``read_dataset`` returns ``{"spec_vol", "elems", "int_spec"}``; ``create_dataset`` stores the
volume energy-axis-last under ``MAPS/mca_arr`` with ``MAPS/int_spec`` and ``MAPS/channel_names``.
HDF5 needs ``h5py``, which is imported lazily; paths ending in ``.npz`` use the same keys in a
NumPy archive so the hot path can be exercised where no HDF5 library exists.  The MAPS
override-file writer/parser is out of scope.
"""
import warnings

import numpy as np

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
