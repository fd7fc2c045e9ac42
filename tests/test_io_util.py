"""Boundary I/O and helper functions that run on the host: dataset round trip (io.py:58-189 semantics,
.npz container where no HDF5 library exists) and the scatter-peak heuristic (util.py:189-229)."""
import numpy as np
import pytest

from mapstorch_b200 import io, util
from mapstorch_b200.default import default_param_vals


def test_dataset_round_trip_npz(tmp_path):
    rng = np.random.default_rng(0)
    vol = rng.poisson(3.0, size=(16, 5, 7)).astype(np.float32)  # energy axis first, as XRF-Maps stores it
    path = str(tmp_path / "scan.npz")
    assert io.create_dataset(vol, 0, path, fit_elems=["Fe", "Cu"], dtype=np.float16) == path
    data = io.read_dataset(path)
    assert set(data) == {"spec_vol", "elems", "int_spec"}
    assert data["spec_vol"].shape == (5, 7, 16) and data["spec_vol"].dtype == np.float32
    assert np.array_equal(data["spec_vol"], np.moveaxis(vol, 0, 2))
    assert np.array_equal(data["int_spec"], vol.sum(axis=(1, 2)))
    assert data["elems"] == ["Fe", "Cu", "COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE"]
    with pytest.raises(KeyError):
        io.read_dataset(path, spec_vol_key="MAPS/does_not_exist")
    with pytest.raises(ValueError):
        io.create_dataset(vol, 3, path)
    with pytest.raises(ValueError):
        io.create_dataset(vol[0], 0, path)


def test_dataset_without_element_names(tmp_path):
    path = str(tmp_path / "scan.npz")
    io.create_dataset(np.ones((2, 3, 8), dtype=np.float32), 2, path)
    with pytest.warns(UserWarning):
        data = io.read_dataset(path)
    assert data["elems"] is None and data["spec_vol"].shape == (2, 3, 8)


def test_scatter_peak_heuristic_recovers_slope_and_angle():
    from oracle import maps_oracle as mo

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0, COMPTON_AMPLITUDE=4.0, COHERENT_SCT_AMPLITUDE=4.5, Fe=3.0, Cu=3.0)
    spec = mo.model_spec(p, (0, 2047), ["Fe", "Cu"]) + 1.0
    upd = util.estimate_and_update_params(spec, (50, 1450), 12.0)
    assert set(upd) == {"COHERENT_SCT_ENERGY", "ENERGY_SLOPE", "COMPTON_ANGLE"}
    assert abs(upd["ENERGY_SLOPE"] / p["ENERGY_SLOPE"] - 1) < 5e-3
    assert abs(upd["COMPTON_ANGLE"] - p["COMPTON_ANGLE"]) < 8.0
    assert util.estimate_and_update_params(spec, (50, 1450), None) == {}
    assert util.estimate_and_update_params([], (0, 10), 12.0) == {}
    assert util.get_channel_from_energy(6.4, np.linspace(0, 12, 1201)) == 640
    assert util.get_channel_from_energy(99.0, np.linspace(0, 12, 1201)) == 1200


def test_open_dataset_is_lazy_and_knows_the_layout(tmp_path):
    from mapstorch_b200 import io as mio

    vol = np.random.default_rng(0).poisson(5, size=(6, 7, 64)).astype(np.float32)
    path = tmp_path / "a.npz"
    mio.create_dataset(vol, 2, str(path), fit_elems=["Fe", "Cu"])
    with mio.open_dataset(str(path)) as h:
        assert h.shape == (6, 7, 64) and h.energy_dim == 2 and h.map_shape == (6, 7) and h.n_channels == 64
        assert h.elems == ["Fe", "Cu", "COHERENT_SCT_AMPLITUDE", "COMPTON_AMPLITUDE"]
        assert np.allclose(h.int_spec, vol.sum(axis=(0, 1)))
    raw = tmp_path / "b.npy"
    np.save(raw, np.moveaxis(vol, 2, 0))
    with mio.open_dataset(str(raw), energy_dim=0) as h:
        assert isinstance(h.volume, np.memmap) and h.map_shape == (6, 7) and h.n_channels == 64
        assert np.array_equal(np.asarray(h.volume[:, 2:4, :]), np.moveaxis(vol, 2, 0)[:, 2:4, :])
    with pytest.raises(KeyError):
        mio.open_dataset(str(path), spec_vol_key="MAPS/none")
