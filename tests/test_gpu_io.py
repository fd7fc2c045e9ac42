"""The data formats either side of the fit on the GPU: channel-first volumes (the raw XRF-Maps
layout), datasets streamed from a file slab by slab, and the integrated spectrum as a by-product of the pass.  Pixels
are independent and the integrated spectrum of integer counts is exact in fp64, so every variant must reproduce the
in-memory energy-last fit BIT FOR BIT."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ELEMS10 = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]


def _volume(h=9, w=13, n_ch=2048, seed=3, flat=2.0):
    from mapstorch_b200.default import default_param_vals
    from oracle import maps_oracle as mo

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    rng = np.random.default_rng(seed)
    B, names = mo.basis(p, (0, n_ch - 1), ELEMS10)
    amps = 10.0 ** rng.uniform(1.5, 3.0, size=(h * w, B.shape[1]))
    y = rng.poisson(amps @ B.astype(np.float64).T + flat).astype(np.float32).reshape(h, w, n_ch)
    return p, y, names


@pytest.mark.parametrize("use_snip,dtype", [(False, np.float32), (True, np.float32), (False, np.float16)])
def test_channel_first_volumes_equal_the_energy_last_fit(use_snip, dtype):
    from mapstorch_b200 import opt

    p, y, names = _volume()
    y = y.astype(dtype)
    er = (0, 2047)
    kw = dict(init_param_vals=p, use_snip=use_snip, int_spec=True)
    ref = opt.fit_map(torch.from_numpy(y).cuda(), er, ELEMS10, **kw)
    cf = np.ascontiguousarray(np.moveaxis(y, 2, 0))  # [C, H, W]
    # host ndarray, CUDA tensor, small slabs (ragged last slab), one slab
    variants = {"host": opt.fit_map(cf, er, ELEMS10, energy_dim=0, rows_per_slab=2, **kw),
                "cuda": opt.fit_map(torch.from_numpy(cf).cuda(), er, ELEMS10, energy_dim=0, rows_per_slab=4, **kw),
                "whole": opt.fit_map(torch.from_numpy(cf), er, ELEMS10, energy_dim=0, rows_per_slab=100, **kw)}
    exact = y.astype(np.float64).sum(axis=(0, 1))
    assert np.array_equal(ref["int_spec"].cpu().numpy(), exact.astype(np.float32))
    for tag, got in variants.items():
        assert set(got) == set(ref)
        for k in ref:
            assert got[k].shape == ref[k].shape and torch.equal(got[k].cuda(), ref[k].cuda()), (tag, k)
    # [H, C, W] (energy_dim=1) is converted like create_dataset does
    mid = opt.fit_map(np.ascontiguousarray(np.moveaxis(y, 2, 1)), er, ELEMS10, energy_dim=1, **kw)
    for k in ref:
        assert torch.equal(mid[k].cuda(), ref[k].cuda()), k


def test_dataset_files_are_streamed_slab_by_slab(tmp_path):
    from mapstorch_b200 import io as mio, opt

    p, y, names = _volume(h=10, w=8)
    er = (50, 1450)
    ref = opt.fit_map(torch.from_numpy(y).cuda(), er, ELEMS10, init_param_vals=p, use_snip=True, device="cpu")
    exact = y.astype(np.float64).sum(axis=(0, 1)).astype(np.float32)
    # (1) a dataset written by create_dataset (energy last), read back through open_dataset / fit_dataset
    path = tmp_path / "scan.npz"
    mio.create_dataset(y, 2, str(path), fit_elems=ELEMS10, compression=None)
    got = opt.fit_dataset(str(path), er, init_param_vals=p, use_snip=True, rows_per_slab=3)
    assert np.array_equal(got["int_spec"].numpy(), exact)
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    # (2) a raw channel-first volume in a memory-mapped .npy file: read rows_per_slab rows at a time
    raw = tmp_path / "mca_arr.npy"
    np.save(raw, np.ascontiguousarray(np.moveaxis(y, 2, 0)))
    with mio.open_dataset(str(raw), energy_dim=0) as handle:
        assert isinstance(handle.volume, np.memmap) and handle.map_shape == (10, 8) and handle.n_channels == 2048
        got2 = opt.fit_map(handle.volume, er, ELEMS10, init_param_vals=p, use_snip=True, energy_dim=0, rows_per_slab=4,
                           int_spec=True, device="cpu")
    assert np.array_equal(got2["int_spec"].numpy(), exact)
    for k in ref:
        assert torch.equal(got2[k], ref[k]), k
    # (3) energy-last memory map (lazy rows)
    last = tmp_path / "vol.npy"
    np.save(last, y)
    got3 = opt.fit_dataset(str(last), er, ELEMS10, init_param_vals=p, use_snip=True, rows_per_slab=7)
    for k in ref:
        assert torch.equal(got3[k], ref[k]), k
    # per-pixel refinement runs on the streamed slabs as well
    fine = opt.fit_dataset(str(raw), er, ELEMS10, energy_dim=0, init_param_vals=p, use_snip=False, rows_per_slab=5,
                           refine=("ENERGY_OFFSET",), refine_iters=6)
    fine_ref = opt.fit_map(torch.from_numpy(y), er, ELEMS10, init_param_vals=p, use_snip=False, refine=("ENERGY_OFFSET",),
                           refine_iters=6, device="cpu")
    for k in fine_ref:
        assert torch.equal(fine[k], fine_ref[k]), k


def test_hdf5_round_trip_when_h5py_is_available(tmp_path):
    """The HDF5 branch of read_dataset / create_dataset / open_dataset (io.py:58-189) -- needs h5py, which this image
    does not ship; the test runs wherever it does."""
    h5py = pytest.importorskip("h5py")
    from mapstorch_b200 import io as mio, opt

    p, y, names = _volume(h=6, w=5)
    path = tmp_path / "scan.h5"
    mio.create_dataset(y, 2, str(path), fit_elems=ELEMS10)
    back = mio.read_dataset(str(path))
    assert np.array_equal(back["spec_vol"], y) and back["elems"][:10] == ELEMS10
    with h5py.File(path, "r") as f:
        assert f["MAPS/mca_arr"].compression == "gzip" and f["MAPS/int_spec"].shape == (2048,)
    # raw XRF-Maps layout: channel-first under MAPS/Spectra/mca_arr
    raw = tmp_path / "raw.h5"
    with h5py.File(raw, "w") as f:
        f.create_dataset("MAPS/Spectra/mca_arr", data=np.moveaxis(y, 2, 0))
    ref = opt.fit_map(torch.from_numpy(y), (0, 2047), ELEMS10, init_param_vals=p, use_snip=False, device="cpu")
    got = opt.fit_dataset(str(raw), (0, 2047), ELEMS10, init_param_vals=p, use_snip=False, rows_per_slab=2, int_spec=False)
    for k in ref:
        assert torch.equal(got[k], ref[k]), k


def test_l1_maps_from_channel_first_and_memory_mapped_volumes(tmp_path):
    """fit_map(loss="l1") on streamed sources -- a channel-first volume (energy_dim=0) in host memory and a memory-mapped
    energy-last file -- equals the fit of the same volume resident in HBM, bit for bit (pixels are independent; the
    slabs only change where a pixel sits inside a launch), with and without the amplitude penalty."""
    from mapstorch_b200 import opt

    p, y, names = _volume(h=6, w=8)
    er = (50, 1450)
    for lam in (0.0, 1e-3):
        kw = dict(init_param_vals=p, use_snip=True, loss="l1", l1_lambda=lam)
        ref = opt.fit_map(torch.from_numpy(y).cuda(), er, ELEMS10, **kw)
        cf = np.ascontiguousarray(np.moveaxis(y, 2, 0))
        got = opt.fit_map(cf, er, ELEMS10, energy_dim=0, rows_per_slab=4, **kw)
        raw = tmp_path / f"vol_{lam}.npy"
        np.save(raw, y)
        lazy = opt.fit_map(np.load(raw, mmap_mode="r"), er, ELEMS10, rows_per_slab=5, **kw)
        for k in ref:
            assert torch.equal(got[k].cuda(), ref[k].cuda()), ("channel-first", lam, k)
            assert torch.equal(lazy[k].cuda(), ref[k].cuda()), ("memmap", lam, k)
