"""K2 (mt_snip) on the GPU vs reference goldens and the oracle.  Indices are host tables
(bit-exact by test_tables); values must be within 1e-5 of the row maximum."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["c2048", "c1401", "c4096"])
def test_snip_bkg_matches_reference(snip_golden, tag):
    from mapstorch_b200 import bkg

    g = snip_golden
    er, spec, ref, pr = tuple(int(v) for v in g[f"{tag}_er"]), g[f"{tag}_spec"], g[f"{tag}_bkg"], g[f"{tag}_params"]
    args = [torch.tensor(float(v)) for v in pr]
    out = bkg.snip_bkg(torch.from_numpy(spec), er, *args, device="cpu").numpy()
    assert out.shape == ref.shape and out.dtype == np.float32
    for row in range(spec.shape[0]):
        scale = np.abs(ref[row]).max()
        assert np.abs(out[row] - ref[row]).max() <= 1e-5 * max(scale, 1e-30), (tag, row)
    # 1-D and [H, W, C] inputs, fp16 storage (exact for these counts)
    one = bkg.snip_bkg(torch.from_numpy(spec[1]), er, *args, device="cpu").numpy()
    assert np.array_equal(one, out[1])
    vol = bkg.snip_bkg(torch.from_numpy(spec.reshape(2, 3, -1)), er, *args, device="cuda")
    assert vol.is_cuda and np.array_equal(vol.cpu().numpy().reshape(6, -1), out)
    small = spec[:3]
    assert small.max() < 2048
    half = bkg.snip_bkg(torch.from_numpy(small).half(), er, *args, device="cpu").numpy()
    assert np.array_equal(half, out[:3])


@pytest.mark.parametrize("n_px", [1, 3, 4, 5, 1000])
def test_ragged_batches_and_modes(snip_golden, n_px):
    from mapstorch_b200 import _native, bkg
    from oracle import maps_oracle as mo

    g = snip_golden
    er, pr = (0, 2047), g["c2048_params"]
    rng = np.random.default_rng(n_px)
    spec = rng.poisson(rng.uniform(0, 30, size=(n_px, 1)) + 400 * np.exp(-0.5 * ((np.arange(2048) - 700) / 9.0) ** 2),
                       ).astype(np.float32)
    ref = mo.snip_bkg(spec, er, *pr)
    dev = torch.from_numpy(spec).cuda()
    lohi, n_pass = bkg.snip_tables_device(2048, er, *[torch.tensor(float(v)) for v in pr])
    out = torch.empty((n_px, 2048), device="cuda")
    bkg.snip_launch(dev, 0, 2048, lohi, n_pass, 5, out, _native.MT_SNIP_BACKGROUND)
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-5 * ref.max()
    res = torch.empty((n_px, 2048), device="cuda")
    bkg.snip_launch(dev, 0, 2048, lohi, n_pass, 5, res, _native.MT_SNIP_RESIDUAL)
    assert np.array_equal(res.cpu().numpy(), spec - got)
    hilo = torch.empty((n_px, 4096), device="cuda")
    bkg.snip_launch(dev, 0, 2048, lohi, n_pass, 5, hilo, _native.MT_SNIP_RESIDUAL_HILO)
    h = hilo.cpu().numpy()
    assert np.array_equal(h[:, :2048] + h[:, 2048:], spec - got)
    assert not np.any(h[:, :2048].view(np.uint32) & 0x1FFF)


def test_cropped_window_inside_wider_rows(snip_golden):
    """energy_range crop: the kernel reads channels [ch0, ch0+n_ch) of rows with a larger pitch."""
    from mapstorch_b200 import _native, bkg

    g = snip_golden
    er, pr = (50, 1450), g["c1401_params"]
    full = np.zeros((6, 2048), dtype=np.float32)
    full[:, 50:1451] = g["c1401_spec"]
    full[:, :50] = 7.0
    lohi, n_pass = bkg.snip_tables_device(1401, er, *[torch.tensor(float(v)) for v in pr])
    out = torch.empty((6, 1401), device="cuda")
    bkg.snip_launch(torch.from_numpy(full).cuda(), 50, 1401, lohi, n_pass, 5, out, _native.MT_SNIP_BACKGROUND)
    ref = g["c1401_bkg"]
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5 * ref.max()


def test_snip_op_single_pass():
    from mapstorch_b200 import bkg

    bg = torch.rand(5, 300)
    w = torch.linspace(0.4, 9.0, 300)
    out = bkg.snip_op(bg, w, 0, 299, "cpu")
    idx = torch.arange(300, dtype=torch.float32)
    lo = torch.clamp(idx - w, 0, 299).long().expand(5, 300)
    hi = torch.clamp(idx + w, 0, 299).long().expand(5, 300)
    assert torch.equal(out, torch.minimum((bg.gather(-1, lo) + bg.gather(-1, hi)) / 2, bg))
