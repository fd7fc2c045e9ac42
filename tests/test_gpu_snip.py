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


def _both_kernels(dev, ch0, n_ch, lohi, n_pass, mode, dtype_out=torch.float32):
    """(strided kernel result, contiguous-run kernel result) for one problem, called through the C ABI."""
    from mapstorch_b200 import _native

    n_px = dev.shape[0]
    width = n_ch * (2 if mode >= _native.MT_SNIP_RESIDUAL_HILO else 1)
    old = torch.full((n_px, width), -7.0, device="cuda").to(dtype_out)
    new = torch.full((n_px, width), -9.0, device="cuda").to(dtype_out)
    flag_old = torch.zeros(1, dtype=torch.int32, device="cuda")
    flag_new = torch.zeros(1, dtype=torch.int32, device="cuda")
    code = _native.dtype_code(dev.dtype)
    if mode == _native.MT_SNIP_RESIDUAL_HILO_F16:
        _native.call("mt_snip_residual_f16", dev.data_ptr(), code, n_px, dev.stride(0), ch0, n_ch, lohi.data_ptr(), n_ch, 0,
                     n_pass, 5, old.data_ptr(), old.stride(0), flag_old.data_ptr(), _native.stream_ptr())
    else:
        _native.call("mt_snip", dev.data_ptr(), code, n_px, dev.stride(0), ch0, n_ch, lohi.data_ptr(), n_pass, 5,
                     old.data_ptr(), old.stride(0), mode, _native.stream_ptr())
    runs = lohi._mt_runs
    assert _native.load().mt_snip_runs_eligible(dev.data_ptr(), code, dev.stride(0), ch0, n_ch, 5, new.data_ptr(),
                                                new.stride(0), mode) == 1
    _native.call("mt_snip_runs", dev.data_ptr(), code, n_px, dev.stride(0), ch0, n_ch, runs.data_ptr(), runs.stride(0),
                 n_pass, 5, new.data_ptr(), new.stride(0), mode, flag_new.data_ptr(), _native.stream_ptr())
    torch.cuda.synchronize()
    assert int(flag_old.item()) == int(flag_new.item())
    return old, new


@pytest.mark.parametrize("n_ch,ch0,n_px,dtype", [
    (4096, 0, 1203, torch.float32), (4096, 0, 37, torch.float16), (3000, 8, 9, torch.float32),
    (2048, 0, 1189, torch.float16), (1400, 48, 6, torch.float32), (1000, 0, 5, torch.float32), (512, 16, 1, torch.float16),
    (8, 0, 3, torch.float32)])
def test_runs_kernel_is_bit_identical_to_the_strided_kernel(snip_golden, n_ch, ch0, n_px, dtype):
    """csrc/mt_snip_runs.cu (contiguous runs, register neighbours, bulk-copy prefetch, persistent CTAs) performs the
    same fp32 operations as csrc/mt_snip.cu: every output mode must agree bit for bit, for every run length, through
    more quads than there are CTAs, with ragged last quads, crops inside wider rows and both storage types."""
    import os

    from mapstorch_b200 import _native, bkg

    if os.environ.get("MT_SNIP_RUNS") == "0":
        pytest.skip("the contiguous-run kernel is switched off (MT_SNIP_RUNS=0)")
    pr = snip_golden["c4096_params"]
    rng = np.random.default_rng(n_ch + n_px)
    ld = ch0 + n_ch + 24
    chan = np.arange(ld)
    peaks = 900 * np.exp(-0.5 * ((chan - 0.4 * ld) / 7.0) ** 2) + 300 * np.exp(-0.5 * ((chan - 0.8 * ld) / 15.0) ** 2)
    spec = rng.poisson(rng.uniform(0, 40, size=(n_px, 1)) + peaks).astype(np.float32)
    spec[0, ch0: ch0 + 3] = 0.0
    dev = torch.from_numpy(spec).cuda().to(dtype)
    lohi, n_pass = bkg.snip_tables_device(n_ch, (ch0, ch0 + n_ch - 1), *[torch.tensor(float(v)) for v in pr])
    runs = lohi._mt_runs.cpu().numpy().view(np.uint32)
    kind, publish = runs[:n_pass, -512:] & 0x1ff, (runs[:n_pass, -512:] >> 16) & 0xff
    if n_ch >= 512:
        assert (kind >= 0x100).any() and (kind == 0).any()  # both the register path and the table path are exercised
        assert (publish == 0).any() and (publish != 0).any()  # records nobody gathers are not published
    for mode in (_native.MT_SNIP_BACKGROUND, _native.MT_SNIP_RESIDUAL, _native.MT_SNIP_RESIDUAL_HILO):
        old, new = _both_kernels(dev, ch0, n_ch, lohi, n_pass, mode)
        assert torch.equal(old, new), (mode, (old != new).sum().item())
    old, new = _both_kernels(dev, ch0, n_ch, lohi, n_pass, _native.MT_SNIP_RESIDUAL_HILO_F16, torch.float16)
    assert torch.equal(old, new)


def test_runs_kernel_overflow_flag_and_unaligned_problems_fall_back(snip_golden):
    import os

    from mapstorch_b200 import _native, bkg

    if os.environ.get("MT_SNIP_RUNS") == "0":
        pytest.skip("the contiguous-run kernel is switched off (MT_SNIP_RUNS=0)")

    pr = snip_golden["c2048_params"]
    lohi, n_pass = bkg.snip_tables_device(2048, (0, 2047), *[torch.tensor(float(v)) for v in pr])
    spec = torch.zeros((5, 2048), device="cuda")
    spec[3, 1000] = 3.0e5  # residual beyond the fp16 range
    old, new = _both_kernels(spec, 0, 2048, lohi, n_pass, _native.MT_SNIP_RESIDUAL_HILO_F16, torch.float16)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.empty((5, 4096), dtype=torch.float16, device="cuda")
    bkg.snip_residual_f16(spec, 0, 2048, lohi, n_pass, 5, out, flag)
    assert int(flag.item()) == 1
    # a crop that starts on an odd channel is not served by the run kernel: the entry point refuses, the wrapper falls back
    wide = torch.rand((4, 2100), device="cuda") * 50
    lohi2, n_pass2 = bkg.snip_tables_device(2048, (3, 2050), *[torch.tensor(float(v)) for v in pr])
    out2 = torch.empty((4, 2048), device="cuda")
    assert _native.load().mt_snip_runs_eligible(wide.data_ptr(), _native.MT_DTYPE_F32, wide.stride(0), 3, 2048, 5,
                                                out2.data_ptr(), out2.stride(0), 0) == 0
    with pytest.raises(_native.NativeError):
        _native.call("mt_snip_runs", wide.data_ptr(), _native.MT_DTYPE_F32, 4, wide.stride(0), 3, 2048,
                     lohi2._mt_runs.data_ptr(), lohi2._mt_runs.stride(0), n_pass2, 5, out2.data_ptr(), out2.stride(0), 0, None,
                     _native.stream_ptr())
    bkg.snip_launch(wide, 3, 2048, lohi2, n_pass2, 5, out2, _native.MT_SNIP_BACKGROUND)
    ref = torch.empty((4, 2048), device="cuda")
    _native.call("mt_snip", wide.data_ptr(), _native.MT_DTYPE_F32, 4, wide.stride(0), 3, 2048, lohi2.data_ptr(), n_pass2, 5,
                 ref.data_ptr(), ref.stride(0), 0, _native.stream_ptr())
    assert torch.equal(out2, ref)


def test_device_built_gather_tables_are_bit_identical_to_the_host_tables():
    """mt_snip_tables_f32 (the table builder fit_spec(tune_params=True) uses at every trial point) against
    tables.snip_tables, itself pinned to the reference's torch ops: 150 random calibrations, crops and sizes."""
    import torch

    from mapstorch_b200 import bkg, tables

    rng = np.random.default_rng(7)
    for trial in range(150):
        n_ch = int(rng.integers(8, 4096)) if trial % 10 else int(rng.choice([1, 2, 5, 16384]))
        ch0 = int(rng.integers(0, 300))
        vals = (rng.normal(0, 0.05), rng.uniform(0.003, 0.03), rng.normal(0, 1e-7) if trial % 3 else 0.0,
                rng.uniform(0.05, 4.0), rng.uniform(0.02, 0.4), rng.uniform(1e-6, 2e-3))
        six = [torch.tensor(float(v)) for v in vals]
        want = tables.snip_tables(n_ch, ch0, *six)
        lohi, n_pass = bkg.snip_tables_device(n_ch, (ch0, ch0 + n_ch - 1), *six, aux=False)
        assert n_pass == want.shape[0]
        got = lohi.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), f"trial {trial}: {np.argwhere(got != want)[:4]}"
    # Python-float operands keep the host route (their arithmetic is partly double precision in the reference)
    floats = [float(v) for v in vals]
    lohi, n_pass = bkg.snip_tables_device(n_ch, (ch0, ch0 + n_ch - 1), *floats, aux=False)
    assert np.array_equal(lohi.cpu().numpy().view(np.uint32), tables.snip_tables(n_ch, ch0, *floats))
