"""Error behaviour of the C ABI and its Python callers ON a B200: bad arguments come back as MT_ERR_*
codes with a message (never a crash, never a silent CPU fallback), and the device stays usable."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _snip_args(n_px=4, n_ch=256):
    from mapstorch_b200 import bkg

    spec = torch.ones((n_px, n_ch), device="cuda")
    lohi, n_pass = bkg.snip_tables_device(n_ch, (0, n_ch - 1), *[torch.tensor(v) for v in
                                                              (0.0, 0.01, 0.0, 0.5, 0.08, 0.00018)])
    out = torch.empty((n_px, n_ch), device="cuda")
    return spec, lohi, n_pass, out


def test_snip_rejects_bad_arguments_and_keeps_working():
    from mapstorch_b200 import _native, bkg

    spec, lohi, n_pass, out = _snip_args()
    with pytest.raises(_native.NativeError) as err:
        bkg.snip_launch(spec, 0, 256, lohi, n_pass, 4, out, _native.MT_SNIP_BACKGROUND)  # even boxcar
    assert err.value.code == _native.MT_ERR_INVALID and "boxcar" in str(err.value)
    with pytest.raises(_native.NativeError) as err:
        bkg.snip_launch(spec, 8, 256, lohi, n_pass, 5, out, _native.MT_SNIP_BACKGROUND)  # window past the row
    assert err.value.code == _native.MT_ERR_INVALID
    with pytest.raises(_native.NativeError) as err:
        bkg.snip_launch(spec, 0, 256, lohi, n_pass, 5, out, 7)                           # unknown output mode
    assert err.value.code == _native.MT_ERR_INVALID and "out_mode" in str(err.value)
    with pytest.raises(_native.NativeError) as err:
        bkg.snip_launch(spec, 0, 256, lohi, n_pass, 5, out, _native.MT_SNIP_RESIDUAL_HILO)  # out too narrow for 2 planes
    assert err.value.code == _native.MT_ERR_INVALID
    with pytest.raises(_native.NativeError) as err:
        _native.call("mt_snip", None, 0, 4, 256, 0, 256, lohi.data_ptr(), n_pass, 5, out.data_ptr(), 256, 0, None)
    assert err.value.code == _native.MT_ERR_INVALID and "null" in str(err.value)
    bad = torch.ones((4, 256), device="cuda", dtype=torch.float64)
    with pytest.raises((KeyError, ValueError, _native.NativeError)):
        bkg.snip_launch(bad, 0, 256, lohi, n_pass, 5, out, _native.MT_SNIP_BACKGROUND)   # fp64 storage is not a thing
    # the library is still fine afterwards
    bkg.snip_launch(spec, 0, 256, lohi, n_pass, 5, out, _native.MT_SNIP_BACKGROUND)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()


def test_snip_channel_limit_is_reported_not_truncated():
    from mapstorch_b200 import _native, bkg

    n_ch = 512 * 16 + 8  # one step past what a CTA keeps on chip
    spec = torch.ones((2, n_ch), device="cuda")
    lohi = torch.zeros((1, n_ch), dtype=torch.int32, device="cuda")
    out = torch.empty((2, n_ch), device="cuda")
    with pytest.raises(_native.NativeError) as err:
        bkg.snip_launch(spec, 0, n_ch, lohi, 1, 5, out, _native.MT_SNIP_BACKGROUND)
    assert err.value.code == _native.MT_ERR_UNSUPPORTED and "on-chip limit" in str(err.value)


def test_fit_argument_checks():
    from mapstorch_b200 import _native, opt
    from mapstorch_b200.default import default_param_vals

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    plan = opt.MapFitPlan((0, 2047), ["Fe", "Cu"], p, use_snip=False)
    good = torch.ones((8, 2048), device="cuda")
    with pytest.raises(ValueError, match="CUDA"):
        plan.fit(torch.ones((8, 2048)))                       # host tensor: fit_host / fit_map is the entry for that
    with pytest.raises(ValueError, match="channels"):
        plan.fit(torch.ones((8, 1024), device="cuda"))        # fewer channels than energy_range needs
    with pytest.raises(ValueError, match="contiguous"):
        plan.fit(torch.ones((2048, 8), device="cuda").t())    # channel stride != 1
    with pytest.raises(ValueError, match="l1_lambda"):
        plan.fit(good, nonneg=False, l1_lambda=1e-6)
    with pytest.raises(ValueError, match="unknown engine"):
        plan.fit(good, engine="cpu")                          # there is no CPU engine
    with pytest.raises(_native.NativeError) as err:           # contraction window outside the operand
        _native.call("mt_fit_contract_simt", good.data_ptr(), 0, 8, 2048, 100, 4096, good.data_ptr(), 2048, 2,
                     good.data_ptr(), 8, _native.MT_X_ELEMENT_MAJOR, None)
    assert err.value.code == _native.MT_ERR_INVALID
    x = plan.fit(good)                                        # and a correct call still works
    torch.cuda.synchronize()
    assert x.shape == (len(plan.names), 8) and torch.isfinite(x).all()


def test_unaligned_rows_take_the_cuda_core_path_not_an_error():
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_param_vals

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    plan = opt.MapFitPlan((0, 2046), ["Fe", "Cu"], p, use_snip=False)
    rng = np.random.default_rng(0)
    y = torch.from_numpy(rng.poisson(5.0, size=(33, 2047)).astype(np.float32)).cuda()  # 8188-byte rows: no TMA
    x = plan.fit(y)
    assert plan._last_engine == "simt"
    padded = torch.zeros((33, 2048), device="cuda")
    padded[:, :2047] = y
    x2 = plan.fit(padded)
    assert plan._last_engine == "tensor"
    assert torch.allclose(x, x2, rtol=2e-4, atol=1e-3 * float(x2.abs().max()))


def test_fit_spec_rejects_unknown_loss_and_optimizer_like_the_reference():
    from mapstorch_b200 import opt

    spec = np.ones(2048, dtype=np.float32)
    with pytest.raises(AssertionError, match="Loss function not supported"):
        opt.fit_spec(spec, (0, 2047), ["Fe"], loss="huber")
    with pytest.raises(AssertionError, match="Optimizer not supported"):
        opt.fit_spec(spec, (0, 2047), ["Fe"], optimizer="sgd")
