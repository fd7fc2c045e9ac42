"""Out-of-bounds guard for the kernels' OUTPUTS (compute-sanitizer refuses to run on this pool).

Every output array is carved out of a larger sentinel-filled buffer, with a row pitch wider than the row and
slack before and after; after the call everything outside the intended region must still be the sentinel.
Sizes are deliberately awkward (odd pixel counts, cropped windows, slab views of a wider map)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SENTINEL = -12345.0


class Guarded:
    """[rows, cols] view with pitch > cols inside a sentinel-filled buffer"""

    def __init__(self, rows, cols, pitch=None, dtype=torch.float32, lead=64, tail=64):
        pitch = pitch or cols + 8
        self.buf = torch.full((lead + rows * pitch + tail,), SENTINEL, dtype=torch.float32, device="cuda").to(dtype) \
            if dtype != torch.int32 else torch.full((lead + rows * pitch + tail,), int(SENTINEL), dtype=dtype, device="cuda")
        self.view = self.buf[lead: lead + rows * pitch].view(rows, pitch)[:, :cols]
        self.rows, self.cols, self.pitch, self.lead = rows, cols, pitch, lead

    def intact(self):
        mask = torch.ones_like(self.buf, dtype=torch.bool)
        inner = mask[self.lead: self.lead + self.rows * self.pitch].view(self.rows, self.pitch)
        inner[:, : self.cols] = False
        outside = self.buf[mask]
        return bool((outside == outside.new_tensor(SENTINEL if self.buf.dtype != torch.int32 else int(SENTINEL))).all())


def _plan(n_ch=2047, snip=False, elems=("Fe", "Cu", "Zn", "Ca", "Ti")):
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_param_vals

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    return opt.MapFitPlan((16, 16 + n_ch - 1), list(elems), p, use_snip=snip), p


@pytest.mark.parametrize("n_px", [1, 37, 515])
def test_snip_outputs_stay_inside_their_rows(n_px):
    from mapstorch_b200 import _native, bkg

    n_ch = 1401
    spec = torch.rand((n_px, 2048), device="cuda") * 50
    lohi, n_pass = bkg.snip_tables_device(n_ch, (50, 1450), *[torch.tensor(v) for v in
                                                           (-0.002, 0.01198, 0.0, 0.5, 0.084, 0.00018)])
    for mode, width in ((_native.MT_SNIP_BACKGROUND, n_ch), (_native.MT_SNIP_RESIDUAL, n_ch),
                        (_native.MT_SNIP_RESIDUAL_HILO, 2 * n_ch)):
        out = Guarded(n_px, width, pitch=2 * n_ch + 4)
        bkg.snip_launch(spec, 50, n_ch, lohi, n_pass, 5, out.view, mode)
        torch.cuda.synchronize()
        assert out.intact(), (n_px, mode)
        assert torch.isfinite(out.view).all() and not (out.view == SENTINEL).all()


@pytest.mark.parametrize("n_px,dtype", [(1, torch.float32), (129, torch.float16), (1000, torch.float32)])
def test_fit_outputs_stay_inside_the_map(n_px, dtype):
    plan, _ = _plan()
    y = (torch.rand((n_px, 2080), device="cuda") * 40).round().to(dtype)
    # the slab's maps are a column range of a wider map, as in slab-wise fitting
    wide = Guarded(len(plan.names), n_px + 300, pitch=n_px + 333)
    x = wide.view[:, 100: 100 + n_px]
    before = wide.view.clone()
    plan.fit(y, out=x, exact_counts=True)
    torch.cuda.synchronize()
    assert wide.intact()
    assert torch.equal(wide.view[:, :100], before[:, :100]) and torch.equal(wide.view[:, 100 + n_px:], before[:, 100 + n_px:])
    assert (x >= 0).all()
    pm = Guarded(n_px, plan.e_pad, pitch=plan.e_pad + 4)
    plan.fit(y, out=pm.view, layout="pe", exact_counts=True)
    torch.cuda.synchronize()
    assert pm.intact()
    assert torch.allclose(pm.view[:, : len(plan.names)].T, x, rtol=0, atol=0)


def test_refine_outputs_stay_inside_their_slab():
    from mapstorch_b200.refine import refine_map

    plan, _ = _plan(n_ch=2048 - 16)
    n_px, total = 77, 300
    y = (torch.rand((n_px, 2048), device="cuda") * 40).round()
    wide = Guarded(len(plan.names), total, pitch=total + 5)
    x = wide.view[:, 200: 200 + n_px]
    plan.fit(y, out=x, exact_counts=True)
    before = wide.view.clone()
    theta, loss, iters = refine_map(plan, y, x, max_iter=4)
    torch.cuda.synchronize()
    assert wide.intact()
    assert torch.equal(wide.view[:, :200], before[:, :200]) and torch.equal(wide.view[:, 200 + n_px:], before[:, 200 + n_px:])
    assert theta.shape == (2, n_px) and torch.isfinite(theta).all() and torch.isfinite(loss).all()
    assert int(iters.min()) >= 0 and int(iters.max()) <= 4 + 8


def test_split_and_penalty_kernels_stay_inside():
    from mapstorch_b200 import _native

    n_px, n_ch = 53, 1001
    spec = (torch.rand((n_px, 1100), device="cuda") * 9000).round()
    out = Guarded(n_px, 2 * n_ch, pitch=2 * n_ch + 6)
    _native.call("mt_split_hilo", spec.data_ptr(), n_px, spec.stride(0), 33, n_ch, out.view.data_ptr(), out.pitch,
                 _native.stream_ptr())
    torch.cuda.synchronize()
    assert out.intact()
    assert torch.equal(out.view[:, :n_ch] + out.view[:, n_ch:], spec[:, 33: 33 + n_ch])
    assert not bool((out.view[:, :n_ch].view(torch.int32) & 0x1FFF).any())
    x = Guarded(7, n_px, pitch=n_px + 9)
    x.view.zero_()
    coef = torch.arange(1, 8, device="cuda", dtype=torch.float32) * 1e-9
    _native.call("mt_penalty_shift", spec.data_ptr(), 0, n_px, spec.stride(0), 33, n_ch, 1, None, coef.data_ptr(),
                 x.view.data_ptr(), x.pitch, _native.MT_X_ELEMENT_MAJOR, 7, _native.stream_ptr())
    torch.cuda.synchronize()
    assert x.intact()
    want = -(spec[:, 33: 33 + n_ch].double() ** 2).sum(dim=1)[None, :] * coef.double()[:, None]
    assert torch.allclose(x.view.double(), want, rtol=1e-5)
