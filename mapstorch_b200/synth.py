"""Seeded synthetic spectral volumes for tests and benchmarks.

Ground-truth amplitude maps are smooth 2-D patterns times ``10**U(lo, hi)`` per element; the
expected counts are ``lambda = A @ B`` with B the model basis for the given global parameters
(+ a flat background for SNIP workloads); spectra are Poisson draws, generated on the device in
row slabs so that no host->device copy is involved.
"""
import torch

ELEMS10 = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
ELEMS20 = "Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As".split()
ELEMS30 = ELEMS20[:19] + "Ba_L Ce_L Gd_L W_L Pt_L Au_L Au_M Pt_M Hg_M W_M Si_Si".split()


def benchmark_params(n_ch):
    """default_param_vals at 12 keV incident energy; 4096-channel volumes span the same 0-24.5 keV."""
    from mapstorch_b200.default import default_param_vals

    p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
    if n_ch == 4096:
        p.update(ENERGY_SLOPE=0.00599, ENERGY_QUADRATIC=0.0)
    return p


def truth_maps(n_comp, height, width, seed, lo=1.5, hi=3.0, row0=0, total_rows=None, device="cuda"):
    """[E, height, width] linear amplitudes; rows row0..row0+height of a total_rows-high map."""
    total_rows = total_rows or height
    gen = torch.Generator(device=device).manual_seed(seed)
    level = lo + (hi - lo) * torch.rand(n_comp, device=device, generator=gen)
    fx = 1 + torch.randint(0, 4, (n_comp,), device=device, generator=gen).float()
    fy = 1 + torch.randint(0, 4, (n_comp,), device=device, generator=gen).float()
    ph = 6.2831853 * torch.rand(n_comp, device=device, generator=gen)
    yy = (torch.arange(row0, row0 + height, device=device).float() / total_rows)[None, :, None]
    xx = (torch.arange(width, device=device).float() / width)[None, None, :]
    pattern = 0.6 + 0.4 * torch.sin(6.2831853 * (fx[:, None, None] * xx + fy[:, None, None] * yy) + ph[:, None, None])
    return (10.0 ** level)[:, None, None] * pattern


def poisson_volume(basis, truth, dtype=torch.float32, flat=0.0, seed=0, slab_px=16384, out=None):
    """basis [E, C] (CUDA fp32), truth [E, H, W] -> volume [H, W, C] of Poisson counts in `dtype`."""
    n_comp, n_ch = basis.shape
    h, w = truth.shape[1:]
    amps = truth.reshape(n_comp, -1).T.contiguous()  # [P, E]
    n_px = amps.shape[0]
    vol = out if out is not None else torch.empty((n_px, n_ch), dtype=dtype, device=basis.device)
    gen = torch.Generator(device=basis.device).manual_seed(seed)
    flat_vol = vol.reshape(n_px, n_ch)
    for p0 in range(0, n_px, slab_px):
        lam = amps[p0: p0 + slab_px] @ basis
        if flat:
            lam += flat
        flat_vol[p0: p0 + slab_px] = torch.poisson(lam, generator=gen).to(dtype)
    return flat_vol.reshape(h, w, n_ch)
