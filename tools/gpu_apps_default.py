"""fit_spec with the apps' default arguments (all 93 default components, 12 tunables, SNIP): time, work counts and a
host profile (diagnostic, GPU box)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapstorch_b200 import opt, refine  # noqa: E402
from mapstorch_b200.default import default_param_vals  # noqa: E402
from oracle import maps_oracle as mo  # noqa: E402  (synthetic spectrum only)

torch.cuda.set_device(0)
p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
elems = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
rng = np.random.default_rng(5)
B, names = mo.basis(p, (0, 2047), elems)
amps = 10.0 ** rng.uniform(3.0, 4.0, size=B.shape[1])
spec = rng.poisson(amps @ B.astype(np.float64).T + 20.0).astype(np.float32)
start = dict(p, ENERGY_OFFSET=p["ENERGY_OFFSET"] + 0.004, FWHM_OFFSET=p["FWHM_OFFSET"] * 1.05)
for n_iter in (25, 500):
    opt.fit_spec(spec, (50, 1450), init_param_vals=start, n_iter=n_iter, progress_bar=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, _, trace = opt.fit_spec(spec, (50, 1450), init_param_vals=start, n_iter=n_iter, progress_bar=False)
    torch.cuda.synchronize()
    print(n_iter, "seconds", time.perf_counter() - t0, refine.LAST_STATS, "loss", trace[0], "->", trace[-1], flush=True)
pr = cProfile.Profile()
pr.enable()
opt.fit_spec(spec, (50, 1450), init_param_vals=start, n_iter=500, progress_bar=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
