#!/usr/bin/env python
"""Small invocation of every kernel for `compute-sanitizer --tool memcheck` (ragged sizes on purpose)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from mapstorch_b200 import opt, synth, bkg
from mapstorch_b200.refine import refine_map

p = synth.benchmark_params(2048)
er = (50, 1450)
plan = opt.MapFitPlan(er, synth.ELEMS10, p, use_snip=True)
truth = synth.truth_maps(len(plan.names), 3, 47, seed=1, device="cuda", lo=0.0, hi=2.5)
full_plan = opt.MapFitPlan((0, 2047), synth.ELEMS10, p, use_snip=False)
vol = synth.poisson_volume(full_plan.basis, truth, dtype=torch.float32, flat=2.0, seed=2).reshape(-1, 2048)
x = plan.fit(vol)                                   # K1, K2 (hi/lo), K3 tf32, epilogue fix-up, fallback solver
x16 = full_plan.fit(vol.half())                     # K3 f16
xs = full_plan.fit(vol, engine="simt")              # CUDA-core contraction + scan solver
xp = full_plan.fit(vol.half(), layout="pe")         # pixel-major
full_plan.epilogue_fixup = False
xl = full_plan.fit(vol.half())                      # list-driven thread-per-pixel solver
plan20 = opt.MapFitPlan((0, 2047), synth.ELEMS20, p, use_snip=False)
t20 = synth.truth_maps(len(plan20.names), 2, 9, seed=3, device="cuda")
v20 = synth.poisson_volume(plan20.basis, t20, seed=4).reshape(-1, 2048)
x20 = plan20.fit(v20, exact_counts=True)
refine_map(plan20, v20, x20, max_iter=4)            # K4
b = bkg.snip_bkg(vol[:5, 50:1451], er, *[torch.tensor(float(p[k])) for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME")], device="cuda")
torch.cuda.synchronize()
print("ok", float(x.sum()), float(x16.sum()), float(xs.sum()), float(xp.sum()), float(xl.sum()), float(x20.sum()), float(b.sum()))
