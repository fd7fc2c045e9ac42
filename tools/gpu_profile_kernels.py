#!/usr/bin/env python
"""Runs each kernel once at benchmark-like sizes (for `ncu -k regex:...` captures)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from mapstorch_b200 import opt, synth, map as mmap, tables, default as d
from mapstorch_b200.refine import refine_map

dev = torch.device("cuda")
# K1: every default component with step + tail terms, 4096 channels
p = synth.benchmark_params(4096)
p.update(F_STEP_OFFSET=0.002, F_STEP_LINEAR=0.0004, COMPTON_F_STEP=0.01)
names = tables.fitted_components(list(d.default_fitting_elems), d.default_energy_consts)
cols = tables.compile_terms(p, names, d.default_energy_consts, d.default_elem_info, True, True)
ev = tables.energy_axis(p, (0, 4095))
for _ in range(3):
    mmap.run_terms(ev, cols)
# K2 + K3 (tf32, hi/lo residual) + NNLS: config 3 shard, 32768 px
p3 = synth.benchmark_params(4096)
plan = opt.MapFitPlan((0, 4095), synth.ELEMS30, p3, use_snip=True)
truth = synth.truth_maps(len(plan.names), 32, 1024, seed=1, device=dev)
vol = synth.poisson_volume(plan.basis, truth, dtype=torch.float32, flat=3.0, seed=2).reshape(-1, 4096)
for _ in range(3):
    plan.fit(vol)
# K4: config 4, 65536 px
p4 = synth.benchmark_params(2048)
plan4 = opt.MapFitPlan((0, 2047), synth.ELEMS20, p4, use_snip=False)
pb = dict(p4, ENERGY_OFFSET=p4["ENERGY_OFFSET"] + 0.003, FWHM_OFFSET=p4["FWHM_OFFSET"] * 1.04)
truth4 = synth.truth_maps(len(plan4.names), 128, 512, seed=3, device=dev)
vol4 = synth.poisson_volume(opt.MapFitPlan((0, 2047), synth.ELEMS20, pb, use_snip=False).basis, truth4, seed=4).reshape(-1, 2048)
for _ in range(2):
    x = plan4.fit(vol4, exact_counts=True)
    refine_map(plan4, vol4, x, max_iter=12)
torch.cuda.synchronize()
print("done")
