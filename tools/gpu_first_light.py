#!/usr/bin/env python
"""Staged bring-up on a GPU box: each stage runs in its own subprocess with a timeout so that a
trapping / hanging kernel cannot hide the results of the others.  Writes gpurun_out/first_light.log."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)

STAGES = {
    "model": """
import numpy as np, torch
from mapstorch_b200 import map as mmap, tables, default as d
from oracle import maps_oracle as mo
p = dict(d.default_param_vals, COHERENT_SCT_ENERGY=12.0)
names = tables.fitted_components(list(d.default_fitting_elems), d.default_energy_consts)
cols = tables.compile_terms(p, names, d.default_energy_consts, d.default_elem_info, True, True)
ev = tables.energy_axis(p, (0, 2047))
out = mmap.run_terms(ev, cols).cpu().numpy()
terms, starts = tables.pack_columns(cols)
ref = mo.eval_terms(terms, starts, ev.numpy())
sc = np.abs(ref).max(axis=1); live = sc > 0
print('K1 worst rel-to-max', (np.abs(out-ref).max(axis=1)[live]/sc[live]).max(), 'live', live.sum())
""",
    "snip": """
import numpy as np, torch
from mapstorch_b200 import bkg
from oracle import maps_oracle as mo
g = np.load('tests/golden/snip_golden.npz')
for tag in ('c2048','c1401','c4096'):
    er = tuple(int(v) for v in g[tag+'_er']); spec=g[tag+'_spec']; ref=g[tag+'_bkg']; pr=g[tag+'_params']
    out = bkg.snip_bkg(torch.from_numpy(spec), er, *[torch.tensor(float(v)) for v in pr], device='cpu').numpy()
    print(tag, 'snip rel-to-max err per row', np.abs(out-ref).max(axis=1)/np.maximum(np.abs(ref).max(axis=1),1e-30))
""",
    "simt_nnls": """
import numpy as np, torch
from mapstorch_b200 import opt
from mapstorch_b200.default import default_param_vals
from oracle import maps_oracle as mo
elems = ["K","Ca","Ti","Cr","Mn","Fe","Co","Ni","Cu","Zn"]
p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0); er=(0,2047)
rng = np.random.default_rng(0)
B, names = mo.basis(p, er, elems)
amps = 10.0 ** rng.uniform(-0.5, 2.5, size=(300, len(names)))
y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
plan = opt.MapFitPlan(er, elems, p, use_snip=False)
x = plan.fit(torch.from_numpy(y).cuda(), nonneg=False, engine='simt').cpu().numpy().T
ref = mo.fit_lstsq(y, B)
print('simt lstsq err rel-to-max', (np.abs(x-ref)/np.abs(ref).max(axis=1,keepdims=True)).max())
x = plan.fit(torch.from_numpy(y).cuda(), nonneg=True, engine='simt').cpu().numpy().T
ref = mo.fit_nnls(y, B)
print('simt nnls err rel-to-max', (np.abs(x-ref)/np.abs(ref).max(axis=1,keepdims=True)).max(), 'zeros frac', (ref==0).mean())
""",
    "tensor_f16": """
import numpy as np, torch
from mapstorch_b200 import opt
from mapstorch_b200.default import default_param_vals
from oracle import maps_oracle as mo
elems = ["K","Ca","Ti","Cr","Mn","Fe","Co","Ni","Cu","Zn"]
p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0); er=(0,2047)
rng = np.random.default_rng(0)
B, names = mo.basis(p, er, elems)
for n_px in (128, 256, 1000, 40000):
    amps = 10.0 ** rng.uniform(1.0, 2.8, size=(n_px, len(names)))
    y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
    plan = opt.MapFitPlan(er, elems, p, use_snip=False)
    ref = mo.fit_lstsq(y, B)
    for dt in (torch.float16, torch.float32):
        x = plan.fit(torch.from_numpy(y).to('cuda', dt), nonneg=False, engine='tensor')
        torch.cuda.synchronize()
        x = x.cpu().numpy().T
        print(n_px, dt, 'tensor lstsq err rel-to-max', (np.abs(x-ref)/np.abs(ref).max(axis=1,keepdims=True)).max(), flush=True)
""",
    "tensor_e30": """
import numpy as np, torch
from mapstorch_b200 import opt
from mapstorch_b200.default import default_param_vals
from oracle import maps_oracle as mo
elems = ("Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge Ba_L Ce_L Gd_L W_L Pt_L Au_L Au_M Pt_M Hg_M W_M Si_Si").split()
p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0, ENERGY_SLOPE=0.00599, ENERGY_QUADRATIC=0.0); er=(0,4095)
rng = np.random.default_rng(0)
B, names = mo.basis(p, er, elems)
n_px = 3000
amps = 10.0 ** rng.uniform(1.5, 3.0, size=(n_px, len(names)))
y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
print('max count', y.max())
plan = opt.MapFitPlan(er, elems, p, use_snip=False)
print('cond(G)', np.linalg.cond(plan.gram))
ref = mo.fit_lstsq(y, B)
for dt in (torch.float16, torch.float32):
    x = plan.fit(torch.from_numpy(y).to('cuda', dt), nonneg=False, engine='tensor').cpu().numpy().T
    e = np.abs(x-ref)/np.abs(ref).max(axis=1,keepdims=True)
    print(dt, 'tensor e30 err rel-to-max: max', e.max(), 'median', np.median(e), 'bias', ((x-ref)/np.abs(ref).max(axis=1,keepdims=True)).mean(), flush=True)
xs = plan.fit(torch.from_numpy(y).cuda(), nonneg=False, engine='simt').cpu().numpy().T
print('simt e30 err', (np.abs(xs-ref)/np.abs(ref).max(axis=1,keepdims=True)).max())
""",
}


def main():
    names = sys.argv[1:] or list(STAGES)
    env = dict(os.environ, PYTHONPATH=str(ROOT), PYTHONDONTWRITEBYTECODE="1")
    with open(OUT / "first_light.log", "w") as log:
        for name in names:
            try:
                r = subprocess.run([sys.executable, "-c", STAGES[name]], cwd=ROOT, env=env, capture_output=True,
                                   text=True, timeout=240)
                msg = f"=== {name}: exit {r.returncode}\n{r.stdout}{r.stderr[-3000:]}\n"
            except subprocess.TimeoutExpired as exc:
                msg = f"=== {name}: TIMEOUT\n{exc.stdout}\n{exc.stderr}\n"
            print(msg, flush=True)
            log.write(msg)


if __name__ == "__main__":
    main()
