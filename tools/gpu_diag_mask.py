import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_refine as t
from mapstorch_b200 import opt
from mapstorch_b200.default import default_param_vals
from oracle import maps_oracle as mo
p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0); er = (0, 2047)
rng = np.random.default_rng(14); names = mo.component_names(t.ELEMS20)
idx = np.nonzero((np.arange(2048) // 48) % 3 != 1)[0]
specs = []
for i in range(4):
    shift, widen = rng.uniform(-0.006, 0.006), rng.uniform(0.93, 1.08)
    pp = dict(p, ENERGY_OFFSET=p["ENERGY_OFFSET"] + shift, FWHM_OFFSET=p["FWHM_OFFSET"] * widen)
    amps = 10.0 ** rng.uniform(2.0, 3.2, size=len(names))
    model, _, _, _ = mo.gaussian_model_jacobian(pp, er, t.ELEMS20, amps)
    spec = rng.poisson(model).astype(np.float32); spec[(np.arange(2048) // 48) % 3 == 1] += 500.0
    specs.append(spec)
y = np.stack(specs); free = ("ENERGY_OFFSET", "FWHM_OFFSET")
for iters, tol in ((30, 1e-6),):
    maps = opt.fit_map(y, er, t.ELEMS20, init_param_vals=p, use_snip=False, indices=idx, refine=free, refine_iters=iters, refine_tol=tol, device="cpu") if "refine_tol" in opt.fit_map.__code__.co_varnames else opt.fit_map(y, er, t.ELEMS20, init_param_vals=p, use_snip=False, indices=idx, refine=free, refine_iters=iters, device="cpu")
    for i in range(4):
        theta, x, loss = mo.refine_spectrum(y[i], p, er, t.ELEMS20, free=free, indices=idx)
        mine = np.array([float(maps[n][i]) for n in names]); interior = x > 1e-3 * x.max()
        print(iters, i, "doff", float(maps["ENERGY_OFFSET"][i]) - theta[0], "fwhm rel", float(maps["FWHM_OFFSET"][i]) / theta[1] - 1,
              "loss ratio", float(maps["loss"][i]) / loss - 1, "amp", (np.abs(mine - x)[interior] / x[interior]).max(), "it", int(maps["iterations"][i]))
