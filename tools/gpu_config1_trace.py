"""Objective trace of config 1's fit_spec (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mapstorch_b200 import opt, refine  # noqa: E402

torch.cuda.set_device(0)
spec, params, elems = bench.config1_spectrum()
kw = dict(elements_to_fit=elems, init_param_vals=params, tune_params=True, init_amp=True, use_snip=True, loss="mse",
          optimizer="adamw", n_iter=500, progress_bar=False, device="cuda")
tensors, spec_fit, bkg, trace = opt.fit_spec(spec, [50, 1450], **kw)
print(refine.LAST_STATS, len(trace))
print(" ".join(f"{v:.6g}" for v in trace))
