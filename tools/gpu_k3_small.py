"""K3 on small maps: whole tiles only (MT_FIT_TAIL=0) against 64-pixel granules in the last tile (GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapstorch_b200 import opt, synth  # noqa: E402

torch.cuda.set_device(0)
n_ch = 2048
params = synth.benchmark_params(n_ch)
plan = opt.MapFitPlan((0, n_ch - 1), synth.ELEMS10, params, use_snip=False)
truth = synth.truth_maps(len(plan.names), 256, 256, seed=1234, device="cuda")
full = synth.poisson_volume(plan.basis, truth, dtype=torch.float32, flat=0.0, seed=99).reshape(-1, n_ch)
for n_px in (1024, 4096, 16384, 37888, 65536):
    flat = full[:n_px]
    out = torch.empty((plan.n_live, n_px), device="cuda")
    graph = plan.make_graph(flat, out)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"MT_FIT_TAIL={os.environ.get('MT_FIT_TAIL', '1')}  {n_px:6d} px  {float(np.median(ts)) * 1e3:8.1f} us")
