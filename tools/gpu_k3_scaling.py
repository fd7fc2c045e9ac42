"""K3 (fit_contract_kernel) time against the number of pixels: fixed cost per launch vs streaming rate (GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapstorch_b200 import opt, synth  # noqa: E402

torch.cuda.set_device(0)
n_ch = 4096
params = synth.benchmark_params(n_ch)
plan = opt.MapFitPlan((0, n_ch - 1), getattr(synth, os.environ.get("ELEMS", "ELEMS30")), params, use_snip=False)
plan.timing = {}
rows = []
# the bench's synthetic scan (512 rows of 2048 pixels: twice the c5 shard)
truth = synth.truth_maps(len(plan.names), 512, 2048, seed=1234, device="cuda")
full = synth.poisson_volume(plan.basis, truth, dtype=torch.float16, flat=0.0, seed=99).reshape(-1, n_ch)
del truth
for n_px in (262144, 524288, 1048576):
    flat = full[:n_px]
    out = torch.empty((plan.n_live, n_px), device="cuda")
    for _ in range(3):
        plan.fit(flat, out=out, exact_counts=True)
    torch.cuda.synchronize()
    ts = []
    graph = plan.make_graph(flat, out)
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    rows.append((n_px, ms))
    print(f"{n_px:8d} px  {ms:.4f} ms  {n_px * (n_ch * 2 + plan.n_live * 4) / ms / 1e6:.0f} GB/s  ({n_px / 256 / 148:.2f} waves of 256-px tiles)")
(n0, t0), (n1, t1) = rows[-2], rows[-1]
rate = (n1 - n0) / (t1 - t0)
print(f"marginal rate {rate * (n_ch * 2 + plan.n_live * 4) / 1e6:.0f} GB/s; fixed cost of a step {t0 - n0 / rate:.4f} ms")
