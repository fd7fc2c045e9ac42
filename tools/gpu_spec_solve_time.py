"""Duration of mt_spec_solve (K7) by CUDA events: config 1's problem size and the apps' default size (diagnostic)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mapstorch_b200 import robust  # noqa: E402
from test_gpu_spec_solve import make_problem  # noqa: E402

torch.cuda.set_device(0)
for n_comp, n_ch, n_prob in ((12, 1401, 1), (12, 1401, 8), (32, 4096, 1), (93, 1401, 1), (96, 4096, 1)):
    basis, y, bkg = make_problem(n_comp, n_ch, seed=1)
    b = torch.as_tensor(basis, device="cuda")
    if n_prob > 1:
        b = b.unsqueeze(0).repeat(n_prob, 1, 1).contiguous()
    yy, bb = torch.as_tensor(y, device="cuda"), torch.as_tensor(bkg, device="cuda")
    for _ in range(3):
        out = robust.spec_solve(b, yy, bb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(20):
        out = robust.spec_solve(b, yy, bb)
    e1.record()
    torch.cuda.synchronize()
    print(f"{n_comp} components x {n_ch} channels x {n_prob} problems: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch, "
          f"{int(out[3][0, 4])} pivot steps")
