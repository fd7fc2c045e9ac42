"""Where do the strided and the contiguous-run SNIP kernels differ?  (diagnostic, run on the GPU box)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mapstorch_b200 import _native, bkg  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "snip_golden.npz"))
pr = g["c4096_params"]
n_ch, n_px = 4096, int(os.environ.get("NPX", 1203))
rng = np.random.default_rng(1)
chan = np.arange(n_ch)
peaks = 900 * np.exp(-0.5 * ((chan - 0.4 * n_ch) / 7.0) ** 2) + 300 * np.exp(-0.5 * ((chan - 0.8 * n_ch) / 15.0) ** 2)
spec = rng.poisson(rng.uniform(0, 40, size=(n_px, 1)) + peaks).astype(np.float32)
dev = torch.from_numpy(spec).cuda()
lohi, n_pass = bkg.snip_tables_device(n_ch, (0, n_ch - 1), *[torch.tensor(float(v)) for v in pr])
runs = lohi._mt_runs
desc = runs.cpu().numpy().view(np.uint32)[:, -512:]
for npass in (0, 1, 3, 6, 8, n_pass):
    old = torch.empty((n_px, n_ch), device="cuda")
    new = torch.empty((n_px, n_ch), device="cuda")
    _native.call("mt_snip", dev.data_ptr(), 0, n_px, n_ch, 0, n_ch, lohi.data_ptr(), npass, 5, old.data_ptr(), n_ch, 0,
                 _native.stream_ptr())
    _native.call("mt_snip_runs", dev.data_ptr(), 0, n_px, n_ch, 0, n_ch, runs.data_ptr(), runs.stride(0), npass, 5,
                 new.data_ptr(), n_ch, 0, None, _native.stream_ptr())
    torch.cuda.synchronize()
    d = (old != new).cpu().numpy()
    o, n = old.cpu().numpy(), new.cpu().numpy()
    print(f"passes {npass}: mismatches {d.sum()}  pixels affected {d.any(1).sum()}  channels affected {d.any(0).sum()}  "
          f"max abs diff {np.abs(o - n).max():.3g}  max rel {np.abs(o - n).max() / o.max():.3g}")
    if d.any():
        ch = np.nonzero(d.any(0))[0]
        print("   first channels", ch[:24], " pixel histogram (px % 4):", np.bincount(np.nonzero(d.any(1))[0] % 4, minlength=4))
        if npass > 0:
            print("   descriptor kinds of the last pass among mismatching runs:",
                  np.unique(desc[npass - 1][np.unique(ch // 8)], return_counts=True))
