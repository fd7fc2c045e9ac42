"""Config 1: work counts, time and final loss against the Broyden refresh interval (diagnostic, GPU box)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mapstorch_b200 import refine  # noqa: E402

torch.cuda.set_device(0)
for refresh in (3, 5, 8, 12, 20, 1000):
    refine.LM_REFRESH = refresh
    r = bench.run_config1("cuda:0", repeats=3)
    print(refresh, json.dumps({k: r[k] for k in ("seconds", "final_loss", "work")}), flush=True)
