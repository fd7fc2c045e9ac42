"""Config 1 (opt.fit_spec with the CLI's arguments) timing, work counts and a host profile (diagnostic, GPU box)."""
import cProfile
import json
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.cuda.set_device(0)
print(json.dumps({k: v for k, v in bench.run_config1("cuda:0", repeats=3).items() if k in ("seconds", "final_loss", "work")}))
if os.environ.get("PROFILE", "1") == "1":
    pr = cProfile.Profile()
    pr.enable()
    bench.run_config1("cuda:0", repeats=1)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
