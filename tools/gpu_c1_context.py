"""Why is config 1's fit_spec slower inside the full bench than on its own?  (diagnostic, GPU box)"""
import argparse
import gc
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
args = argparse.Namespace(gpus=1, steps=5, warmup=3, workload="c5", impl="ours", e2e_steps=0, no_cpu_baseline=True, no_e2e=True,
                          no_graph=False, no_extra=True, nccl_gather=False, gather="owner")


def c1(tag):
    r = bench.run_config1(dev, repeats=3)
    print(f"{tag:40s} {r['seconds']:.3f} s   threads={torch.get_num_threads()}  reserved={torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)


c1("fresh process")
for name in ("c2", "c4", "c3", "c5"):
    bench.run_workload(name, args, 0, 0, 1, dev, steps=3, warmup=3, e2e_steps=0, with_cpu_baseline=False)
    c1(f"after workload {name}")
gc.collect()
torch.cuda.empty_cache()
c1("after gc + empty_cache")
