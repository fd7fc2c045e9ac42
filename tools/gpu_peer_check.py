#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/gpu_peer_check.py : the fused peer-memory gather of the element
maps (dist.PeerMaps) against the NCCL all-gather (dist.gather_maps), bit for bit, plus timings."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from mapstorch_b200 import opt, synth
from mapstorch_b200.dist import PeerMaps, gather_maps

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
rows, width, n_ch = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 512, 4096)))
lo, hi = (float(a) for a in (sys.argv[4:6] if len(sys.argv) > 5 else (0.5, 2.5)))
p = synth.benchmark_params(n_ch)
plan = opt.MapFitPlan((0, n_ch - 1), synth.ELEMS30, p, use_snip=False)
truth = synth.truth_maps(len(plan.names), rows, width, seed=5, row0=rank * rows, total_rows=world * rows, device=dev, lo=lo, hi=hi)
vol = synth.poisson_volume(plan.basis, truth, dtype=torch.float16, seed=100 + rank).reshape(-1, n_ch)
n_comp, n_px = len(plan.names), vol.shape[0]
local = torch.zeros((n_comp, n_px), device=dev)
plan.fit(vol, out=local)
ref = gather_maps(local, world * rows, width)

def timed(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

def nccl_step():
    plan.fit(vol, out=local)
    return gather_maps(local, world * rows, width)

results = {"nccl_ms": timed(nccl_step), "fit_only_ms": timed(lambda: plan.fit(vol, out=local))}
for mc, layout in (("owner", "ep"), ("gather", "ep"), ("allgather", "ep")):
    try:
        maps = PeerMaps(n_comp, world * rows, width, mode=mc, layout=layout, e_pad=plan.e_pad)
        kw = dict(out=maps.local, layout=layout)
        def step():
            plan.fit(vol, peers=maps.peers, **kw)
            maps.barrier()
        step(); torch.cuda.synchronize()
        got = maps.full[:, :n_comp].T if layout == "pe" else maps.full
        if mc == "owner":
            ok = torch.tensor([int(all(torch.equal(maps.full[j], ref[e]) for j, e in enumerate(maps.owned)))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            same = bool(ok.item())
        else:
            same = torch.equal(got, ref) if maps.has_full_maps() else None
        results[f"peer_multicast={mc}_{layout}"] = {
            "bit_identical_to_nccl_on_rank0": same, "ms": timed(step),
            "no_barrier_ms": timed(lambda: plan.fit(vol, peers=maps.peers, **kw)),
            "contract_only_ms": timed(lambda: plan.fit(vol, peers=maps.peers, nonneg=False, **kw)),
            "contract_local_ms": timed(lambda: plan.fit(vol, nonneg=False, **kw)),
            "fit_local_ms": timed(lambda: plan.fit(vol, **kw))}
    except Exception as exc:
        results[f"peer_multicast={mc}_{layout}"] = f"FAILED: {type(exc).__name__}: {exc}"
if rank == 0:
    import json
    print("world", world, "px/rank", n_px, json.dumps(results, indent=1), flush=True)
dist.destroy_process_group()
