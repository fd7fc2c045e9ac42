#!/usr/bin/env python
"""Benchmark of the per-pixel XRF map fit (BASELINE.json metric: pixels/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A step is one fit of
the rank's resident row shard (plus, for N > 1, the all-gather of the element maps).  Rank 0
prints ONE JSON line.  `--impl reference` times the CPU baseline (the oracle's composition of the
reference functions) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

# per-GPU workloads (weak scaling: every rank gets one of these; see DESIGN.md "Measurement")
WORKLOADS = {
    # BASELINE config 5 = 2048x2048x4096 fp16 on 8 GPUs -> each GPU owns 256 rows
    "c5": dict(rows=256, width=2048, n_ch=4096, dtype="f16", elems="ELEMS30", snip=False, flat=0.0,
               desc="BASELINE config 5 row shard: 256x2048x4096 fp16 per GPU (8 GPUs = the 2048x2048x4096 map), "
                    "30 elements + 2 scatter peaks"),
    # BASELINE config 2 (the whole map on every GPU)
    "c2": dict(rows=256, width=256, n_ch=2048, dtype="f32", elems="ELEMS10", snip=False, flat=0.0,
               desc="BASELINE config 2: 256x256x2048 fp32, 10 elements + 2 scatter peaks"),
    # BASELINE config 3 = 1024x1024x4096 fp32 + SNIP on 8 GPUs -> each GPU owns 128 rows
    "c3": dict(rows=128, width=1024, n_ch=4096, dtype="f32", elems="ELEMS30", snip=True, flat=3.0,
               desc="BASELINE config 3 row shard: 128x1024x4096 fp32 per GPU (8 GPUs = the 1024x1024x4096 map), "
                    "30 elements + 2 scatter peaks, SNIP background"),
    # BASELINE config 4: 512x512x2048 fp32, 20 elements, per-pixel refinement of peak offset / width
    "c4": dict(rows=512, width=512, n_ch=2048, dtype="f32", elems="ELEMS20", snip=False, flat=0.0,
               refine=("ENERGY_OFFSET", "FWHM_OFFSET"), refine_iters=12,
               desc="BASELINE config 4: 512x512x2048 fp32, 20 elements + 2 scatter peaks, per-pixel Gauss-Newton "
                    "refinement of ENERGY_OFFSET and FWHM_OFFSET + amplitudes (<= 12 iterations)"),
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc, self.path = None, None
        try:
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in Path(self.path).read_text().strip().splitlines() if r.strip()]
            sm = [float(r[0]) for r in rows]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.strip().lower() == "active"})
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                   "samples": len(rows), "power_w_max": max(float(r[2]) for r in rows)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


def build_workload(name, rank, world, device):
    from mapstorch_b200 import opt, synth

    w = WORKLOADS[name]
    elems = getattr(synth, w["elems"])
    params = synth.benchmark_params(w["n_ch"])
    er = (0, w["n_ch"] - 1)
    plan = opt.MapFitPlan(er, elems, params, use_snip=w["snip"])
    truth = synth.truth_maps(len(plan.names), w["rows"], w["width"], seed=1234, row0=rank * w["rows"],
                             total_rows=world * w["rows"], device=device)
    dtype = torch.float16 if w["dtype"] == "f16" else torch.float32
    if w.get("refine"):
        # row bands with their own energy offset / peak width, so that the refinement has work to do
        vol = torch.empty((w["rows"], w["width"], w["n_ch"]), dtype=dtype, device=device)
        bands = 8
        rows = w["rows"] // bands
        for b in range(bands):
            pb = dict(params, ENERGY_OFFSET=params["ENERGY_OFFSET"] + 0.004 * (b - 3.5) / 3.5,
                      FWHM_OFFSET=params["FWHM_OFFSET"] * (1 + 0.05 * (b - 3.5) / 3.5))
            band_plan = opt.MapFitPlan(er, elems, pb, use_snip=False)
            synth.poisson_volume(band_plan.basis, truth[:, b * rows:(b + 1) * rows], dtype=dtype, flat=w["flat"],
                                 seed=99 + rank * bands + b, out=vol[b * rows:(b + 1) * rows].reshape(-1, w["n_ch"]))
    else:
        vol = synth.poisson_volume(plan.basis, truth, dtype=dtype, flat=w["flat"], seed=99 + rank)
    return w, elems, params, er, plan, truth, vol


def _refine_worker(job):
    """one pixel of the CPU refinement baseline (runs in a forked worker; numpy / scipy only)"""
    from oracle import maps_oracle as mo

    y, params, er, elems, free = job
    theta, x, loss = mo.refine_spectrum(y, params, er, elems, free=free)
    return float(loss)


def cpu_baseline_refine(w, params, er, elems, budget_s=12.0):
    """Config 4 on the host cores: the oracle's per-pixel refinement (variable projection over the free
    calibration parameters + scipy least_squares, converging to the same optimum as the reference's per-pixel
    fit_spec(fitting_params=..., tune_params=True) in far fewer model evaluations than its 500 Adam steps),
    one worker process per core, on a bounded sample."""
    import multiprocessing as mp

    from oracle import maps_oracle as mo

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    B, _ = mo.basis(params, er, elems)

    def sample(n):
        amps = 10.0 ** rng.uniform(1.5, 3.0, size=(n, B.shape[1]))
        y = rng.poisson(amps @ B.astype(np.float64).T + w["flat"]).astype(np.float32)
        return [(y[i], params, er, list(elems), tuple(w["refine"])) for i in range(n)]

    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_refine_worker, sample(cores))  # start the workers, import scipy
        t0 = time.perf_counter()
        pool.map(_refine_worker, sample(2 * cores), chunksize=1)
        per_round = (time.perf_counter() - t0) / 2  # seconds for one pixel on every core
        n = cores * max(2, int(budget_s / max(per_round, 1e-3)))
        jobs = sample(n)
        t0 = time.perf_counter()
        pool.map(_refine_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "px/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"{n} synthetic pixels of the same workload: oracle.refine_spectrum per pixel (free = "
                      f"{', '.join(w['refine'])}; amplitudes by non-negative least squares inside scipy least_squares), "
                      f"one process per core, {dt:.1f} s"}


def cpu_baseline(w, params, er, elems, budget_s=12.0, threads=None):
    """Oracle composition of the reference functions on the host cores, on a bounded sample."""
    from oracle import maps_oracle as mo

    if w.get("refine"):
        return cpu_baseline_refine(w, params, er, elems, budget_s)
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    B, names = mo.basis(params, er, elems)
    snip_args = [params[k] for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET",
                                     "FWHM_FANOPRIME")] if w["snip"] else None

    def sample(n):
        amps = 10.0 ** rng.uniform(1.5, 3.0, size=(n, B.shape[1]))
        return rng.poisson(amps @ B.astype(np.float64).T + w["flat"]).astype(np.float32)

    n = 512
    y = sample(n)
    t0 = time.perf_counter()
    mo.fit_map_composed(y, B, snip_args, er)
    dt = time.perf_counter() - t0
    n_big = int(min(max(n, n * budget_s / max(dt, 1e-6)), 1 << 18))
    y = sample(n_big)
    t0 = time.perf_counter()
    mo.fit_map_composed(y, B, snip_args, er)
    dt = time.perf_counter() - t0
    # the reference's own per-spectrum fit: Adam, 500 iterations (restated with analytic gradients)
    t1 = time.perf_counter()
    for i in range(4):
        mo.fit_spec_adam(y[i] if not w["snip"] else y[i], B, None, np.full(B.shape[1], 2.0, np.float32), n_iter=500)
    adam = 4 / (time.perf_counter() - t1)
    return {"value": n_big / dt, "unit": "px/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"{n_big} synthetic pixels of the same workload: oracle composition of the reference functions "
                      f"({'snip_bkg + ' if w['snip'] else ''}fp32 BLAS contraction with pinv(B) + scipy NNLS on negative "
                      f"pixels), {dt:.1f} s",
            "reference_adam_500it_px_s": adam}


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    from mapstorch_b200 import synth

    w = WORKLOADS[args.workload]
    elems = getattr(synth, w["elems"])
    params = synth.benchmark_params(w["n_ch"])
    er = (0, w["n_ch"] - 1)
    vals = []
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(w, params, er, elems, budget_s=per_step_budget)
        if i >= args.warmup:
            vals.append(cb)
    value = float(np.mean([c["value"] for c in vals]))
    cb = dict(vals[-1], value=value)
    line = {"impl": "reference", "metric": "pixels/sec", "value": value, "unit": "px/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([c["seconds"] for c in vals])) * 1e3,  # one step = one bounded CPU sample
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "name": args.workload},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def bind_near_gpu(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (same NUMA node / PCIe root), so
    that the pinned staging buffers of the end-to-end leg are allocated next to the GPU they feed.  With 8 ranks
    streaming 4.3 GB each, remote-socket pinned memory halves the per-GPU host-to-device rate.  Returns the
    previous affinity (to restore) or None when NVML / sched_setaffinity is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        previous = os.sched_getaffinity(0)
        cpus &= previous
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return previous
    except Exception as exc:  # noqa: BLE001 - best effort, the run is valid without it
        print(f"[bench] no NUMA binding: {exc}", file=sys.stderr)
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true", help="gather the maps with NCCL instead of peer-memory stores")
    ap.add_argument("--gather", default="owner", choices=["owner", "gather", "allgather"],
                    help="peer-memory exchange of the element maps: every component assembled on its owner GPU "
                         "(balanced all-to-all, default), everything on rank 0, or everything everywhere")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    # libraries (NCCL's version banner) write to fd 1; keep stdout for the ONE JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from mapstorch_b200 import _native, opt
    from mapstorch_b200.dist import gather_maps

    early_cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and WORKLOADS[args.workload].get("refine"):
        # this baseline forks one worker per core: run it before the CUDA context exists
        from mapstorch_b200 import synth as _synth

        _w = WORKLOADS[args.workload]
        early_cb = cpu_baseline(_w, _synth.benchmark_params(_w["n_ch"]), (0, _w["n_ch"] - 1), getattr(_synth, _w["elems"]))

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    w, elems, params, er, plan, truth, vol = build_workload(args.workload, rank, world, device)
    flat = vol.reshape(-1, vol.shape[-1])
    n_px = flat.shape[0]
    n_comp = len(plan.names)
    out = torch.zeros((n_comp, n_px), dtype=torch.float32, device=device)
    total_rows = world * w["rows"]

    peer_maps = None
    if world > 1 and not w["snip"] and not args.nccl_gather:
        try:
            from mapstorch_b200.dist import PeerMaps
            peer_maps = PeerMaps(n_comp, total_rows, w["width"], mode=args.gather)
        except Exception as exc:  # pragma: no cover - needs symmetric memory support
            print(f"bench: peer-memory maps unavailable ({exc}); gathering with NCCL", file=sys.stderr)
    graph = None
    if not args.no_graph and world == 1:
        try:
            graph = plan.make_graph(flat, out)
        except Exception as exc:  # pragma: no cover - depends on the driver
            print(f"bench: CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)

    refine_state = {}

    def step(gather=True, eager=False):
        if peer_maps is not None and gather:
            # fused: the fit kernels store the amplitudes into the destination GPUs' map buffers over NVLink
            plan.fit(flat, out=peer_maps.local, peers=peer_maps.peers)
            peer_maps.barrier()
            return peer_maps.full
        if graph is not None and not eager:
            graph.replay()
        else:
            plan.fit(flat, out=out)
        if w.get("refine"):
            from mapstorch_b200.refine import refine_map
            refine_state["out"] = refine_map(plan, flat, out, free=w["refine"], max_iter=w["refine_iters"])
        if world > 1 and gather:
            return gather_maps(out, total_rows, w["width"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step()
    # accuracy of what is being timed: amplitudes against the ground truth (Poisson-limited) and
    # against a float64 solve of a pixel sample
    got = plan.fit(flat, out=out).clone()
    sample = torch.arange(0, n_px, max(1, n_px // 512), device=device)[:512]
    y64 = flat[sample].double()
    if w["snip"]:
        from mapstorch_b200 import bkg as _bkg
        pr = [torch.tensor(float(params[k])) for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH",
                                                       "FWHM_OFFSET", "FWHM_FANOPRIME")]
        y64 = y64 - _bkg.snip_bkg(flat[sample].float(), er, *pr, device="cuda").double()
    ref = torch.linalg.lstsq(plan.basis.double().T, y64.T).solution.T  # [n, E]
    mine = got[:, sample].T.double()
    free = (ref > 0).all(dim=1)  # pixels where the non-negativity constraint is inactive
    scale = ref.abs().max(dim=1, keepdim=True).values
    interior = (ref > 1e-3 * scale) & free[:, None]
    check = {"pixels": int(free.sum()),
             "max_abs_err_over_pixel_max": float((((mine - ref).abs() / scale)[free]).max()) if free.any() else None,
             "max_rel_err_interior": float(((mine - ref).abs() / ref.abs())[interior].max()) if interior.any() else None}
    if int(free.sum()) < 64:
        # (SNIP workloads: almost every pixel has an active constraint) compare a subsample with float64 NNLS instead
        from scipy.optimize import nnls as _nnls

        b64 = plan.basis.double().cpu().numpy().T  # [C, E]
        sub = list(range(0, len(sample), max(1, len(sample) // 96)))[:96]
        ys = y64[sub].cpu().numpy()
        ref_nn = np.stack([_nnls(b64, row, maxiter=50 * b64.shape[1])[0] for row in ys])
        got_nn = mine[sub].cpu().numpy()
        top = np.abs(ref_nn).max(axis=1, keepdims=True)
        inner = ref_nn > 1e-3 * top
        check.update({"nnls_pixels": len(sub),
                      "nnls_max_abs_err_over_pixel_max": float((np.abs(got_nn - ref_nn) / top).max()),
                      "nnls_max_rel_err_interior": float((np.abs(got_nn - ref_nn) / np.maximum(ref_nn, 1e-30))[inner].max())})

    # ---- device-resident timing (value) ---------------------------------------------------------
    sampler = ClockSampler(device)
    time.sleep(0.6)  # let nvidia-smi finish its NVML start-up before the timed region
    for _ in range(2):
        step()
    ms_total = timed(step, args.steps)
    # same steps launched eagerly with CUDA events around the dominant kernel (roofline numerator)
    plan.timers = []
    launches0 = _native.launch_count
    ms_eager = timed(lambda: step(gather=False, eager=True), args.steps)
    launches = _native.launch_count - launches0
    clocks = sampler.stop()
    timers, plan.timers = plan.timers, None
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in timers]))
    kernel_px = float(np.mean([n for _, _, n in timers]))
    ms_fit_only = timed(lambda: step(gather=False), args.steps) if world > 1 else ms_total
    ms_per_step = ms_total / args.steps
    value = world * n_px / (ms_per_step * 1e-3)

    es = flat.element_size()
    in_row_bytes = w["n_ch"] * es * (1 if not w["snip"] else 1)
    bytes_per_px = w["n_ch"] * (es if not w["snip"] else 4 * 2) + n_comp * 4  # what the contraction kernel streams
    peak, peak_src = measured_peak()
    achieved = bytes_per_px * kernel_px / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "fit_contract_kernel (tcgen05 + TMA)", "bytes_per_px": bytes_per_px,
                "px_per_launch": kernel_px, "kernel_ms": kernel_ms, "peak_source": peak_src,
                "job_frac": (w["n_ch"] * es + n_comp * 4) * n_px / (ms_fit_only / args.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the public API, host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        previous_affinity = bind_near_gpu(local_rank)
        host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
        host.copy_(flat)
        torch.cuda.synchronize()

        def e2e_step():
            maps = opt.fit_map(host.reshape(vol.shape), er, elems, init_param_vals=params, use_snip=w["snip"],
                               device="cpu", refine=w.get("refine"), refine_iters=w.get("refine_iters", 12))
            return maps

        e2e_step()
        ms = timed(e2e_step, args.e2e_steps) / args.e2e_steps
        e2e = {"value": world * n_px / (ms * 1e-3), "unit": "px/s", "h2d_bytes_per_step": world * flat.numel() * es,
               "d2h_bytes_per_step": world * n_comp * n_px * 4, "ms_per_step": ms, "steps": args.e2e_steps,
               "api": "mapstorch_b200.opt.fit_map(pinned host volume) -> host maps; plan build, chunked H2D "
                      "overlapped with the fit, D2H of the maps all inside the timed region",
               "numa_bound": previous_affinity is not None}
        del host
        if previous_affinity is not None:
            os.sched_setaffinity(0, previous_affinity)  # the CPU baseline below uses every core again

    cb = early_cb
    if cb is None and rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(w, params, er, elems)

    if rank == 0:
        line = {"metric": "pixels/sec", "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 x f16 -> f32 (tcgen05)" if w["dtype"] == "f16" else "tf32 x tf32 -> f32 (tcgen05)",
                "data": "synthetic",
                "config": {"workload": w["desc"], "name": args.workload, "pixels_per_gpu": n_px, "channels": w["n_ch"],
                           "components": n_comp, "storage": w["dtype"], "snip": w["snip"],
                           "l2": f"inputs larger than L2 ({flat.numel() * es / 1e6:.0f} MB per GPU per step), no flush",
                           "step": "fit of the resident shard (contraction + non-negative fix-up)" +
                                   ("" if world == 1 else " + element maps assembled over NVLink by the fit kernels' own peer stores "
                                    + {"owner": "(every component's full map on its owner GPU: balanced all-to-all)",
                                       "gather": "(all maps on rank 0)", "allgather": "(all maps on every GPU)"}[peer_maps.mode]
                                    + " + barrier" if peer_maps is not None
                                    else " + NCCL all-gather of the element maps")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "cuda_graph": graph is not None,
                "ms_per_step_eager_launch": ms_eager / args.steps, "roofline": roofline, "cpu_baseline": cb,
                "fit_only_px_s": world * n_px / (ms_fit_only / args.steps * 1e-3),
                "refine": ({"free": list(w["refine"]), "mean_iterations": float(refine_state["out"][2].float().mean()),
                            "mean_loss": float(refine_state["out"][1].mean())} if w.get("refine") else None),
                "check_vs_float64_lstsq": check}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
