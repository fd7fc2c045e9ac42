#!/usr/bin/env python
"""Benchmark of the per-pixel XRF map fit (BASELINE.json metric: pixels/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  A step is one fit of the rank's
resident row shard plus, for N > 1, the exchange of the element maps (fused into the fit kernels as NVLink peer
stores, dist.ShardedMapFit).  Rank 0 prints ONE JSON line: the headline workload (default c5) at the top level
and, in "workloads", short runs of the other BASELINE configurations (c2, c3, c4 and config 1's fit_spec), each
with the roofline of ITS dominant kernel.  `--impl reference` times the reference's own CPU code
(baseline/_ref, installed by tools/install_reference.py) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings
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
SNIP_KEYS = ("ENERGY_OFFSET", "ENERGY_SLOPE", "ENERGY_QUADRATIC", "SNIP_WIDTH", "FWHM_OFFSET", "FWHM_FANOPRIME")
C1_DESC = ("BASELINE config 1: opt.fit_spec on one synthetic 2048-channel integrated spectrum, 10 K-line elements + 2 "
           "scatter peaks, 12 keV, the CLI's arguments (energy_range=[50,1450], tune_params=True, init_amp=True, "
           "use_snip=True, loss='mse', optimizer='adamw', n_iter=500)")


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


# ---- CPU baselines ------------------------------------------------------------------------------------------------
def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def reference_modules():
    """The unmodified reference, imported from baseline/_ref (tools/install_reference.py); None if absent."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "mapstorch" / "opt.py").exists():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    warnings.filterwarnings("ignore")
    try:
        import mapstorch.bkg as rbkg
        import mapstorch.map as rmap
        import mapstorch.opt as ropt
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] reference in baseline/_ref does not import: {exc}", file=sys.stderr)
        return None
    return ropt, rmap, rbkg


def reference_basis(rmap, params, er, elems):
    """B [E, C] from the reference's own model functions at unit amplitude (what model_spec sums, map.py:282-304)."""
    t = {k: torch.tensor(float(v)) for k, v in params.items()}
    for e in list(elems) + ["COMPTON_AMPLITUDE", "COHERENT_SCT_AMPLITUDE"]:
        t[e] = torch.tensor(0.0)
    ch = torch.arange(er[0], er[1] + 1, dtype=torch.float32)
    ev = t["ENERGY_OFFSET"] + t["ENERGY_SLOPE"] * ch + t["ENERGY_QUADRATIC"] * (ch ** 2)
    cols = [rmap.model_elem_spec(t, e, ev, True, False) for e in elems]
    cols.append(rmap.elastic_peak(t, ev, t["ENERGY_SLOPE"]))  # component order of MapFitPlan.names / oracle.basis
    cols.append(rmap.compton_peak(t, ev, t["ENERGY_SLOPE"], True, False))
    return torch.stack(cols)


def reference_composed(mods, y, params, er, elems, snip):
    """The reference's functions composed into a map fit, best case: its batched snip_bkg (bkg.py:69-114), its basis,
    one fp32 contraction with pinv(B) in torch and NNLS (scipy) for the pixels whose solution has a negative
    component -- what its 10**a parametrisation converges to.  The reference itself ships no per-pixel fitter."""
    from scipy.optimize import nnls

    ropt, rmap, rbkg = mods
    yt = torch.from_numpy(y)
    if snip:
        yt = yt - rbkg.snip_bkg(yt, er, *(torch.tensor(float(params[k])) for k in SNIP_KEYS))
    B = reference_basis(rmap, params, er, elems)
    B64 = B.double()
    pinv = torch.linalg.solve(B64 @ B64.T, B64).float()
    x = (yt @ pinv.T).numpy()
    bad = np.nonzero((x < 0).any(axis=1))[0]
    b_np, y_np = B64.numpy().T, yt.double().numpy()
    for i in bad:
        x[i] = nnls(b_np, y_np[i], maxiter=50 * b_np.shape[1])[0]
    return x


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

    cores = host_cores()
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
    """The map fit on the host cores, on a bounded sample: the REFERENCE's own functions composed (kind
    "reference") when baseline/_ref is installed, otherwise the oracle's restatement of the same composition
    (kind "port")."""
    from oracle import maps_oracle as mo

    if w.get("refine"):
        return cpu_baseline_refine(w, params, er, elems, budget_s)
    cores = threads or host_cores()
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    B, names = mo.basis(params, er, elems)
    snip_args = [params[k] for k in SNIP_KEYS] if w["snip"] else None
    mods = reference_modules()

    def sample(n):
        amps = 10.0 ** rng.uniform(1.5, 3.0, size=(n, B.shape[1]))
        return rng.poisson(amps @ B.astype(np.float64).T + w["flat"]).astype(np.float32)

    def run(y):
        if mods is not None:
            return reference_composed(mods, y, params, er, elems, w["snip"])
        return mo.fit_map_composed(y, B, snip_args, er)

    n = 512
    y = sample(n)
    t0 = time.perf_counter()
    run(y)
    dt = time.perf_counter() - t0
    n_big = int(min(max(n, n * budget_s / max(dt, 1e-6)), 1 << 18))
    y = sample(n_big)
    t0 = time.perf_counter()
    x = run(y)
    dt = time.perf_counter() - t0
    out = {"value": n_big / dt, "unit": "px/s", "cores": cores, "kind": "reference" if mods is not None else "port",
           "seconds": dt}
    if mods is not None:
        # the composition is only a baseline if it computes the same maps: compare a few pixels with the oracle
        ref = mo.fit_nnls(y[:32], B, mo.snip_bkg(y[:32], er, *snip_args) if w["snip"] else None)
        scale = ref.max(axis=1, keepdims=True)
        out["max_abs_diff_vs_oracle_over_pixel_max"] = float((np.abs(x[:32] - ref) / scale).max())
        out["sample"] = (f"{n_big} synthetic pixels of the same workload through the reference's own functions from "
                         f"baseline/_ref ({'mapstorch.bkg.snip_bkg batched + ' if w['snip'] else ''}mapstorch.map basis + "
                         f"one fp32 torch contraction with pinv(B) + scipy NNLS on negative pixels), {dt:.1f} s")
    else:
        out["sample"] = (f"{n_big} synthetic pixels of the same workload: oracle composition of the reference functions "
                         f"({'snip_bkg + ' if w['snip'] else ''}fp32 BLAS contraction with pinv(B) + scipy NNLS on negative "
                         f"pixels), {dt:.1f} s")
    return out


def _reference_fit_worker(job):
    """The reference's actual per-spectrum fit, one pixel: fit_spec(tune_params=False) with Adam (opt.py:165-327)."""
    spec, er, elems, params, n_iter, use_snip = job
    torch.set_num_threads(1)
    ropt, _, _ = reference_modules()
    t0 = time.perf_counter()
    ropt.fit_spec(spec, er, elements_to_fit=list(elems), init_param_vals=params, tune_params=False, init_amp=True,
                  use_snip=use_snip, loss="mse", optimizer="adam", n_iter=n_iter, progress_bar=False)
    return time.perf_counter() - t0


def reference_fit_spec_rate(w, params, er, elems, budget_s=25.0):
    """px/s of the reference's own fit_spec(tune_params=False, n_iter=500) per pixel, one process per core.  The
    per-iteration cost is constant, so the iterations actually timed are sized to the budget and the 500-iteration
    rate follows from the measured seconds per iteration (both reported)."""
    import multiprocessing as mp

    from oracle import maps_oracle as mo

    cores = min(host_cores(), 64)
    rng = np.random.default_rng(7)
    B, _ = mo.basis(params, er, elems)
    amps = 10.0 ** rng.uniform(1.5, 3.0, size=(cores, B.shape[1]))
    y = rng.poisson(amps @ B.astype(np.float64).T + w["flat"]).astype(np.float32)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_reference_fit_worker, [(y[i], er, elems, params, 2, w["snip"]) for i in range(cores)], chunksize=1)  # imports
        warm = pool.map(_reference_fit_worker, [(y[i], er, elems, params, 8, w["snip"]) for i in range(cores)], chunksize=1)
        per_it = float(np.mean(warm)) / 8
        n_iter = int(min(500, max(16, budget_s / max(per_it, 1e-4))))
        t0 = time.perf_counter()
        secs = pool.map(_reference_fit_worker, [(y[i], er, elems, params, n_iter, w["snip"]) for i in range(cores)], chunksize=1)
        wall = time.perf_counter() - t0
    s_per_it = float(np.mean(secs)) / n_iter
    return {"px_s_at_500_iterations": cores / (500 * s_per_it), "pixels": cores,
            "processes": cores, "iterations_timed": n_iter, "seconds_per_iteration": s_per_it, "wall_seconds": wall,
            "call": "mapstorch.opt.fit_spec(pixel, energy_range, elements, init_param_vals, tune_params=False, "
                    "init_amp=True, use_snip=%s, optimizer='adam', n_iter=...)" % w["snip"]}


def config1_spectrum():
    """Synthetic integrated spectrum of BASELINE config 1 (2048 channels, 10 K-line elements, 12 keV)."""
    from oracle import maps_oracle as mo

    from mapstorch_b200 import synth

    params = synth.benchmark_params(2048)
    rng = np.random.default_rng(11)
    B, names = mo.basis(params, (0, 2047), synth.ELEMS10)
    amps = 10.0 ** rng.uniform(3.0, 4.5, size=B.shape[1])
    amps[-1] *= 30.0  # a dominant elastic peak, as in measured spectra (SURVEY.md 8d on the scatter-peak heuristic)
    spec = rng.poisson(B.astype(np.float64) @ amps + 40.0).astype(np.float32)
    return spec, params, list(synth.ELEMS10)


def reference_config1():
    """Wall time of the reference's fit_spec with the CLI's arguments on config 1's spectrum (scripts/fit_spec.py)."""
    mods = reference_modules()
    if mods is None:
        return None
    spec, params, elems = config1_spectrum()
    torch.set_num_threads(host_cores())
    t0 = time.perf_counter()
    tensors, spec_fit, bkg, trace = mods[0].fit_spec(spec, [50, 1450], elements_to_fit=elems, init_param_vals=params,
                                                     tune_params=True, init_amp=True, use_snip=True, loss="mse",
                                                     optimizer="adamw", n_iter=500, progress_bar=False, device="cpu")
    return {"seconds": time.perf_counter() - t0, "final_loss": float(trace[-1]), "threads": torch.get_num_threads(),
            "workload": C1_DESC}


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    from mapstorch_b200 import synth

    w = WORKLOADS[args.workload]
    elems = getattr(synth, w["elems"])
    params = synth.benchmark_params(w["n_ch"])
    er = (0, w["n_ch"] - 1)
    extra = {}
    if reference_modules() is not None and not args.no_extra:
        # forked workers first (before this process spins up its own thread pools)
        if not w.get("refine"):
            extra["reference_fit_spec"] = reference_fit_spec_rate(w, params, er, elems)
        extra["reference_config1"] = reference_config1()
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
    line.update(extra)
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


# ---- one workload on the GPU(s) --------------------------------------------------------------------------------------
def accuracy_check(plan, flat, got, params, er, w, device):
    """Amplitudes of what is being timed against a float64 solve of a pixel sample."""
    n_px = flat.shape[0]
    sample = torch.arange(0, n_px, max(1, n_px // 512), device=device)[:512]
    y64 = flat[sample].double()
    if w["snip"]:
        from mapstorch_b200 import bkg as _bkg
        pr = [torch.tensor(float(params[k])) for k in SNIP_KEYS]
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
    return check


def run_workload(name, args, rank, local_rank, world, device, steps, warmup, e2e_steps, with_cpu_baseline,
                 extra_modes=False):
    """Build, check, time one workload; returns the record of its measurements (rank 0's view)."""
    import torch.distributed as dist
    from mapstorch_b200 import _native, opt
    from mapstorch_b200.dist import ShardedMapFit, gather_maps

    w, elems, params, er, plan, truth, vol = build_workload(name, rank, world, device)
    del truth
    flat = vol.reshape(-1, vol.shape[-1])
    n_px = flat.shape[0]
    n_comp = len(plan.names)
    out = torch.zeros((n_comp, n_px), dtype=torch.float32, device=device)
    total_rows = world * w["rows"]
    use_graph = not args.no_graph and not w.get("refine")

    sharded = None
    if world > 1:
        sharded = ShardedMapFit(plan, total_rows, w["width"], gather=args.gather, buffers=2,
                                use_peers=not args.nccl_gather)
        if not sharded.fused:
            print(f"bench: maps gathered with NCCL ({sharded.why_not})", file=sys.stderr)
    graph = None
    if use_graph and world == 1:
        try:
            graph = plan.make_graph(flat, out)
        except Exception as exc:  # pragma: no cover - depends on the driver
            print(f"bench: CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)

    refine_state = {}

    def refine():
        from mapstorch_b200.refine import refine_map
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        refine_state["out"] = refine_map(plan, flat, out, free=w["refine"], max_iter=w["refine_iters"])
        t1.record()
        refine_state.setdefault("events", []).append((t0, t1))

    def step(gather=True, eager=False):
        if sharded is not None and gather:
            # fused: the fit kernels store the amplitudes into the destination GPUs' map buffers over NVLink; the
            # completion of step k is awaited at the start of step k+1 (other buffer) and once more after the loop
            sharded.step(flat, graph=use_graph and sharded.fused and not eager)
            return
        if graph is not None and not eager:
            graph.replay()
        else:
            plan.fit(flat, out=out)
        if w.get("refine"):
            refine()

    def finish():
        if sharded is not None:
            sharded.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(warmup):
        step()
    finish()
    got = plan.fit(flat, out=out).clone()
    check = accuracy_check(plan, flat, got, params, er, w, device)

    # ---- device-resident timing (value) ---------------------------------------------------------
    sampler = ClockSampler(device)
    time.sleep(0.6)  # let nvidia-smi finish its NVML start-up before the timed region
    for _ in range(2):
        step()
    finish()
    launches0 = _native.launch_count
    ms_total = timed(step, steps, after=finish)
    launches_graph = _native.launch_count - launches0
    # same steps launched eagerly (launch count, eager step time) ...
    launches0 = _native.launch_count
    ms_eager = timed(lambda: step(gather=False, eager=True), steps)
    launches = _native.launch_count - launches0
    # ... and once more with CUDA events around the kernels (roofline numerators).  Every probed step is queued behind a
    # 0.25 ms device-side spin, so its launches are in the stream before the GPU reaches the first event: without it the
    # interval of a 0.1 ms kernel (c2) contains the host's launch latency and reads 20 % long.
    plan.timers = []
    refine_state.pop("events", None)
    spin_cycles = int(0.25e-3 * 1.9e9)

    def probe():
        torch.cuda._sleep(spin_cycles)
        step(gather=False, eager=True)

    timed(probe, steps)
    clocks = sampler.stop()
    timers, plan.timers = plan.timers, None
    ms_fit_only = timed(lambda: step(gather=False), steps) if world > 1 else ms_total
    ms_per_step = ms_total / steps
    value = world * n_px / (ms_per_step * 1e-3)

    def kernel_stats(kind):
        sel = [(a.elapsed_time(b), n) for k, a, b, n in timers if k == kind]
        if not sel:
            return None
        per_step_ms = sum(t for t, _ in sel) / steps
        return {"ms_per_step": per_step_ms, "launches_per_step": len(sel) / steps,
                "ms_per_launch": float(np.mean([t for t, _ in sel])), "px_per_launch": float(np.mean([n for _, n in sel]))}

    es = flat.element_size()
    peak, peak_src = measured_peak()
    alg_bytes_px = w["n_ch"] * es + n_comp * 4  # read the spectrum once, write the amplitudes once
    job_frac = alg_bytes_px * n_px / (ms_fit_only / steps * 1e-3) / 1e9 / peak
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(name)
        except Exception:
            traffic = None
    k3 = kernel_stats("contract")
    k2 = kernel_stats("snip")
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    eager_ms = ms_eager / steps
    # what K3 streams: the spectrum, or on the SNIP path the two residual planes (fp16 pair = 4 B per channel, tf32 pair = 8)
    plane_bytes = 4 if getattr(plan, "_last_snip_half", False) else 8
    contract_bytes_px = w["n_ch"] * (es if not w["snip"] else plane_bytes) + n_comp * 4
    k3_roof = None
    if k3:
        achieved = contract_bytes_px * k3["px_per_launch"] / (k3["ms_per_launch"] * 1e-3) / 1e9
        k3_roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                   "traffic": traffic, "kernel": "fit_contract_kernel (tcgen05 + TMA)", "bytes_per_px": contract_bytes_px,
                   "px_per_launch": k3["px_per_launch"], "kernel_ms": k3["ms_per_launch"], "peak_source": peak_src,
                   "share_of_step": k3["ms_per_step"] / eager_ms}
    if w["snip"] and k2:
        # K2 is bound by shared-memory bandwidth, not HBM: passes * C * (2 fetches + 1 store) * 4 B per pixel
        n_pass = int(plan._snip[1])
        smem_bytes_px = n_pass * w["n_ch"] * 12
        achieved = smem_bytes_px * k2["px_per_launch"] / (k2["ms_per_launch"] * 1e-3) / 1e9
        smem_peak = n_sm * 128 * sm_mhz * 1e6 / 1e9
        roofline = {"bound": "smem", "achieved": achieved, "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
                    "traffic": traffic, "kernel": "snip_runs_kernel (K2b)" if getattr(plan._snip[0], "_mt_runs", None) is not None else "snip_kernel (K2)", "bytes_per_px": smem_bytes_px, "passes": n_pass,
                    "px_per_launch": k2["px_per_launch"], "kernel_ms": k2["ms_per_launch"],
                    "peak_source": f"{n_sm} SMs x 128 B/clk x {sm_mhz:.0f} MHz (SM clock sampled during the run)",
                    "share_of_step": k2["ms_per_step"] / eager_ms, "contraction": k3_roof}
    elif w.get("refine") and refine_state.get("events"):
        ev_ms = float(np.mean([a.elapsed_time(b) for a, b in refine_state["events"]]))
        cfg, _, _, _, _, window, _ = next(iter(plan._refine_tables.values()))
        pairs = int(((window >> 16).astype(np.int64) - (window & 0xFFFF).astype(np.int64)).sum())
        evals = float(refine_state["out"][2].float().mean()) + 1.0  # model evaluations per pixel (iterations + the first)
        achieved = pairs * evals * n_px / (ev_ms * 1e-3) / 1e9
        sfu_peak = n_sm * 16 * sm_mhz * 1e6 / 1e9
        roofline = {"bound": "sfu", "achieved": achieved, "peak": sfu_peak, "unit": "Gexp/s", "frac": achieved / sfu_peak,
                    "traffic": traffic, "kernel": "refine_kernel (K4)", "exp_per_px": pairs * evals,
                    "line_channel_pairs": pairs, "evaluations_per_px": evals, "px_per_launch": n_px, "kernel_ms": ev_ms,
                    "peak_source": f"{n_sm} SMs x 16 MUFU.EX2/clk x {sm_mhz:.0f} MHz; algorithmic count = one exp per (line, "
                                   "channel) pair inside the 5.5 sigma windows and model evaluation",
                    "share_of_step": ev_ms / eager_ms, "contraction": k3_roof}
    else:
        roofline = k3_roof
    roofline["job_frac"] = job_frac  # the complete step against the HBM streaming roofline (C*sizeof(in) + E*4 per pixel)

    # ---- multi-GPU: what the timed step assembled must be what NCCL would have gathered, bit for bit ----
    peer_check = None
    if sharded is not None and sharded.fused:
        b = sharded.step(flat, graph=use_graph)
        sharded.finish()
        torch.cuda.synchronize()
        sharded.peer_maps.check()
        ref_full = gather_maps(plan.fit(flat, out=out), total_rows, w["width"])
        res = sharded.result(b)
        owned = sharded.owned()
        same = all(torch.equal(res[j] if sharded.gather == "owner" else res[e], ref_full[e]) for j, e in enumerate(owned))
        flag = torch.tensor([1 if same else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        peer_check = {"bit_identical": bool(int(flag)), "mode": sharded.gather, "graph": bool(use_graph),
                      "components_checked_on_rank0": len(owned), "against": "NCCL all-gather of the same fit"}
        # the other two exchange patterns of the same fit (everything on rank 0 / everything everywhere), short runs:
        # same step, same check; reported beside the headline mode, never instead of it
        exchange_modes = {}
        if extra_modes:
            for mode in ("owner", "root", "all"):
                if mode == sharded.gather:
                    exchange_modes[mode] = {"ms_per_step": ms_total / steps, "value": value, "unit": "px/s",
                                            "bit_identical": peer_check["bit_identical"], "headline": True}
                    continue
                other = ShardedMapFit(plan, total_rows, w["width"], gather=mode, buffers=2)
                ok_all = torch.tensor([1 if other.fused else 0], device=device)
                dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
                if not int(ok_all):
                    exchange_modes[mode] = {"unavailable": other.why_not or "not fused on every rank"}
                    del other
                    continue

                def other_step():
                    other.step(flat, graph=use_graph)

                for _ in range(3):
                    other_step()
                other.finish()
                n_short = max(3, min(steps, 10))
                ms_o = timed(other_step, n_short, after=other.finish) / n_short
                bo = other.step(flat, graph=use_graph)
                other.finish()
                torch.cuda.synchronize()
                other.peer_maps.check()
                res_o = other.result(bo)
                same_o = all(torch.equal(res_o[e], ref_full[e]) for e in other.owned())
                flag_o = torch.tensor([1 if same_o else 0], device=device)
                dist.all_reduce(flag_o, op=dist.ReduceOp.MIN)
                exchange_modes[mode] = {"ms_per_step": ms_o, "value": world * n_px / (ms_o * 1e-3), "unit": "px/s",
                                        "steps": n_short, "bit_identical": bool(int(flag_o))}
                del other, res_o
                torch.cuda.empty_cache()
            peer_check["exchange_modes"] = exchange_modes
        del ref_full

    # ---- end to end through the public API, host buffers ------------------------------------------
    e2e = None
    if e2e_steps > 0:
        previous_affinity = bind_near_gpu(local_rank)
        host = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
        host.copy_(flat)
        torch.cuda.synchronize()

        def e2e_step():
            return opt.fit_map(host.reshape(vol.shape), er, elems, init_param_vals=params, use_snip=w["snip"],
                               device="cpu", refine=w.get("refine"), refine_iters=w.get("refine_iters", 12))

        e2e_step()
        ms = timed(e2e_step, e2e_steps) / e2e_steps
        e2e = {"value": world * n_px / (ms * 1e-3), "unit": "px/s", "h2d_bytes_per_step": world * flat.numel() * es,
               "d2h_bytes_per_step": world * n_comp * n_px * 4, "ms_per_step": ms, "steps": e2e_steps,
               "api": "mapstorch_b200.opt.fit_map(pinned host volume) -> host maps; plan build, chunked H2D "
                      "overlapped with the fit, D2H of the maps all inside the timed region",
               "numa_bound": previous_affinity is not None}
        del host
        if previous_affinity is not None:
            os.sched_setaffinity(0, previous_affinity)  # the CPU baseline below uses every core again

    cb = None
    if with_cpu_baseline and rank == 0 and world == 1 and not w.get("refine"):
        cb = cpu_baseline(w, params, er, elems)

    if sharded is None:
        step_desc = "fit of the resident shard (contraction + non-negative fix-up)"
    elif sharded.fused:
        step_desc = ("fit of the resident shard + element maps assembled over NVLink by the fit kernels' own peer stores "
                     + {"owner": "(every component's full map on its owner GPU: balanced all-to-all)",
                        "root": "(all maps on rank 0)", "all": "(all maps on every GPU)"}[sharded.gather]
                     + "; completion signalled per step (mt_peer_signal), awaited before the next use of the same "
                       "double-buffered destination and after the last step (mt_peer_wait)")
    else:
        step_desc = "fit of the resident shard + NCCL all-gather of the element maps"
    rec = {"value": value, "unit": "px/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "dtype": "f16 x f16 -> f32 (tcgen05)" if w["dtype"] == "f16" else "tf32 x tf32 -> f32 (tcgen05)",
           "config": {"workload": w["desc"], "name": name, "pixels_per_gpu": n_px, "channels": w["n_ch"],
                      "components": n_comp, "storage": w["dtype"], "snip": w["snip"],
                      "l2": f"inputs larger than L2 ({flat.numel() * es / 1e6:.0f} MB per GPU per step), no flush",
                      "step": step_desc},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
           "cuda_graph": (graph is not None) or bool(sharded is not None and sharded.fused and use_graph),
           "gpu_launches_in_graphed_region": launches_graph,
           "ms_per_step_eager_launch": eager_ms, "roofline": roofline, "cpu_baseline": cb,
           "fit_only_px_s": world * n_px / (ms_fit_only / steps * 1e-3),
           "refine": ({"free": list(w["refine"]), "mean_iterations": float(refine_state["out"][2].float().mean()),
                       "mean_loss": float(refine_state["out"][1].mean())} if w.get("refine") else None),
           "check_vs_float64_lstsq": check, "peer_check": peer_check}
    del sharded, graph, vol, flat, out, got
    torch.cuda.empty_cache()
    return rec


def run_config1(device, repeats=3):
    """BASELINE config 1 through opt.fit_spec (global calibration fit of one integrated spectrum, on the GPU)."""
    from mapstorch_b200 import opt

    spec, params, elems = config1_spectrum()
    kw = dict(elements_to_fit=elems, init_param_vals=params, tune_params=True, init_amp=True, use_snip=True, loss="mse",
              optimizer="adamw", n_iter=500, progress_bar=False, device="cuda")
    opt.fit_spec(spec, [50, 1450], **kw)  # warm (tables, kernels)
    torch.cuda.synchronize()
    secs, trace = [], None
    for _ in range(repeats):
        t0 = time.perf_counter()
        tensors, spec_fit, bkg, trace = opt.fit_spec(spec, [50, 1450], **kw)
        torch.cuda.synchronize()
        secs.append(time.perf_counter() - t0)
    from mapstorch_b200 import refine as _refine

    return {"seconds": float(np.median(secs)), "unit": "s per fit_spec call", "repeats": repeats, "final_loss": float(trace[-1]),
            "work": dict(_refine.LAST_STATS),
            "first_loss": float(trace[0]), "workload": C1_DESC,
            "fitted": {k: float(tensors[k]) for k in ("ENERGY_OFFSET", "ENERGY_SLOPE", "FWHM_OFFSET", "COMPTON_ANGLE")
                       if k in tensors}}


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
    ap.add_argument("--no-extra", action="store_true", help="only the headline workload (no c2/c3/c4/config-1 sub-records)")
    ap.add_argument("--nccl-gather", action="store_true", help="gather the maps with NCCL instead of peer-memory stores")
    ap.add_argument("--gather", default="owner", choices=["owner", "root", "all"],
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
    head = run_workload(args.workload, args, rank, local_rank, world, device, args.steps, args.warmup,
                        0 if args.no_e2e else args.e2e_steps, not args.no_cpu_baseline,
                        extra_modes=world > 1 and not args.no_extra and not args.nccl_gather)
    if early_cb is not None:
        head["cpu_baseline"] = early_cb

    # ---- the other BASELINE configurations, short runs (same process, same box) ----
    extra = {}
    if not args.no_extra and args.workload == "c5":
        others = ["c2", "c3", "c4"] if world == 1 else ["c3"]
        for name in others:
            try:
                rec = run_workload(name, args, rank, local_rank, world, device, steps=min(args.steps, 6), warmup=3,
                                   e2e_steps=0 if args.no_e2e else 2, with_cpu_baseline=False)
                extra[name] = {k: rec[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "config", "e2e",
                                                  "roofline", "cuda_graph", "ms_per_step_eager_launch", "fit_only_px_s",
                                                  "refine", "check_vs_float64_lstsq", "peer_check", "gpu_launches")}
            except Exception as exc:  # noqa: BLE001 - a sub-record must not cost the headline line
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
                print(f"bench: workload {name} failed: {exc}", file=sys.stderr)
        if world == 1:
            try:
                extra["c1"] = run_config1(device)
            except Exception as exc:  # noqa: BLE001
                extra["c1"] = {"error": f"{type(exc).__name__}: {exc}"}

    failed = False
    if rank == 0:
        line = {"metric": "pixels/sec", "value": head["value"], "unit": "px/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic", "config": head["config"]}
        for k in ("clocks", "e2e", "gpu_launches", "cuda_graph", "gpu_launches_in_graphed_region", "ms_per_step_eager_launch",
                  "roofline", "cpu_baseline", "fit_only_px_s", "refine", "check_vs_float64_lstsq", "peer_check"):
            line[k] = head[k]
        if extra:
            line["workloads"] = extra
        emit(line)
    checks = [head.get("peer_check")] + [v.get("peer_check") for v in extra.values() if isinstance(v, dict)]
    failed = any(c is not None and not c["bit_identical"] for c in checks)
    if world > 1:
        dist.destroy_process_group()
    if failed:
        print("bench: peer-store maps differ from the NCCL gather", file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
