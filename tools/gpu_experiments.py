#!/usr/bin/env python
"""Ad-hoc GPU experiments (accuracy of the tensor-core accumulation vs K-chunking, e2e breakdown)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from mapstorch_b200 import opt, synth, _native
from oracle import maps_oracle as mo

def acc_experiment():
    elems = synth.ELEMS30
    p = synth.benchmark_params(4096); er = (0, 4095)
    rng = np.random.default_rng(0)
    B, names = mo.basis(p, er, elems)
    n_px = 3000
    amps = 10.0 ** rng.uniform(1.5, 3.0, size=(n_px, len(names)))
    y = rng.poisson(amps @ B.astype(np.float64).T).astype(np.float32)
    plan = opt.MapFitPlan(er, elems, p, use_snip=False)
    ref = mo.fit_lstsq(y, B)
    sc = np.abs(ref).max(axis=1, keepdims=True)
    for dt in (torch.float16, torch.float32):
        dev = torch.from_numpy(y).to('cuda', dt)
        for S in (1, 2, 4, 8, 16, 32):
            step = 4096 // S
            acc = np.zeros((n_px, len(names)))
            x = torch.empty((len(names), n_px), device='cuda')
            for s in range(S):
                plan._contract_untimed(dev, x, 0, 1, s * step, (s + 1) * step, 'tensor', None)
                # zero outside window: need pack zeros -> emulate by restricting k window only (pack has full W)
                acc += x.double().cpu().numpy().T
            e = np.abs(acc - ref) / sc
            per_comp = e.max(axis=0)
            print(dt, 'chunks', S, 'max', e.max(), 'median', np.median(e), 'bias', ((acc - ref) / sc).mean(), 'worst comp', names[int(per_comp.argmax())], flush=True)
    # b = Y B^T formulation, H applied in float64 on host
    class P2: pass
    pinv_saved = plan.pinv
    b64 = plan.basis.double().cpu().numpy()
    plan.pinv = b64; plan._packs = {}
    for dt in (torch.float16, torch.float32):
        dev = torch.from_numpy(y).to('cuda', dt)
        x = torch.empty((len(names), n_px), device='cuda')
        plan._contract_untimed(dev, x, 0, 1, 0, 4096, 'tensor', None)
        b = x.double().cpu().numpy().T
        bref = y.astype(np.float64) @ b64.T
        print(dt, 'b=YB^T rel err max', (np.abs(b - bref) / np.abs(bref)).max(), 'bias', ((b - bref) / bref).mean())
        xx = b @ plan.hinv.T
        e = np.abs(xx - ref) / sc
        print(dt, 'x = H b: max', e.max(), 'median', np.median(e))
    plan.pinv = pinv_saved; plan._packs = {}

def e2e_breakdown():
    w = dict(rows=256, width=2048, n_ch=4096)
    p = synth.benchmark_params(4096); er = (0, 4095)
    t0 = time.perf_counter(); plan = opt.MapFitPlan(er, synth.ELEMS30, p, use_snip=False); torch.cuda.synchronize(); print('plan build s', time.perf_counter() - t0)
    t0 = time.perf_counter(); plan = opt.MapFitPlan(er, synth.ELEMS30, p, use_snip=False); torch.cuda.synchronize(); print('plan build (2nd) s', time.perf_counter() - t0)
    n_px = 524288
    host = torch.empty((n_px, 4096), dtype=torch.float16, pin_memory=True); host.zero_()
    dev = torch.empty((n_px, 4096), dtype=torch.float16, device='cuda')
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('H2D full 4.29 GB: %.1f ms, %.1f GB/s' % (dt * 1e3, 4.295 / dt))
    for chunk in (4096, 16384, 65536):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for p0 in range(0, n_px, chunk):
            dev[p0:p0 + chunk].copy_(host[p0:p0 + chunk], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('H2D chunk %d px: %.1f ms, %.1f GB/s' % (chunk, dt * 1e3, 4.295 / dt))
    x = torch.zeros((32, n_px), device='cuda')
    for name, fn in (('fit(no nnls)', lambda: plan.fit(dev, out=x, nonneg=False)), ('fit(nnls)', lambda: plan.fit(dev, out=x, nonneg=True))):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(); print(name, 'ms', (time.perf_counter() - t0) * 100)
    for chunk in (16384, 65536):
        plan.fit_host(host, chunk_px=chunk); torch.cuda.synchronize(); t0 = time.perf_counter()
        plan.fit_host(host, chunk_px=chunk); torch.cuda.synchronize(); print('fit_host chunk', chunk, 'ms', (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); h = plan.to_host(x); print('to_host first ms', (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); h = plan.to_host(x); print('to_host second ms', (time.perf_counter() - t0) * 1e3)

if __name__ == '__main__':
    for name in sys.argv[1:]:
        globals()[name]()
