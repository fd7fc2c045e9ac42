import sys; sys.path.insert(0,'.')
import torch, numpy as np
from mapstorch_b200 import opt, synth
p = synth.benchmark_params(4096); er=(0,4095)
plan = opt.MapFitPlan(er, synth.ELEMS30, p, use_snip=False)
truth = synth.truth_maps(len(plan.names), 64, 2048, seed=1234, device='cuda')
vol = synth.poisson_volume(plan.basis, truth, dtype=torch.float16, seed=99).reshape(-1,4096)
xu = plan.fit(vol, nonneg=False)
neg = (xu < 0).sum(dim=0)
print('pixels', vol.shape[0], 'flagged frac', float((neg>0).float().mean()), 'hist of #neg', torch.bincount(neg.clamp(max=8)).tolist())
nf = torch.zeros(1, dtype=torch.int32, device='cuda')
plan.fit(vol, n_fixed=nf); print('fallback listed', int(nf))
x = plan.fit(vol); zeros = (x==0).sum(dim=0); print('hist of final zero-set size', torch.bincount(zeros.clamp(max=8)).tolist())
import time
for fix in (True, False):
    plan.epilogue_fixup = fix
    out = torch.zeros_like(xu)
    for _ in range(3): plan.fit(vol, out=out)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(20): plan.fit(vol, out=out)
    torch.cuda.synchronize(); print('epilogue_fixup', fix, 'ms', (time.perf_counter()-t)*50)
