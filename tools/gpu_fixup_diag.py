import sys; sys.path.insert(0,'.')
import torch, numpy as np
from mapstorch_b200 import opt, synth
torch.set_printoptions(precision=6, linewidth=200)
p = synth.benchmark_params(4096); er=(0,4095)
plan = opt.MapFitPlan(er, synth.ELEMS30, p, use_snip=False)
truth = synth.truth_maps(len(plan.names), 32, 2048, seed=1234, device='cuda', lo=0.0, hi=2.5)
vol = synth.poisson_volume(plan.basis, truth, dtype=torch.float16, seed=99).reshape(-1,4096)
xu = plan.fit(vol, nonneg=False)
plan.epilogue_fixup = False; ref = plan.fit(vol).clone()
plan.epilogue_fixup = True; nf = torch.zeros(1, dtype=torch.int32, device='cuda'); fused = plan.fit(vol, n_fixed=nf).clone()
print('fallbacks', int(nf), 'of', vol.shape[0])
basis = plan.basis.double(); y = vol.double()
obj = lambda x: ((y - x.double().T @ basis) ** 2).sum(dim=1)
of, orf = obj(fused), obj(ref)
exc = (of - orf)
print('objective excess: max', float(exc.max()), 'min', float(exc.min()), 'rel max', float((exc/orf).max()))
d = (fused-ref).abs() / ref.abs().max(dim=0, keepdim=True).values
print('max diff rel-to-max', float(d.max()))
w = int(exc.argmax())
print('worst pixel', w, 'xu neg comps', (xu[:, w] < 0).nonzero().flatten().tolist())
print('xu   ', xu[:, w])
print('ref  ', ref[:, w])
print('fused', fused[:, w])
G = torch.from_numpy(plan.gram).cuda()
for name, x in (('ref', ref), ('fused', fused)):
    g = G @ (x[:, w].double() - xu[:, w].double())   # gradient of the QP at x
    print(name, 'zero set', (x[:, w] == 0).nonzero().flatten().tolist(), 'min dual on zero set', float(g[x[:, w] == 0].min()) if (x[:, w]==0).any() else None, 'max |grad| on free', float(g[x[:, w] > 0].abs().max()))
