"""Multi-GPU parity (needs >= 2 GPUs): the element maps assembled by the fit kernels' own peer stores
(dist.PeerMaps / dist.ShardedMapFit, every gather mode, eager and CUDA-graph launches, double buffering), the
NCCL gather and the public ``opt.fit_map(group=...)`` must all equal the unsharded single-GPU fit BIT FOR BIT --
pixels are independent, so sharding may not change a single value."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ELEMS10 = ["K", "Ca", "Ti", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
ELEMS30 = ("Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge "
           "Ba_L Ce_L Gd_L W_L Pt_L Au_L Au_M Pt_M Hg_M W_M Si_Si").split()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _volume(plan, rows, width, seed, low=False):
    """the same synthetic counts on every rank (seeded on the CPU)"""
    rng = np.random.default_rng(seed)
    basis = plan.basis.double().cpu().numpy()
    lo, hi = (-0.5, 1.5) if low else (1.0, 3.0)
    amps = 10.0 ** rng.uniform(lo, hi, size=(rows * width, basis.shape[0]))
    return torch.from_numpy(rng.poisson(amps @ basis + 2.0).astype(np.float32)).reshape(rows, width, -1)


def _worker(rank, world, port):
    import torch.distributed as dist

    from mapstorch_b200 import dist as mdist, opt
    from mapstorch_b200.default import default_param_vals

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        p = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0, ENERGY_SLOPE=0.00599, ENERGY_QUADRATIC=0.0)
        er = (0, 4095)
        rows, width = 7, 96  # uneven row blocks: 4 + 3
        plan = opt.MapFitPlan(er, ELEMS30, p, use_snip=False)
        vol = _volume(plan, rows, width, seed=5, low=True).half().cuda()  # low counts: many non-negative fix-ups
        whole = plan.fit(vol.reshape(-1, 4096)).clone()  # [E, P] unsharded, on every rank
        assert int((whole == 0).sum()) > 0
        r0, r1 = mdist.shard_rows(rows, rank, world)
        shard = vol[r0:r1].reshape(-1, 4096)
        # NCCL gather of the local fit
        local = plan.fit(shard).clone()
        assert torch.equal(local, whole[:, r0 * width: r1 * width])
        assert torch.equal(mdist.gather_maps(local, rows, width), whole)
        # fused exchange, every mode, eager and graphed, several steps through both buffers
        for gather in ("owner", "all", "root"):
            for graph in (False, True):
                sh = mdist.ShardedMapFit(plan, rows, width, gather=gather, buffers=2)
                assert sh.fused, sh.why_not
                for step in range(3):
                    b = sh.step(shard, graph=graph)
                    sh.finish()
                    torch.cuda.synchronize()
                    sh.peer_maps.check()
                    got = sh.result(b)
                    owned = sh.owned()
                    if gather == "owner":
                        assert owned == [e for e in range(len(plan.names)) if e % world == rank]
                    for j, e in enumerate(owned):
                        row = got[j] if gather == "owner" else got[e]
                        assert torch.equal(row, whole[e]), (gather, graph, step, e)
                    got.zero_()  # the next use of this buffer must refill it completely
                    dist.barrier()
                maps = None
        # the general solver path (no in-epilogue fix-up) pushes its pixels to the owners as well
        plan.epilogue_fixup = False
        sh = mdist.ShardedMapFit(plan, rows, width, gather="owner", buffers=1)
        b = sh.step(shard)
        sh.finish()
        torch.cuda.synchronize()
        plan.epilogue_fixup = True
        ref = whole
        plan2 = opt.MapFitPlan(er, ELEMS30, p, use_snip=False)
        plan2.epilogue_fixup = False
        ref2 = plan2.fit(vol.reshape(-1, 4096))
        for j, e in enumerate(sh.owned()):
            assert torch.equal(sh.result(b)[j], ref2[e]), e
        # public entry point, with and without SNIP, fp32 and fp16
        for use_snip, dtype in ((False, torch.float16), (True, torch.float32)):
            full = opt.fit_map(vol.to(dtype), er, ELEMS30, init_param_vals=p, use_snip=use_snip)
            for gather in ("all", "owner", "root"):
                part = opt.fit_map(vol[r0:r1].to(dtype), er, ELEMS30, init_param_vals=p, use_snip=use_snip,
                                   group=dist.group.WORLD, gather=gather)
                if gather == "all":
                    assert set(part) == set(full)
                if gather == "root":
                    assert (len(part) == len(full)) if rank == 0 else (len(part) == 0)
                for name, m in part.items():
                    assert m.shape == (rows, width) and torch.equal(m, full[name]), (use_snip, gather, name)
        # 12 components through the 2048-channel fp32 path, twice in a row (buffer reuse across calls)
        p10 = dict(default_param_vals, COHERENT_SCT_ENERGY=12.0)
        plan10 = opt.MapFitPlan((0, 2047), ELEMS10, p10, use_snip=False)
        for seed in (1, 2):
            v = _volume(plan10, 6, 40, seed=seed).cuda()
            full = opt.fit_map(v, (0, 2047), ELEMS10, init_param_vals=p10, use_snip=False)
            a0, a1 = mdist.shard_rows(6, rank, world)
            part = opt.fit_map(v[a0:a1], (0, 2047), ELEMS10, init_param_vals=p10, use_snip=False, group=dist.group.WORLD)
            for name in full:
                assert torch.equal(part[name], full[name]), name
        # per-pixel refinement and the L1 objective in sharded form: the rank's block through the single-GPU call, maps
        # (incl. calibration maps, loss, iterations) by NCCL; every gather mode
        v = _volume(plan10, 6, 40, seed=3).cuda()
        a0, a1 = mdist.shard_rows(6, rank, world)
        for kw in (dict(refine=("ENERGY_OFFSET", "FWHM_OFFSET"), refine_iters=8), dict(loss="l1"),
                   dict(loss="l1", l1_lambda=1e-3)):
            full = opt.fit_map(v, (0, 2047), ELEMS10, init_param_vals=p10, use_snip=False, **kw)
            for gather in ("all", "owner", "root"):
                part = opt.fit_map(v[a0:a1], (0, 2047), ELEMS10, init_param_vals=p10, use_snip=False,
                                   group=dist.group.WORLD, gather=gather, **kw)
                keys = list(full)
                want = keys if gather == "all" else ([k for i, k in enumerate(keys) if i % world == rank]
                                                     if gather == "owner" else (keys if rank == 0 else []))
                assert list(part) == want, (kw, gather)
                for name, m in part.items():
                    assert m.dtype == full[name].dtype and torch.equal(m, full[name]), (kw, gather, name)
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_fit_equals_the_unsharded_fit_bit_for_bit():
    import torch.multiprocessing as mp

    os.environ.setdefault("NCCL_DEBUG", "WARN")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
