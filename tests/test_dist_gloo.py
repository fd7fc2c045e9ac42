"""World-size-2 (and 3) CPU runs of the multi-GPU host logic over gloo: row sharding and the
element-map gather layout.  The per-rank fit itself is covered by the shard-equivalence GPU test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mapstorch_b200.dist import gather_maps, shard_rows, shard_sizes, allreduce_spectrum


def test_shard_rows_cover_the_grid():
    for n_rows in (1, 7, 8, 256, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n_rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n_rows, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, width, n_comp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_comp * n_rows * width, dtype=torch.float32).reshape(n_comp, n_rows, width)
        r0, r1 = shard_rows(n_rows, rank, world)
        local = full[:, r0:r1].reshape(n_comp, -1).clone()
        out = gather_maps(local, n_rows, width)
        assert out.shape == (n_comp, n_rows * width)
        assert torch.equal(out.reshape(n_comp, n_rows, width), full)
        spec = torch.full((16,), float(rank + 1))
        assert torch.equal(allreduce_spectrum(spec), torch.full((16,), float(sum(range(1, world + 1)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 8), (2, 7), (3, 8)])
def test_gather_maps_layout_gloo(world, n_rows):
    mp.spawn(_worker, args=(world, _free_port(), n_rows, 5, 4), nprocs=world, join=True)
