"""Row-wise sharding of the pixel grid across the GPUs of one box and the gather of element maps.

Pixels are independent given the global parameters, so the fit itself needs no
collective: rank r owns a contiguous block of rows of the [H, W, C] volume, the basis / Gram
matrix / index tables are replicated (<= 1.5 MB), and only the small element maps
([E, P_local] fp32, 128 bytes per pixel at E = 32 against 8-16 KB of spectrum) cross NVLink.
One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_rows(n_rows, rank, world):
    """Half-open row range [r0, r1) of rank `rank`: blocks differ by at most one row."""
    base, extra = divmod(int(n_rows), int(world))
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def shard_sizes(n_rows, world):
    return [b - a for a, b in (shard_rows(n_rows, r, world) for r in range(world))]


def gather_maps(x_local, n_rows, width, group=None):
    """All-gather per-rank element maps x_local [E, rows_local * width] into [E, n_rows * width] on
    every rank, laid out so that out[e].reshape(n_rows, width) is the full map of component e.

    Each component is gathered straight into its final place (one all-gather per component, issued
    as one coalesced group on NCCL), so no permute pass over the result is needed.  Uneven shards
    are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_comp = x_local.shape[0]
    sizes = [s * width for s in shard_sizes(n_rows, world)]
    assert x_local.shape[1] == sizes[rank], (x_local.shape, sizes[rank])
    if world == 1:
        return x_local
    even = len(set(sizes)) == 1
    pad = max(sizes)
    src = x_local.contiguous()
    if not even and sizes[rank] != pad:
        src = torch.zeros((n_comp, pad), dtype=x_local.dtype, device=x_local.device)
        src[:, : sizes[rank]] = x_local
    out = torch.empty((n_comp, world * pad), dtype=x_local.dtype, device=x_local.device)

    def issue():
        for e in range(n_comp):
            dist.all_gather_into_tensor(out[e], src[e], group=group)

    if x_local.is_cuda and hasattr(dist, "_coalescing_manager"):
        with dist._coalescing_manager(group=group, device=x_local.device, async_ops=False):
            issue()
    else:
        issue()
    if even:
        return out
    blocks = out.reshape(n_comp, world, pad)
    return torch.cat([blocks[:, r, : sizes[r]] for r in range(world)], dim=1)


def allreduce_spectrum(int_spec_local, group=None):
    """Sum of the per-rank integrated spectra (the MAPS/int_spec of io.create_dataset, io.py:159)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(int_spec_local, group=group)
    return int_spec_local
