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


class PeerMaps:
    """Element maps of the WHOLE grid assembled over NVLink by the fit kernels themselves.

    The buffer lives in torch symmetric memory, so every rank holds pointers to every other rank's
    copy.  ``plan.fit(shard, out=maps.local, peers=maps.peers)`` makes the K3 epilogue and the
    non-negative fix-up store each amplitude of the rank's shard into the destination copies while
    the contraction is still streaming -- the gather of SURVEY.md 8e without a separate collective;
    ``barrier()`` then orders the remote stores before anybody reads ``maps.full``.

    mode "owner"     : component e is assembled on rank e % world (its "owner"): every rank stores its
                       pixels of component e into the owner's buffer, a balanced all-to-all in which
                       each GPU receives (world-1)/world of ONE share of the maps instead of world-1
                       shares at a single root.  ``owned`` lists this rank's components, ``full`` is
                       [ceil(E/world), P_total] with row j = component owned[j].
    mode "gather"    : every rank stores into rank `root`'s buffer (one peer store per value);
                       only the root ends up with the full maps (root inbound bound on 8 GPUs).
    mode "allgather" : every rank stores into all other ranks' buffers (world-1 peer stores).
    mode "multicast" : one store to the NVSwitch multicast mapping reaches every rank.  Cheap for
                       the coalesced epilogue stores, slow for the scattered fix-up stores, and
                       stores of consecutive kernels are NOT ordered at the destinations (measured:
                       stale values on 8 GPUs) -- kept for experiments only."""

    def __init__(self, n_comp, n_rows, width, group=None, mode="gather", root=0, layout="ep", e_pad=None):
        import torch.distributed._symmetric_memory as symm_mem

        from mapstorch_b200 import _native

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.layout, self.n_comp, self.mode, self.root = layout, n_comp, mode, root
        device = torch.device("cuda", torch.cuda.current_device())
        r0, r1 = shard_rows(n_rows, self.rank, self.world)
        self.owned = list(range(n_comp))
        if mode == "owner":
            if layout != "ep" or n_comp > 32:
                raise NotImplementedError("owner mode: element-major layout, at most 32 components")
            n_rows_own = -(-n_comp // self.world)
            self.owned = [e for e in range(n_comp) if e % self.world == self.rank]
            self.full = symm_mem.empty((n_rows_own, n_rows * width), dtype=torch.float32, device=device)
            self.local = torch.zeros((n_comp, (r1 - r0) * width), dtype=torch.float32, device=device)
            offset = 4 * r0 * width
        elif layout == "pe":
            e_pad = e_pad or -(-n_comp // 16) * 16
            self.full = symm_mem.empty((n_rows * width, e_pad), dtype=torch.float32, device=device)
            self.local = self.full[r0 * width: r1 * width]
            offset = 4 * r0 * width * e_pad
        else:
            self.full = symm_mem.empty((n_comp, n_rows * width), dtype=torch.float32, device=device)
            self.local = self.full[:, r0 * width: r1 * width]
            offset = 4 * r0 * width
        self.full.zero_()
        self.handle = symm_mem.rendezvous(self.full, group)
        # every rank's zero-fill must have completed before any peer starts storing into the buffer
        self.handle.barrier()
        self.peers = _native.MtPeers()
        self.peers.n_peers = 0
        self.peers.multicast = None
        self.multicast = False
        self.peers.n_owned = 0
        if mode == "owner":
            self.peers.n_owned = n_comp
            row_bytes = 4 * n_rows * width
            for e in range(n_comp):
                owner, slot = e % self.world, e // self.world
                self.peers.owner_row[e] = int(self.handle.buffer_ptrs[owner]) + slot * row_bytes + offset
        elif mode == "multicast":
            mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
            if not mc:
                raise RuntimeError("this system has no NVLink multicast support")
            self.peers.multicast = mc + offset
            self.multicast = True
        else:
            if mode == "gather":
                targets = [root] if self.rank != root else []
            elif mode == "allgather":
                targets = [r for r in range(self.world) if r != self.rank]
            else:
                raise ValueError(f"unknown PeerMaps mode {mode!r}")
            self.peers.n_peers = len(targets)
            for k, r in enumerate(targets):
                self.peers.peer[k] = int(self.handle.buffer_ptrs[r]) + offset

    def fit_chunked(self, plan, spec2d, rounds_per_chunk=4, **fit_kw):
        """Fit the rank's shard in chunks of whole waves of the persistent kernel and copy every
        finished chunk of the maps to the root's buffer on a side stream, so that the NVLink
        traffic of chunk k overlaps the contraction of chunk k+1.  The kernels store locally only
        (no scattered remote stores from the fix-up), the copies are coalesced."""
        if self.layout != "ep":
            raise NotImplementedError("chunked gather uses the element-major layout")
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream()
            self._root_view = self.handle.get_buffer(self.root, tuple(self.full.shape), torch.float32)
        n_px = spec2d.shape[0]
        sms = torch.cuda.get_device_properties(spec2d.device).multi_processor_count
        chunk = sms * 256 * rounds_per_chunk
        g0 = self.local.storage_offset() - self.full.storage_offset()  # first pixel of this rank in the full map
        for p0 in range(0, n_px, chunk):
            n = min(chunk, n_px - p0)
            plan.fit(spec2d[p0: p0 + n], out=self.local[:, p0: p0 + n], **fit_kw)
            if self.rank != self.root:
                done = torch.cuda.Event()
                done.record(main)
                self._side.wait_event(done)
                with torch.cuda.stream(self._side):
                    self._root_view[:, g0 + p0: g0 + p0 + n].copy_(self.local[:, p0: p0 + n], non_blocking=True)
        if self.rank != self.root:
            joined = torch.cuda.Event()
            joined.record(self._side)
            main.wait_event(joined)

    def has_full_maps(self):
        return self.mode in ("allgather", "multicast") or (self.mode == "gather" and self.rank == self.root)

    def component(self, index, n_rows, width):
        """[n_rows, width] map of one component (a strided view for the pixel-major layout)."""
        flat = self.full[:, index] if self.layout == "pe" else self.full[index]
        return flat.reshape(n_rows, width)

    def barrier(self):
        """All ranks' stores issued so far on the current stream are visible everywhere afterwards."""
        self.handle.barrier()
