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
                       pixels of component e into the owner's buffer (and ONLY there: ``local`` is scratch
                       for the non-negative solver in this mode), a balanced all-to-all in which
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

    def __init__(self, n_comp, n_rows, width, group=None, mode="gather", root=0, layout="ep", e_pad=None, buffers=1):
        """buffers=2 double-buffers the destination maps: step k stores into buffer k % 2, so its completion
        (signal / wait below) may overlap step k+1, which stores into the other buffer."""
        import torch.distributed._symmetric_memory as symm_mem

        from mapstorch_b200 import _native

        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.layout, self.n_comp, self.mode, self.root = layout, n_comp, mode, root
        self.n_rows, self.width, self.buffers, self.current = n_rows, width, int(buffers), 0
        device = torch.device("cuda", torch.cuda.current_device())
        r0, r1 = shard_rows(n_rows, self.rank, self.world)
        self.rows = (r0, r1)
        self.owned = list(range(n_comp))
        nb = self.buffers
        if mode == "owner":
            if layout != "ep" or n_comp > 32:
                raise NotImplementedError("owner mode: element-major layout, at most 32 components")
            n_rows_own = -(-n_comp // self.world)
            self.owned = [e for e in range(n_comp) if e % self.world == self.rank]
            self._all = symm_mem.empty((nb, n_rows_own, n_rows * width), dtype=torch.float32, device=device)
            self.fulls = [self._all[b] for b in range(nb)]
            self.locals = [torch.zeros((n_comp, (r1 - r0) * width), dtype=torch.float32, device=device) for _ in range(nb)]
            offset = 4 * r0 * width
        elif layout == "pe":
            e_pad = e_pad or -(-n_comp // 16) * 16
            self._all = symm_mem.empty((nb, n_rows * width, e_pad), dtype=torch.float32, device=device)
            self.fulls = [self._all[b] for b in range(nb)]
            self.locals = [f[r0 * width: r1 * width] for f in self.fulls]
            offset = 4 * r0 * width * e_pad
        else:
            self._all = symm_mem.empty((nb, n_comp, n_rows * width), dtype=torch.float32, device=device)
            self.fulls = [self._all[b] for b in range(nb)]
            self.locals = [f[:, r0 * width: r1 * width] for f in self.fulls]
            offset = 4 * r0 * width
        self._all.zero_()
        self.handle = symm_mem.rendezvous(self._all, group)
        # completion flags of the fused exchange (mt_peer_signal / mt_peer_wait): one uint32 per source rank
        self._flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self._flags.zero_()
        self._flag_handle = symm_mem.rendezvous(self._flags, group)
        self._epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self._timed_out = torch.zeros(1, dtype=torch.int32, device=device)
        self.sync = _native.MtPeerSync()
        self.sync.n_ranks, self.sync.rank = self.world, self.rank
        for r in range(self.world):
            self.sync.flags[r] = int(self._flag_handle.buffer_ptrs[r])
        # every rank's zero-fill must have completed before any peer starts storing into the buffers
        torch.cuda.synchronize()
        self.handle.barrier()
        self._flag_handle.barrier()
        torch.cuda.synchronize()
        buf_bytes = self._all[0].numel() * 4
        self.peers_list = []
        self.multicast = False
        for b in range(nb):
            peers = _native.MtPeers()
            peers.n_peers = 0
            peers.multicast = None
            peers.n_owned = 0
            base = b * buf_bytes
            if mode == "owner":
                peers.n_owned = n_comp
                row_bytes = 4 * n_rows * width
                for e in range(n_comp):
                    owner, slot = e % self.world, e // self.world
                    peers.owner_row[e] = int(self.handle.buffer_ptrs[owner]) + base + slot * row_bytes + offset
            elif mode == "multicast":
                mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
                if not mc:
                    raise RuntimeError("this system has no NVLink multicast support")
                peers.multicast = mc + base + offset
                self.multicast = True
            else:
                if mode == "gather":
                    targets = [root] if self.rank != root else []
                elif mode == "allgather":
                    targets = [r for r in range(self.world) if r != self.rank]
                else:
                    raise ValueError(f"unknown PeerMaps mode {mode!r}")
                peers.n_peers = len(targets)
                for k, r in enumerate(targets):
                    peers.peer[k] = int(self.handle.buffer_ptrs[r]) + base + offset
            self.peers_list.append(peers)

    # the buffer the next fit stores into
    @property
    def full(self):
        return self.fulls[self.current]

    @property
    def local(self):
        return self.locals[self.current]

    @property
    def peers(self):
        return self.peers_list[self.current]

    def flip(self):
        """Make the other buffer the target of the next fit (no-op with a single buffer); returns the index of the
        buffer that was current."""
        was = self.current
        self.current = (self.current + 1) % self.buffers
        return was

    def signal(self, stream=None):
        """Publish "everything this rank has stored so far (on `stream`) has landed" to every rank."""
        import ctypes
        import os

        from mapstorch_b200 import _native

        if os.environ.get("MT_PEER_SYNC", "") == "none":  # measurement knob: what the signalling costs
            return

        _native.call("mt_peer_signal", ctypes.byref(self.sync), self._epoch.data_ptr(), _native.stream_ptr(stream))

    def wait(self, lag=0, stream=None):
        """Block `stream` until every rank has signalled (this rank's count - lag) times."""
        import ctypes
        import os

        from mapstorch_b200 import _native

        if os.environ.get("MT_PEER_SYNC", "") in ("none", "signal"):
            return

        _native.call("mt_peer_wait", ctypes.byref(self.sync), self._epoch.data_ptr(), int(lag),
                     self._timed_out.data_ptr(), _native.stream_ptr(stream))

    def check(self):
        """Raise if a wait expired (a peer never signalled).  Synchronises."""
        silent = int(self._timed_out.item())
        if silent:
            raise RuntimeError(f"PeerMaps: rank {silent - 1} never signalled completion of its stores")

    def fit_chunked(self, plan, spec2d, rounds_per_chunk=4, **fit_kw):
        """Fit the rank's shard in chunks of whole waves of the persistent kernel and copy every
        finished chunk of the maps to the root's buffer on a side stream, so that the NVLink
        traffic of chunk k overlaps the contraction of chunk k+1.  The kernels store locally only
        (no scattered remote stores from the fix-up), the copies are coalesced."""
        if self.layout != "ep" or self.buffers != 1:
            raise NotImplementedError("chunked gather uses the element-major layout and a single buffer")
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream()
            self._root_view = self.handle.get_buffer(self.root, tuple(self.full.shape), torch.float32)
        n_px = spec2d.shape[0]
        sms = torch.cuda.get_device_properties(spec2d.device).multi_processor_count
        chunk = sms * 256 * rounds_per_chunk
        g0 = self.rows[0] * self.width  # first pixel of this rank in the full map
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

    def barrier(self, native=True):
        """All ranks' stores issued so far on the current stream are visible everywhere afterwards."""
        if native:
            self.signal()
            self.wait()
        else:
            self.handle.barrier()


class ShardedMapFit:
    """The multi-GPU map fit as one object: every rank fits its row shard, and the element maps of the whole grid
    are assembled over NVLink by the fit kernels' own peer stores (PeerMaps) -- or, where those do not apply
    (more than 32 components, components without an open line, the l1_lambda penalty, no symmetric memory), by a
    NCCL all-gather.  ``bench.py`` and ``opt.fit_map(..., group=...)`` both run through this class.

    gather: "owner" (component e's full map on rank e % world: the balanced exchange), "all" (every map on every
    rank) or "root" (every map on rank 0).

    step(shard2d) launches fit + signal into the current buffer (with two buffers followed by a wait for everybody's
    PREVIOUS step) and flips to the other buffer; finish() waits until this rank holds every peer's share of all
    steps launched so far.  With two buffers the ranks therefore run with one step of slack: the completion latency
    of step k hides behind the fit of step k+1."""

    MODES = {"owner": "owner", "all": "allgather", "root": "gather"}

    def __init__(self, plan, n_rows, width, group=None, gather="owner", buffers=2, use_peers=True):
        if gather not in self.MODES:
            raise ValueError(f"gather must be one of {sorted(self.MODES)}, got {gather!r}")
        self.plan, self.n_rows, self.width, self.gather = plan, int(n_rows), int(width), gather
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.rows = shard_rows(self.n_rows, self.rank, self.world)
        self.n_comp = len(plan.names)
        self.peer_maps, self.why_not = None, None
        self._graphs = {}
        self._nccl_out = None
        self._last = None
        dense = plan.n_live == self.n_comp and len(plan.groups) == 1
        if not use_peers:
            self.why_not = "disabled by the caller"
        elif not dense:
            self.why_not = "components without an open line or more than one tensor-core group"
        elif gather == "owner" and self.n_comp > 32:
            self.why_not = "owner exchange holds at most 32 components"
        else:
            try:
                self.peer_maps = PeerMaps(self.n_comp, self.n_rows, self.width, group=self.group,
                                          mode=self.MODES[gather], buffers=buffers)
            except Exception as exc:  # no symmetric memory / NVLink on this system
                self.why_not = f"symmetric memory unavailable: {exc}"
        if self.peer_maps is None:
            n_local = (self.rows[1] - self.rows[0]) * self.width
            self._local = torch.zeros((self.n_comp, n_local), dtype=torch.float32, device="cuda")

    @property
    def fused(self):
        return self.peer_maps is not None

    def step(self, shard2d, graph=False, **fit_kw):
        """Fit the rank's shard [P_local, C] (CUDA) and start the exchange.  Returns the index of the buffer that
        receives this step's maps.  graph=True replays a CUDA graph of (fit + signal) captured on first use."""
        pm = self.peer_maps
        if pm is None:
            self.plan.fit(shard2d, out=self._local, **fit_kw)
            self._nccl_out = gather_maps(self._local, self.n_rows, self.width, group=self.group)
            self._last = 0
            return 0
        b = pm.current
        if graph:
            key = (b, shard2d.data_ptr(), tuple(shard2d.shape), tuple(sorted(fit_kw)))
            if key not in self._graphs:
                self._graphs[key] = self._capture(shard2d, fit_kw)
            self._graphs[key].replay()
        else:
            self.plan.fit(shard2d, out=pm.local, peers=pm.peers, **fit_kw)
            pm.signal()
            self._trailing_wait()
        self._last = b
        pm.flip()
        return b

    def _capture(self, shard2d, fit_kw):
        pm, plan = self.peer_maps, self.plan
        saved, plan.timers = plan.timers, None
        # warm run outside the capture (operand packs, tables, the count-range check of fp32 volumes); its signal
        # keeps every rank's epoch in step with the eager path
        plan.fit(shard2d, out=pm.local, peers=pm.peers, **fit_kw)
        pm.signal()
        self._trailing_wait()
        torch.cuda.synchronize()
        kw = dict(fit_kw)
        if shard2d.dtype == torch.float32 and "exact_counts" not in kw:  # no host sync inside the capture
            if plan.use_snip:
                kw["exact_counts"] = True if plan._last_snip_half else "split"
            else:
                kw["exact_counts"] = "split" if plan._last_wide else True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan.fit(shard2d, out=pm.local, peers=pm.peers, **kw)
            pm.signal()
            self._trailing_wait()
        plan.timers = saved
        return g

    def _trailing_wait(self):
        """With two buffers a step ends by waiting for everybody's PREVIOUS step: the maps of step k-1 (other buffer)
        are then complete on this rank, and nobody can still be storing into the buffer that step k+1 reuses -- the
        ranks run with one step of slack instead of in lock-step.  A rank that reads maps(b) between steps on the same
        stream is ordered before its next signal, hence before any peer overwrites that buffer two steps later."""
        if self.peer_maps.buffers >= 2:
            self.peer_maps.wait(lag=1)

    def finish(self):
        """Block the current stream until the maps of every step launched so far are complete on this rank."""
        if self.peer_maps is not None:
            self.peer_maps.wait()

    def result(self, b=None):
        """Tensor holding the gathered maps of buffer b (default: the last step's): [n_owned, P] for "owner"
        (row j = component owned[j]), [E, P] otherwise (complete on every rank for "all", on rank 0 for "root")."""
        if self.peer_maps is None:
            return self._nccl_out
        return self.peer_maps.fulls[self._last if b is None else b]

    def owned(self):
        """Indices of the components whose complete map this rank holds after finish()."""
        if self.peer_maps is None or self.gather == "all":
            return list(range(self.n_comp))
        if self.gather == "root":
            return list(range(self.n_comp)) if self.rank == 0 else []
        return list(self.peer_maps.owned)

    def maps(self, b=None):
        """{name: [n_rows, width] map} of the components this rank holds in full."""
        full = self.result(b)
        if self.peer_maps is not None and self.gather == "owner":
            return {self.plan.names[e]: full[j].reshape(self.n_rows, self.width) for j, e in enumerate(self.peer_maps.owned)}
        return {self.plan.names[e]: full[e].reshape(self.n_rows, self.width) for e in self.owned()}


_SHARDED_CACHE = {}


def fit_map_sharded(spec_shard, energy_range, elements_to_fit=None, group=None, gather="all", plan=None, **fit_map_kw):
    """``opt.fit_map`` for a grid that is row-sharded over the ranks of `group`: `spec_shard` is THIS rank's block of
    rows [rows_local, W, C] (blocks as dist.shard_rows cuts them), the result is ``{name: [H_total, W]}`` for the
    components this rank holds in full -- all of them with gather="all", the rank's own with "owner" (component e
    lives on rank e % world), all on rank 0 and none elsewhere with "root".  The fit needs no collective (pixels
    are independent); the element maps travel by peer stores fused into the fit kernels."""
    from mapstorch_b200 import opt
    from mapstorch_b200.default import default_fitting_elems

    group = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    if spec_shard.dim() != 3:
        raise ValueError("fit_map_sharded expects this rank's [rows_local, W, C] block of the volume")
    rows_local, width = int(spec_shard.shape[0]), int(spec_shard.shape[1])
    counts = torch.tensor([rows_local, width], dtype=torch.int64, device="cuda")
    every = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(every, counts, group=group)
    rows = [int(t[0]) for t in every]
    n_rows = sum(rows)
    if any(int(t[1]) != width for t in every) or rows != shard_sizes(n_rows, world):
        raise ValueError(f"row blocks {rows} (width {[int(t[1]) for t in every]}) are not the shard_rows partition of {n_rows} rows")
    elements = default_fitting_elems if elements_to_fit is None else elements_to_fit
    supported = {k: fit_map_kw.pop(k) for k in list(fit_map_kw) if k in ("init_param_vals", "fixed_param_vals", "indices",
                 "use_snip", "e_consts", "elem_info", "detector_type", "escape_factor")}
    l1_lambda = float(fit_map_kw.pop("l1_lambda", 0.0))
    nonneg = bool(fit_map_kw.pop("nonneg", True))
    loss = fit_map_kw.pop("loss", "mse")
    refine = fit_map_kw.pop("refine", None)
    refine_kw = {k: fit_map_kw.pop(k) for k in list(fit_map_kw) if k in ("refine_iters", "refine_bounds")}
    if fit_map_kw:
        raise TypeError(f"fit_map_sharded does not take {sorted(fit_map_kw)}")
    if refine or loss != "mse":
        # Per-pixel refinement and the L1 objective produce more than amplitudes (calibration maps, loss, iteration
        # counts) and run their own kernels after the contraction: every rank fits its block with the single-GPU call
        # and the maps are assembled by one NCCL all-gather (what the fused exchange is checked against); entry i of
        # the result lives on rank i % world with gather="owner", everything on rank 0 with "root".
        local = opt.fit_map(spec_shard, energy_range, elements, l1_lambda=l1_lambda, nonneg=nonneg, loss=loss,
                            refine=refine, device="cuda", **supported, **refine_kw)
        keys = list(local)
        stacked = torch.stack([local[k].reshape(-1).to(torch.float32) for k in keys])
        full = gather_maps(stacked, n_rows, width, group=group)
        rank = dist.get_rank(group)
        out = {}
        for i, k in enumerate(keys):
            if gather == "all" or (gather == "root" and rank == 0) or (gather == "owner" and i % world == rank):
                out[k] = full[i].reshape(n_rows, width).to(local[k].dtype)
        return out
    if plan is None:
        params = dict(supported.get("init_param_vals", opt.default_param_vals))
        params.update(supported.get("fixed_param_vals", {}))
        plan = opt._cached_plan(energy_range, elements, params, supported.get("indices"), supported.get("use_snip", True),
                                supported.get("e_consts", opt.default_energy_consts), supported.get("elem_info", opt.default_elem_info),
                                supported.get("detector_type", "Si"), supported.get("escape_factor", 0.0))
    key = (id(plan), n_rows, width, gather, id(group))
    sharded = _SHARDED_CACHE.get(key)
    if sharded is None:
        if len(_SHARDED_CACHE) >= 4:
            _SHARDED_CACHE.pop(next(iter(_SHARDED_CACHE)))
        sharded = _SHARDED_CACHE[key] = ShardedMapFit(plan, n_rows, width, group=group, gather=gather, buffers=1,
                                                      use_peers=not l1_lambda > 0)
    flat = spec_shard.reshape(-1, spec_shard.shape[-1])
    if flat.dtype not in (torch.float32, torch.float16):
        flat = flat.float()
    if flat.device.type != "cuda":
        flat = flat.cuda(non_blocking=True)
    kw = dict(nonneg=nonneg)
    if l1_lambda > 0:
        kw["l1_lambda"] = l1_lambda
    if sharded.fused:
        sharded.peer_maps.barrier()  # every rank has finished reading the buffers of the previous call
    sharded.step(flat, **kw)
    sharded.finish()
    out = {name: m.clone() for name, m in sharded.maps().items()}  # the buffers are reused by the next call
    if sharded.fused:
        sharded.peer_maps.check()
    return out
