"""Element-sharded meshes across the GPUs of one node (no reference counterpart --
the reference is a single CPU process).

One process per GPU.  Rank ``p`` owns a contiguous block of elements ``[e_p, e_{p+1})`` and a contiguous
range of vertex rows ``[v_p, v_{p+1})`` with ``v_p`` = smallest vertex id its elements reference, so every
vertex an element touches is owned by its own rank or a HIGHER one.  For rank ``q``:

* own elements                 evaluated here;
* ghost vertices               referenced by own elements, owned by a higher rank.  ``interface="exchange"``: the
                               partial gradient rows and Hessian block-rows computed here travel to the owner
                               (one NCCL send per neighbour per assembly) and are added there, in rank order;
* pattern-only elements        lower ranks' elements that touch vertices owned here.  ``interface="recompute"``
                               (default): this rank evaluates them too, so its owned rows are complete without any
                               communication; ``"exchange"``: they are never evaluated and only reserve CSR slots
                               (``skb_plan_create_sharded``).  Either way the rows this rank owns have the full
                               global pattern (columns reaching into the lower neighbour = "halo" vertices);
* local vertex numbering       sorted global ids of (halo | owned | ghost) -- monotone, so sorted local
                               patterns are sorted global patterns.

Both sides derive the exchange lists from the same element sets (the sender's own elements touching the
receiver's vertices ARE the receiver's pattern-only elements), sorted by global id, so no index metadata has
to be communicated.

Everything in :class:`ShardLayout` is host-side numpy and is unit-tested on CPU (``tests/test_sharding.py``,
including a world_size-2 gloo run); :class:`Shard` adds the device plan and the NCCL exchange.
"""

import os

import numpy as np


def element_cuts(t, world, align=1):
    """Contiguous element blocks; cuts rounded to multiples of ``align`` (grid slabs: elements per plane of cells)."""
    cuts = [int(round(p * (t / align) / world)) * align for p in range(world + 1)]
    cuts[0], cuts[-1] = 0, t
    return np.asarray(cuts, dtype=np.int64)


def vertex_cuts(T, ecuts, n):
    """v_p = smallest vertex referenced by block p (v_0 = 0, v_world = n); must be non-decreasing."""
    world = len(ecuts) - 1
    v = np.zeros(world + 1, dtype=np.int64)
    v[world] = n
    for p in range(1, world):
        blk = T[ecuts[p]:ecuts[p + 1]]
        v[p] = blk.min() if blk.size else v[p - 1]
    for p in range(world - 1, 0, -1):           # empty blocks / non-monotone meshes: keep ranges ordered
        v[p] = min(v[p], v[p + 1])
    assert not np.any(np.diff(v) < 0)
    return v


class ShardLayout:
    """Host-side description of one rank of a sharded mesh (numpy only)."""

    def __init__(self, rank, world, dim, T_own, T_pattern, pattern_rank, vcuts):
        """``T_own``: (t_own, K) global ids of this rank's elements; ``T_pattern``: (t_pat, K) global ids of
        lower ranks' elements touching vertices owned here, ``pattern_rank`` their owning ranks; ``vcuts`` the
        global vertex ownership cuts."""
        self.rank, self.world, self.dim = rank, world, dim
        K = dim + 1
        T_own = np.asarray(T_own, dtype=np.int64).reshape(-1, K)
        T_pattern = np.asarray(T_pattern, dtype=np.int64).reshape(-1, K)
        self.vcuts = np.asarray(vcuts, dtype=np.int64)
        self.v_lo, self.v_hi = int(self.vcuts[rank]), int(self.vcuts[rank + 1])
        self.t_own, self.t_pattern = T_own.shape[0], T_pattern.shape[0]
        if T_own.size and T_own.min() < self.v_lo:
            raise ValueError("an element references a vertex owned by a lower rank")
        owned = np.arange(self.v_lo, self.v_hi, dtype=np.int64)
        self.l2g = np.unique(np.concatenate([owned, T_own.ravel(), T_pattern.ravel()]))
        self.n_local = self.l2g.size
        self.own_lo = int(np.searchsorted(self.l2g, self.v_lo))      # local ids [own_lo, own_hi) are owned
        self.own_hi = int(np.searchsorted(self.l2g, self.v_hi))
        self.T_local = np.searchsorted(self.l2g, np.concatenate([T_own, T_pattern], axis=0)).astype(np.int64)
        self.T_own_global, self.T_pattern_global = T_own, T_pattern
        self.pattern_rank = np.asarray(pattern_rank, dtype=np.int64).reshape(-1)

        # ---- sends: ghost rows grouped by owner (higher ranks), derived from own elements
        self.send = {}
        ghost = T_own[T_own >= self.v_hi] if T_own.size else np.zeros(0, np.int64)
        for q in np.unique(np.searchsorted(self.vcuts, ghost, side="right") - 1):
            q = int(q)
            mask = np.any((T_own >= self.vcuts[q]) & (T_own < self.vcuts[q + 1]), axis=1)
            self.send[q] = _interface_rows(T_own[mask], int(self.vcuts[q]), int(self.vcuts[q + 1]))
        # ---- receives: from each lower rank p, the same lists derived from its pattern-only elements
        self.recv = {}
        for p in np.unique(self.pattern_rank):
            p = int(p)
            self.recv[p] = _interface_rows(T_pattern[self.pattern_rank == p], self.v_lo, self.v_hi)

        # ---- vector halo lists: which owned vertices each neighbour needs / which local copies it refreshes
        #   to a LOWER rank p: its ghost vertices owned here  == rows of recv[p]
        #   to a HIGHER rank q: the vertices owned here among my elements that touch q's vertices (q's halo)
        self.vsend, self.vrecv = {}, {}
        for p, (rows, _, _) in self.recv.items():
            self.vsend[p] = rows
        for q in self.send:
            mask = np.any((T_own >= self.vcuts[q]) & (T_own < self.vcuts[q + 1]), axis=1)
            vs = np.unique(T_own[mask])
            self.vsend[q] = vs[(vs >= self.v_lo) & (vs < self.v_hi)]
        for q, (rows, _, _) in self.send.items():          # my ghosts owned by q
            self.vrecv[q] = rows
        for p in self.recv:                                # my halo vertices owned by p
            vs = np.unique(T_pattern[self.pattern_rank == p])
            self.vrecv[p] = vs[(vs >= self.vcuts[p]) & (vs < self.vcuts[p + 1])]

    # ------------------------------------------------------------------ index maps
    def vector_maps(self):
        """``(send, recv)``: rank -> int32 local DOF indices for refreshing the non-owned copies of a vector."""
        d = self.dim
        f = lambda vs: (self.to_local(vs)[:, None] * d + np.arange(d)[None, :]).ravel().astype(np.int32)  # noqa: E731
        return {r: f(v) for r, v in self.vsend.items()}, {r: f(v) for r, v in self.vrecv.items()}

    def to_local(self, g):
        loc = np.searchsorted(self.l2g, g)
        assert np.array_equal(self.l2g[loc], g)
        return loc

    def exchange_maps(self, bptr, bcol):
        """Scalar index lists into this rank's gradient (n_local*dim) and CSR values for every neighbour.

        Returns ``(send, recv)``: dicts ``rank -> (g_idx, h_idx)`` int32 arrays.  ``send[q]`` gathers the partial
        gradient entries / Hessian values of the ghost rows owned by ``q``; ``recv[p]`` are the positions in the
        OWNED rows where the values arriving from ``p`` are added.  Orders match by construction."""
        d = self.dim
        out_s, out_r = {}, {}
        for table, out in ((self.send, out_s), (self.recv, out_r)):
            for r, (rows, rptr, cols) in table.items():
                lr = self.to_local(rows)
                lc = self.to_local(cols)
                g_idx = (lr[:, None] * d + np.arange(d)[None, :]).ravel()
                h_idx = _row_value_positions(lr, rptr, lc, bptr, bcol, d)
                out[r] = (g_idx.astype(np.int32), h_idx.astype(np.int32))
        return out_s, out_r


def _interface_rows(Tsub, v_lo, v_hi):
    """For the elements ``Tsub`` (global ids): rows = sorted vertices in [v_lo, v_hi) they touch, and for each
    row its sorted column vertices (all vertices sharing an element with it).  Returns (rows, rptr, cols)."""
    K = Tsub.shape[1]
    a = np.repeat(Tsub, K, axis=1).ravel()                     # row vertex of every (a, b) pair
    b = np.tile(Tsub, (1, K)).ravel()
    keep = (a >= v_lo) & (a < v_hi)
    pairs = np.unique(np.stack([a[keep], b[keep]], axis=1), axis=0)
    if pairs.size == 0:
        return np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.int64)
    rows, first = np.unique(pairs[:, 0], return_index=True)
    rptr = np.concatenate([first, [pairs.shape[0]]]).astype(np.int64)
    return rows, rptr, pairs[:, 1].copy()


def _row_value_positions(lrows, rptr, lcols, bptr, bcol, d):
    """Positions, in the canonical scalar-CSR value array, of the blocks (row, col) listed per row, in the order
    row -> block-row i -> listed column -> k  (the order both sides of an exchange use)."""
    out = []
    for j, r in enumerate(lrows):
        b0, b1 = int(bptr[r]), int(bptr[r + 1])
        nb = b1 - b0
        cols = lcols[rptr[j]:rptr[j + 1]]
        s = np.searchsorted(bcol[b0:b1], cols)
        if np.any(s >= nb) or not np.array_equal(bcol[b0:b1][s], cols):
            raise ValueError("interface block missing from the local pattern")
        base = b0 * d * d + s * d                               # (ncols,)
        pos = base[None, :, None] + (np.arange(d) * nb * d)[:, None, None] + np.arange(d)[None, None, :]
        out.append(pos.reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


# ---------------------------------------------------------------------------------- builders
def layout_from_global(T, n, dim, rank, world, align=1):
    """Shard layout of ``rank`` from the full element list (tests and small meshes).

    The caller may list the elements in any order: they are first sorted (stably) by their smallest vertex id, which
    makes the contiguous blocks own contiguous, non-decreasing vertex ranges (SURVEY §8e "after ordering elements by
    minimum vertex id").  ``layout.own_elements`` are the caller's indices of this rank's elements, in the order the
    shard lists them (per-element inputs of the rank are ``mu[layout.own_elements]`` etc.)."""
    T = np.asarray(T, dtype=np.int64)
    order = np.argsort(T.min(axis=1), kind="stable")
    T = T[order]
    ecuts = element_cuts(T.shape[0], world, align)
    vcuts = vertex_cuts(T, ecuts, n)
    own = T[ecuts[rank]:ecuts[rank + 1]]
    lower = T[:ecuts[rank]]
    touch = np.any((lower >= vcuts[rank]) & (lower < vcuts[rank + 1]), axis=1)
    pat = lower[touch]
    pat_rank = np.searchsorted(ecuts, np.nonzero(touch)[0], side="right") - 1
    lay = ShardLayout(rank, world, dim, own, pat, pat_rank, vcuts)
    lay.own_elements = order[ecuts[rank]:ecuts[rank + 1]]
    return lay, ecuts


def layout_grid_slab(cells, rank, world):
    """Same layout for the synthetic Kuhn / 2-triangle grids WITHOUT materialising the global mesh: rank ``p``
    gets the cell planes ``[i_p, i_{p+1})`` along the slowest axis (elements are cell-major, so this is a
    contiguous element block) plus, as pattern-only elements, the lower neighbour's last plane of cells."""
    from . import synthetic as syn
    cells = tuple(cells)
    dim = len(cells)
    mx = cells[0]
    per_cell = 6 if dim == 3 else 2
    plane_cells = int(np.prod(cells[1:]))
    plane_verts = int(np.prod([c + 1 for c in cells[1:]]))
    icuts = [int(round(p * mx / world)) for p in range(world + 1)]
    i0, i1 = icuts[rank], icuts[rank + 1]
    n = (mx + 1) * plane_verts
    vcuts = np.asarray([icuts[p] * plane_verts for p in range(world)] + [n], dtype=np.int64)
    T_own = syn.grid_elements(cells, i0, i1)
    if rank > 0 and i0 > 0:
        T_pat = syn.grid_elements(cells, i0 - 1, i0)
        T_pat = T_pat[np.any(T_pat >= vcuts[rank], axis=1)]
        pat_rank = np.full(T_pat.shape[0], np.searchsorted(icuts, i0 - 1, side="right") - 1, dtype=np.int64)
    else:
        T_pat = np.zeros((0, dim + 1), dtype=np.int64)
        pat_rank = np.zeros(0, dtype=np.int64)
    lay = ShardLayout(rank, world, dim, T_own, T_pat, pat_rank, vcuts)
    lay.t_total = mx * plane_cells * per_cell
    lay.n_total = n
    return lay


class Shard:
    """One rank's device plan, its interface handling (recompute or NCCL exchange) and the distributed solve (needs a GPU
    and an initialised process group)."""

    def __init__(self, layout, X_local, device=0, tile_elems=0, interface=None):
        """``interface``: how the rows this rank owns get the contributions of the lower neighbour's interface elements
        (the one layer of elements that touches them):

        * ``"exchange"``  -- north_star's scheme: the neighbour evaluates them, packs the partial gradient rows and
          Hessian block-rows of its ghost vertices and sends them over NVLink (one grouped NCCL send/recv per
          assembly); they are added here in rank order;
        * ``"recompute"`` -- this rank evaluates that layer itself (ghost elements: +1 plane of cells, 5.8 % more
          elements at 8 ranks of the 139^3 grid) and the assembly needs no communication at all; the energy still
          sums every element once (``t_energy``).  Same owned rows to rounding (another summation order).

        Default: ``SKB_SHARD_INTERFACE`` or ``"recompute"`` (faster at every rank count measured)."""
        import torch
        from .plan import MeshPlan
        self.layout = layout
        self.X_local = np.ascontiguousarray(X_local, dtype=np.float64)
        self.coarse = None
        self.interface = interface or os.environ.get("SKB_SHARD_INTERFACE", "recompute")
        if self.interface not in ("exchange", "recompute"):
            raise ValueError("interface must be 'exchange' or 'recompute'")
        if self.interface == "recompute":
            self.plan = MeshPlan(X=X_local, T=layout.T_local, device=device, tile_elems=tile_elems,
                                 t_energy=layout.t_own)
        else:
            self.plan = MeshPlan(X=X_local, T=layout.T_local, device=device, tile_elems=tile_elems,
                                 t_active=layout.t_own)
        bptr, bcol = self.plan.block_pattern()
        send, recv = layout.exchange_maps(bptr, bcol)
        dev = torch.device("cuda", device)
        self.device = dev
        mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        self.send = {r: (mk(g), mk(h)) for r, (g, h) in send.items()}
        self.recv = {r: (mk(g), mk(h)) for r, (g, h) in recv.items()}
        f64 = torch.float64
        self.sbuf = {r: torch.empty(g.numel() + h.numel(), dtype=f64, device=dev) for r, (g, h) in self.send.items()}
        self.rbuf = {r: torch.empty(g.numel() + h.numel(), dtype=f64, device=dev) for r, (g, h) in self.recv.items()}
        self.exchange_launches = 2 * (len(self.send) + len(self.recv)) if self.interface == "exchange" else 0
        self.exchange_bytes = 8 * sum(b.numel() for b in self.sbuf.values()) if self.interface == "exchange" else 0
        d = layout.dim
        self.nnz_owned = int(bptr[layout.own_hi] - bptr[layout.own_lo]) * d * d

    def exchange(self, g_d, vals_d):
        """Adds the lower neighbours' interface contributions into the owned rows of ``g_d`` / ``vals_d`` (device
        tensors in this rank's local numbering).  Pack -> NCCL send/recv -> scatter-add, all on the current stream."""
        import torch
        import torch.distributed as dist
        from ._lib import check, load
        if self.interface == "recompute":      # the owned rows are complete: this rank evaluated every element they see
            return
        lib = load()
        st = torch.cuda.current_stream().cuda_stream
        ops = []
        for q, (gi, hi) in sorted(self.send.items()):
            buf = self.sbuf[q]
            check(lib.skb_gather_dev(g_d.data_ptr(), gi.data_ptr(), gi.numel(), buf.data_ptr(), st))
            check(lib.skb_gather_dev(vals_d.data_ptr(), hi.data_ptr(), hi.numel(), buf.data_ptr() + 8 * gi.numel(), st))
            ops.append(dist.P2POp(dist.isend, buf, q))
        for p in sorted(self.recv):
            ops.append(dist.P2POp(dist.irecv, self.rbuf[p], p))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for p, (gi, hi) in sorted(self.recv.items()):          # fixed rank order -> deterministic sums
            buf = self.rbuf[p]
            check(lib.skb_scatter_add_dev(g_d.data_ptr(), gi.data_ptr(), gi.numel(), buf.data_ptr(), st))
            check(lib.skb_scatter_add_dev(vals_d.data_ptr(), hi.data_ptr(), hi.numel(), buf.data_ptr() + 8 * gi.numel(), st))


    # ------------------------------------------------------------------ vectors
    def _vector_lists(self):
        if not hasattr(self, "_vsend"):
            import torch
            vs, vr = self.layout.vector_maps()
            mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
            self._vsend = {r: mk(i) for r, i in vs.items()}
            self._vrecv = {r: mk(i) for r, i in vr.items()}
            f64 = torch.float64
            self._vsbuf = {r: torch.empty(i.numel(), dtype=f64, device=self.device) for r, i in self._vsend.items()}
            self._vrbuf = {r: torch.empty(i.numel(), dtype=f64, device=self.device) for r, i in self._vrecv.items()}
        return self._vsend, self._vrecv

    def halo_exchange(self, v_d):
        """Refreshes the non-owned (halo / ghost) entries of the local vector ``v_d`` from their owners."""
        import torch
        import torch.distributed as dist
        from ._lib import check, load
        lib = load()
        vsend, vrecv = self._vector_lists()
        st = torch.cuda.current_stream().cuda_stream
        ops = []
        for r, idx in sorted(vsend.items()):
            check(lib.skb_gather_dev(v_d.data_ptr(), idx.data_ptr(), idx.numel(), self._vsbuf[r].data_ptr(), st))
            ops.append(dist.P2POp(dist.isend, self._vsbuf[r], r))
        for r in sorted(vrecv):
            ops.append(dist.P2POp(dist.irecv, self._vrbuf[r], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for r, idx in sorted(vrecv.items()):
            check(lib.skb_scatter_dev(v_d.data_ptr(), idx.data_ptr(), idx.numel(), self._vrbuf[r].data_ptr(), st))

    # ------------------------------------------------------------------ native NCCL driving (opt-in)
    def enable_native_nccl(self, graph=False):
        """Drives the distributed PCG from C++ with NCCL called directly (``csrc/capi_nccl.cu``) instead of from this
        module's Python loop: same kernels, same order -- the iterates agree to rounding (bit for bit at two ranks, where
        the all-reduce has one summation order) -- without Python and torch.distributed between the steps.  Collective: every rank calls it.  The
        library creates its own communicator from an id made on rank 0 and broadcast here.  Also switched on by
        ``SKB_NATIVE_NCCL=1`` in ``make_shard`` (``=2`` / ``graph=True``: CUDA-graph replay of the iterations)."""
        import torch
        import torch.distributed as dist
        from ._lib import check, load, ptr
        lib = load()
        lay = self.layout
        idbuf = np.zeros(128, dtype=np.uint8)
        if lay.rank == 0:
            check(lib.skb_nccl_unique_id(ptr(idbuf), idbuf.size))
        t = torch.from_numpy(idbuf).to(self.device)
        dist.broadcast(t, 0)
        idbuf = np.ascontiguousarray(t.cpu().numpy())
        check(lib.skb_nccl_init(self.plan._h, ptr(idbuf), idbuf.size, lay.rank, lay.world))
        vsend, vrecv = self._vector_lists()
        peers = sorted(set(vsend) | set(vrecv))
        i64 = lambda f: np.array([f(r) for r in peers], dtype=np.int64)  # noqa: E731
        arrs = (np.array(peers, dtype=np.int32),
                i64(lambda r: vsend[r].numel() if r in vsend else 0),
                i64(lambda r: vsend[r].data_ptr() if r in vsend else 0),
                i64(lambda r: self._vsbuf[r].data_ptr() if r in vsend else 0),
                i64(lambda r: vrecv[r].numel() if r in vrecv else 0),
                i64(lambda r: vrecv[r].data_ptr() if r in vrecv else 0),
                i64(lambda r: self._vrbuf[r].data_ptr() if r in vrecv else 0))
        check(lib.skb_nccl_set_halo(self.plan._h, len(peers), *[ptr(a) for a in arrs]))
        self._native = True
        self._native_graph = bool(graph)   # full chunks of `check_every` iterations replay one CUDA graph

    def enable_peer_transport(self):
        """Maps every rank's solver slab into every other rank (CUDA IPC over NVLink/NVSwitch) so that the per-iteration
        halo exchange and the reduction of ``skb_dist_pcg2`` become stores into the peers' memory issued by the
        producing kernels (``csrc/capi_pcg2.cu``, transport 1) instead of NCCL calls.  Collective.  Returns whether
        ALL ranks succeeded (decided together, so every rank takes the same branch afterwards)."""
        import torch
        import torch.distributed as dist
        from ._lib import load, ptr
        lib = load()
        if not getattr(self, "_native", False):
            self.enable_native_nccl()
        world = self.layout.world
        h = np.zeros(64, dtype=np.uint8)
        meta = np.zeros(1 + world, dtype=np.int64)
        ok = lib.skb_pcg2_peer_export(self.plan._h, ptr(h), ptr(meta), meta.size) == 0
        blob = np.concatenate([h, meta.view(np.uint8)])
        t = torch.from_numpy(blob).to(self.device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            blobs = [o.cpu().numpy() for o in out]
            handles = np.ascontiguousarray(np.concatenate([b[:64] for b in blobs]))
            metas = np.ascontiguousarray(np.concatenate([b[64:] for b in blobs]).view(np.int64))
            ok = lib.skb_pcg2_peer_import(self.plan._h, ptr(handles), ptr(metas)) == 0
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self._peer = int(flag.item()) == 1
        if not self._peer:
            self._peer_error = lib.skb_last_error().decode()
        return self._peer

    def _pcg2(self, vals_d, diag_d, rhs_d, x_d, rtol, max_iter, check_every=25, graph=True, transport=0):
        """Single-reduction PCG with the coarse restriction folded into its one all-reduce (``csrc/capi_pcg2.cu``):
        per iteration 7 kernels, one grouped NCCL send/recv and one NCCL all-reduce, replayed as a CUDA graph."""
        import ctypes
        import torch
        from ._lib import DistPcg2Args, check, load
        lib = load()
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        a = DistPcg2Args(P(vals_d), P(diag_d), P(rhs_d), P(x_d), torch.cuda.current_stream().cuda_stream, float(rtol),
                         int(self.layout.own_lo), int(self.layout.own_hi), int(max_iter), int(check_every),
                         1 if graph else 0, 1 if self.coarse is not None else 0, int(transport), 0)
        iters = ctypes.c_int32(0)
        relres = ctypes.c_double(0.0)
        check(lib.skb_dist_pcg2(self.plan._h, ctypes.byref(a), ctypes.byref(iters), ctypes.byref(relres)))
        tm = np.zeros(4)
        check(lib.skb_dist_pcg2_times(self.plan._h, tm.ctypes.data_as(ctypes.c_void_p)))
        self.last_solve_ms = dict(setup=float(tm[0]), coarse_inverse=float(tm[1]), iterations=float(tm[2]), total=float(tm[3]),
                                  per_iteration=float(tm[2]) / max(int(iters.value), 1))
        return int(iters.value), float(relres.value)

    def _pcg_native(self, vals_d, diag_d, rhs_d, x_d, rtol, max_iter, check_every):
        import ctypes
        import torch
        from ._lib import DistPcgArgs, check, load
        lib = load()
        w = self._work()
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        c = self.coarse
        a = DistPcgArgs(P(vals_d), P(diag_d), P(rhs_d), P(x_d), P(w["dinv"]), P(w["r"]), P(w["z"]), P(w["p"]), P(w["q"]),
                        P(w["s"]), P(w["work"]), None if c is None else P(c["Ac"]), None if c is None else P(c["rc"]),
                        None if c is None else P(c["zc"]), torch.cuda.current_stream().cuda_stream, float(rtol),
                        int(self.layout.own_lo), int(self.layout.own_hi), int(max_iter), int(check_every),
                        1 if getattr(self, "_native_graph", False) else 0, 0)
        iters = ctypes.c_int32(0)
        relres = ctypes.c_double(0.0)
        check(lib.skb_dist_pcg_native(self.plan._h, ctypes.byref(a), ctypes.byref(iters), ctypes.byref(relres)))
        return int(iters.value), float(relres.value)

    # ------------------------------------------------------------------ distributed PCG / Newton
    def _work(self):
        if not hasattr(self, "_w"):
            import torch
            f64 = torch.float64
            nd = self.plan.ndof
            z = lambda n: torch.zeros(n, dtype=f64, device=self.device)  # noqa: E731
            self._w = dict(r=z(nd), z=z(nd), p=z(nd), q=z(nd), dx=z(nd), rhs=z(nd), diag=z(nd), g=z(nd), xtrial=z(nd),
                           dinv=z(self.plan.n * self.layout.dim ** 2), vals=z(self.plan.nnz), s=z(8), work=z(3 * 2048),
                           ls=z(6))
        return self._w

    def set_coarse_space(self, n_agg_target=729):
        """Two-level PCG preconditioner on the sharded mesh (``csrc/coarse.cuh``): every rank bins ITS local vertices
        with the global bounding box, so aggregate ids agree across ranks; centres are global centroids of the owned
        vertices.  ``n_agg_target = 0`` removes it."""
        import torch
        import torch.distributed as dist
        from ._lib import check, load, ptr
        lib = load()
        lay, dim = self.layout, self.layout.dim
        if not n_agg_target:
            check(lib.skb_dist_coarse_set(self.plan._h, 0, None, None, lay.own_lo, lay.own_hi))
            self.coarse = None
            return 0
        X = self.X_local
        f64 = torch.float64
        lo = torch.from_numpy(X.min(axis=0)).to(self.device)
        hi = torch.from_numpy(X.max(axis=0)).to(self.device)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
        ext = np.maximum(hi - lo, 1e-300)
        edge = (np.prod(ext) / float(min(int(n_agg_target), 2048))) ** (1.0 / dim)
        nb = np.maximum(1, np.floor(ext / edge + 0.5).astype(np.int64))
        while int(np.prod(nb)) > 2048:
            nb[np.argmax(nb)] -= 1
        ib = np.minimum((np.floor((X - lo) / ext * nb)).astype(np.int64), nb - 1)
        flat = ib[:, 0]
        for a in range(1, dim):
            flat = flat * nb[a] + ib[:, a]
        nbox = int(np.prod(nb))
        own = slice(lay.own_lo, lay.own_hi)
        sums = np.zeros((nbox, dim + 1))
        sums[:, 0] = np.bincount(flat[own], minlength=nbox)
        for a in range(dim):
            sums[:, 1 + a] = np.bincount(flat[own], weights=X[own, a], minlength=nbox)
        sums_d = torch.from_numpy(sums).to(self.device)
        dist.all_reduce(sums_d)                                   # every vertex is owned by exactly one rank
        sums = sums_d.cpu().numpy()
        used = np.nonzero(sums[:, 0] > 0)[0]                      # drop empty boxes, identically on every rank
        remap = -np.ones(nbox, dtype=np.int64)
        remap[used] = np.arange(used.size)
        agg = remap[flat]
        assert (agg >= 0).all()
        cen = sums[used, 1:] / sums[used, :1]
        xrel = np.ascontiguousarray(X - cen[agg])
        agg32 = np.ascontiguousarray(agg.astype(np.int32))
        n_agg = int(used.size)
        check(lib.skb_dist_coarse_set(self.plan._h, n_agg, ptr(agg32), ptr(xrel), lay.own_lo, lay.own_hi))
        nc = (6 if dim == 3 else 3) * n_agg
        z = lambda n: torch.zeros(n, dtype=f64, device=self.device)  # noqa: E731
        self.coarse = dict(n_agg=n_agg, nc=nc, Ac=z(nc * nc), rc=z(nc), zc=z(nc))
        return n_agg

    def _coarse_correct(self, r_d, z_d, p_d, s, slot, work, st):
        """z += P Ainv P^T r (restriction all-reduced); s[slot] = r.z over the owned dofs."""
        import torch.distributed as dist
        from ._lib import check, load
        lib = load()
        c = self.coarse
        v0, v1 = self.layout.own_lo, self.layout.own_hi
        check(lib.skb_dist_coarse_restrict_dev(self.plan._h, r_d.data_ptr(), c["rc"].data_ptr(), st))
        dist.all_reduce(c["rc"])
        check(lib.skb_dist_coarse_correct_dev(self.plan._h, v0, v1, c["Ac"].data_ptr(), c["rc"].data_ptr(), c["zc"].data_ptr(),
                                              r_d.data_ptr(), z_d.data_ptr(), 0 if p_d is None else p_d.data_ptr(),
                                              s.data_ptr(), slot, work.data_ptr(), st))

    def pcg(self, vals_d, diag_d, rhs_d, x_d, rtol=1e-10, max_iter=20000, check_every=10):
        """Block-Jacobi PCG on the distributed matrix (owned rows per rank, complete after the interface exchange).
        Per iteration: halo exchange of p, SpMV + p.q, all-reduce, fused update, all-reduce, direction."""
        mode = getattr(self, "solver", None) or os.environ.get("SKB_DIST_PCG", "peer")
        if mode in ("peer", "peer_eager"):
            # default: the C++-driven single-reduction solve whose exchanges are NVLink stores into the peers' memory;
            # NCCL transport if the IPC mappings cannot be set up.  Collective: every rank takes the same branch.
            if not hasattr(self, "_peer"):
                self.enable_peer_transport()
            if self._peer:
                return self._pcg2(vals_d, diag_d, rhs_d, x_d, rtol, max_iter, graph=(mode == "peer"), transport=1)
            mode = "pcg2" if mode == "peer" else "pcg2_eager"
        if mode in ("pcg2", "pcg2_eager"):
            # the same solve over NCCL (grouped send/recv + one all-reduce per iteration)
            if not getattr(self, "_native", False):
                self.enable_native_nccl()
            return self._pcg2(vals_d, diag_d, rhs_d, x_d, rtol, max_iter, graph=(mode == "pcg2"))
        if mode in ("native", "native_graph"):            # the textbook loop below, issued from C++ (capi_nccl.cu)
            if not getattr(self, "_native", False):
                self.enable_native_nccl()
            self._native_graph = (mode == "native_graph")
            return self._pcg_native(vals_d, diag_d, rhs_d, x_d, rtol, max_iter, check_every)
        if mode != "python":
            raise ValueError("unknown distributed solver %r" % (mode,))
        import torch
        import torch.distributed as dist
        from ._lib import check, load
        lib = load()
        w = self._work()
        h = self.plan._h
        v0, v1 = self.layout.own_lo, self.layout.own_hi
        st = torch.cuda.current_stream().cuda_stream
        s = w["s"]
        P = lambda t: t.data_ptr()  # noqa: E731
        dp = 0 if diag_d is None else P(diag_d)
        coarse = self.coarse
        if coarse is not None:
            # coarse matrix of this system: owned fine blocks per rank, summed over the ranks, inverted by every rank
            check(lib.skb_dist_coarse_assemble_dev(h, P(vals_d), dp, P(coarse["Ac"]), st))
            dist.all_reduce(coarse["Ac"])
            if lib.skb_dist_coarse_invert_dev(h, P(coarse["Ac"]), st) != 0:
                coarse = None        # degenerate aggregate (same matrix on every rank): block-Jacobi for this solve
        check(lib.skb_dist_pcg_init_dev(h, P(vals_d), dp, v0, v1, P(rhs_d), P(w["dinv"]), P(x_d), P(w["r"]), P(w["z"]),
                                        P(w["p"]), P(s), P(w["work"]), st))
        if coarse is not None:
            self._coarse_correct(w["r"], w["z"], w["p"], s, 0, w["work"], st)
        dist.all_reduce(s[0:2])
        bb = float(s[1].item())
        it = 0
        rr = bb
        if not bb > 0.0:
            return 0, 0.0
        while it < max_iter:
            for _ in range(min(check_every, max_iter - it)):
                self.halo_exchange(w["p"])
                check(lib.skb_dist_spmv_dot_dev(h, P(vals_d), dp, v0, v1, P(w["p"]), P(w["q"]), P(s), P(w["work"]), st))
                dist.all_reduce(s[2:3])
                check(lib.skb_dist_pcg_update_dev(h, v0, v1, P(w["dinv"]), P(w["p"]), P(w["q"]), P(x_d), P(w["r"]),
                                                  P(w["z"]), P(s), P(w["work"]), st))
                if coarse is not None:
                    self._coarse_correct(w["r"], w["z"], None, s, 3, w["work"], st)
                dist.all_reduce(s[3:5])
                check(lib.skb_dist_pcg_direction_dev(h, v0, v1, P(w["z"]), P(w["p"]), P(s), st))
                it += 1
            rr = float(s[1].item())
            if not rr > rtol * rtol * bb:
                break
        return it, float(np.sqrt(rr / bb))

    def newton_step(self, material, x_d, x_tilde_d=None, mass_d=None, kin_scale=0.0, fext_d=None, psd_mode=1,
                    max_iter=1, do_line_search=True, tolerance=1e-6, ls_alpha=0.01, ls_beta=0.5, ls_max_iter=100,
                    ls_threshold=1e-12, pcg_rtol=1e-10, pcg_max_iter=20000, pin_k_d=None, pin_target_d=None,
                    contact_plane=None, contact_sphere=None):
        """One implicit step on the sharded mesh (same loop as ``skb_newton`` / solvers/newton.py:42-70): assembly +
        interface exchange, distributed PCG, Armijo backtracking on the all-reduced total energy.  ``x_d`` (local
        numbering, all local vertices) is updated in place; materials must have been set.

        The per-vertex terms every reference caller adds before the solve act on the OWNED vertices of
        each rank: ``pin_k_d`` / ``pin_target_d`` -- device vectors (local dofs) of a Dirichlet penalty
        ``1/2 sum k_i (x_i - t_i)^2`` (``dirichlet_penalty.py:65-142`` with a diagonal ``SGamma``); ``contact_plane`` =
        ``dict(k=, p=, n=[, w_d=])`` and ``contact_sphere`` = ``dict(k=, p=, r=[, w_d=])`` -- penalty springs against a
        plane / sphere (``energies/contact_springs_plane.py``, ``contact_springs_sphere.py``), ``w_d`` the per-vertex
        weights (device, local vertices; default 1).  Their energies join the all-reduce of the line search."""
        import torch
        import torch.distributed as dist
        from ._lib import MATERIAL_IDS, check, load
        lib = load()
        w = self._work()
        h = self.plan._h
        v0, v1 = self.layout.own_lo, self.layout.own_hi
        st = torch.cuda.current_stream().cuda_stream
        P = lambda t: 0 if t is None else t.data_ptr()  # noqa: E731
        mat = MATERIAL_IDS[material]
        info = dict(iters=-1, alphas=[], pcg_iters=0, pcg_relres=0.0, step_norm=0.0)
        ls = w["ls"]
        dim = self.layout.dim
        contacts = []
        for kind, c in ((0, contact_plane), (1, contact_sphere)):
            if c is not None and c.get("k"):
                from ._lib import f64, ptr
                pp = f64(np.asarray(c["p"], dtype=np.float64).reshape(-1))
                nn = f64(np.asarray(c.get("n", np.zeros(dim)), dtype=np.float64).reshape(-1))
                if pp.size != dim or nn.size != dim:
                    raise ValueError("contact: p and n must have dim entries")
                contacts.append((kind, float(c["k"]), pp, nn, float(c.get("r", 0.0)), c.get("w_d")))

        def contact(xv_d, g_d, vals_d, e_slot):
            from ._lib import ptr
            for ci, (kind, k, pp, nn, r, w_d) in enumerate(contacts):
                check(lib.skb_dist_contact_dev(h, v0, v1, P(xv_d), kind, k, ptr(pp), ptr(nn), r, P(w_d), P(g_d), P(vals_d),
                                               0 if e_slot is None else P(ls) + 8 * (e_slot + ci), P(w["work"]), st))

        def total_energy(sstep, with_g):
            ls.zero_()
            check(lib.skb_dist_newton_terms_dev(h, v0, v1, P(x_d), P(w["dx"]), float(sstep), P(fext_d), P(mass_d),
                                                P(x_tilde_d), float(kin_scale), P(pin_k_d), P(pin_target_d),
                                                P(w["g"]) if with_g else 0, P(w["xtrial"]), P(ls), P(w["work"]), st))
            self.halo_exchange(w["xtrial"])
            check(lib.skb_energy_dev(h, mat, P(w["xtrial"]), 0, P(ls) + 24, st))
            contact(w["xtrial"], None, None, 4)
            dist.all_reduce(ls)
            e = ls.cpu().numpy()
            return float(e[0] + e[3] + e[4] + e[5]), float(e[1]), float(e[2])

        self.halo_exchange(x_d)
        for it in range(max_iter):
            self.gradient_hessian_dev(material, psd_mode, x_d, w["g"], w["vals"])
            contact(x_d, w["g"], w["vals"], None)       # gradient rows and diagonal blocks of the owned vertices
            check(lib.skb_dist_newton_rhs_dev(h, v0, v1, P(x_d), P(fext_d), P(mass_d), P(x_tilde_d), float(kin_scale),
                                              P(pin_k_d), P(pin_target_d), P(w["g"]), P(w["rhs"]), P(w["diag"]), st))
            pit, relres = self.pcg(w["vals"], w["diag"], w["rhs"], w["dx"], rtol=pcg_rtol, max_iter=pcg_max_iter)
            if not np.isfinite(relres):
                from ._lib import SimkitB200Error
                raise SimkitB200Error("distributed Newton step: the linear solve produced a non-finite residual; the system "
                                      "must be symmetric positive definite")
            alpha = 1.0
            if do_line_search:
                e0, gdx, dx2 = total_energy(0.0, True)
                t, ok = 1.0, False
                for _ in range(ls_max_iter):
                    e1, _, _ = total_energy(t, False)
                    if e1 <= e0 + ls_alpha * t * gdx + ls_threshold:
                        ok = True
                        break
                    t *= ls_beta
                alpha = t if ok else 0.0
            else:
                _, _, dx2 = total_energy(1.0, False)
            if alpha > 0.0:
                x_d.copy_(w["xtrial"])          # holds x + alpha*dx with refreshed halo / ghost copies
            step = alpha * float(np.sqrt(dx2))
            info["iters"] = it
            info["alphas"].append(alpha)
            info["pcg_iters"] += pit
            info["pcg_relres"] = relres
            info["step_norm"] = step
            if step < tolerance:
                break
        return info

    # ------------------------------------------------------------------ reduced (subspace) tier
    def reduced(self, material, B_local, z, x0_local=None, psd_mode=1, want=("E", "g", "H")):
        """Reduced energy / gradient / Hessian ``B^T g``, ``B^T H B`` of the GLOBAL mesh (``elastic_*_z``,
        energies/elastic.py:749-782): every rank contracts its own elements against the rows of the basis that belong to
        its local vertices (``B_local``: ``(n_local*dim, r)``; ``None``: the basis set with ``plan.set_basis``), then ONE
        all-reduce sums the ``1 + r + r*r`` numbers over the ranks (320 KB at r = 200).  Collective."""
        import torch
        import torch.distributed as dist
        E, g, H = self.plan.reduced(material, B_local, z, x0=x0_local, psd_mode=psd_mode, want=want)
        r = int(np.asarray(z).size)
        buf = np.zeros(1 + r + r * r)
        buf[0] = E
        if g is not None:
            buf[1:1 + r] = g.ravel()
        if H is not None:
            buf[1 + r:] = H.ravel()
        t = torch.from_numpy(buf).to(self.device)
        dist.all_reduce(t)
        buf = t.cpu().numpy()
        self.reduced_allreduce_bytes = buf.nbytes
        return (float(buf[0]), None if g is None else buf[1:1 + r].reshape(r, 1).copy(),
                None if H is None else buf[1 + r:].reshape(r, r).copy())

    def lumped_mass_dofs(self, rho=1.0):
        """Device vector (local dofs) of the lumped masses ``massmatrix.py:41-49`` of the GLOBAL mesh on this rank's
        owned dofs: local element sums, then the same interface exchange as the gradient."""
        import torch
        w = self._work()
        d = self.layout.dim
        m = np.repeat(self.plan.vertex_masses(rho), d)
        m_d = torch.from_numpy(m).to(self.device)
        w["vals"].zero_()
        self.exchange(m_d, w["vals"])
        return m_d

    # ------------------------------------------------------------------ public entry points
    def set_materials(self, mu, lam, vol=None):
        self.plan.set_materials(mu, lam, self.plan.volume() if vol is None else vol)

    def gradient_hessian_dev(self, material, psd_mode, x_d, g_d, vals_d):
        """Device-resident sharded assembly: local fused kernel chain, then the interface exchange.  On return
        (stream-ordered) the OWNED rows ``[own_lo*dim, own_hi*dim)`` of ``g_d`` / ``vals_d`` are globally complete."""
        import torch
        from ._lib import MATERIAL_IDS, check, load
        st = torch.cuda.current_stream().cuda_stream
        check(load().skb_gradient_hessian_dev(self.plan._h, MATERIAL_IDS[material], int(psd_mode), x_d.data_ptr(), None,
                                              g_d.data_ptr(), vals_d.data_ptr(), st))
        self.exchange(g_d, vals_d)

    def gradient_hessian(self, material, x_host, mu, lam, vol=None, psd_mode=1, g_out=None, vals_out=None):
        """Host-pointer version (the call a user makes per rank): uploads this rank's state, assembles, exchanges and
        returns the owned gradient rows and the owned CSR value rows (local numbering, see ``owned_csr``)."""
        import torch
        f64 = torch.float64
        if not hasattr(self, "_x_d"):
            self._x_d = torch.empty(self.plan.ndof, dtype=f64, device=self.device)
            self._g_d = torch.empty(self.plan.ndof, dtype=f64, device=self.device)
            self._vals_d = torch.empty(self.plan.nnz, dtype=f64, device=self.device)
        self.set_materials(mu, lam, vol)
        xh = torch.from_numpy(np.ascontiguousarray(np.asarray(x_host, dtype=np.float64).reshape(-1)))
        self._x_d.copy_(xh, non_blocking=True)
        self.gradient_hessian_dev(material, psd_mode, self._x_d, self._g_d, self._vals_d)
        d = self.layout.dim
        g0, g1 = self.layout.own_lo * d, self.layout.own_hi * d
        v0, v1 = self.owned_value_range()
        g_out = np.empty(g1 - g0) if g_out is None else g_out
        vals_out = np.empty(v1 - v0) if vals_out is None else vals_out
        torch.from_numpy(g_out).copy_(self._g_d[g0:g1], non_blocking=True)
        torch.from_numpy(vals_out).copy_(self._vals_d[v0:v1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return g_out, vals_out

    def owned_value_range(self):
        """[v0, v1): the contiguous slice of the local CSR value array that holds the owned rows."""
        if not hasattr(self, "_vrange"):
            bptr, _ = self.plan.block_pattern()
            d = self.layout.dim
            self._vrange = (int(bptr[self.layout.own_lo]) * d * d, int(bptr[self.layout.own_hi]) * d * d)
        return self._vrange

    def owned_csr(self, vals_owned):
        """scipy CSR of the owned rows in GLOBAL numbering, shape (n_owned*dim, n_total*dim)."""
        import scipy.sparse as sps
        d = self.layout.dim
        indptr, indices = self.plan.csr_pattern()
        r0, r1 = self.layout.own_lo * d, self.layout.own_hi * d
        ip = indptr[r0:r1 + 1].astype(np.int64)
        gdof = (self.layout.l2g[:, None] * d + np.arange(d)[None, :]).ravel()
        cols = gdof[indices[ip[0]:ip[-1]]]
        n_total = int(self.layout.vcuts[-1])
        return sps.csr_matrix((vals_owned, cols, ip - ip[0]), shape=(r1 - r0, n_total * d))


def make_shard(workload, rank, world, device=0, tile_elems=0, sigma=0.1, interface=None):
    """Shard of a named synthetic config with its jittered state (bench.py, N > 1)."""
    from . import synthetic as syn
    cfg = syn.CONFIGS[workload]
    lay = layout_grid_slab(cfg["cells"], rank, world)
    X_local = syn.grid_vertices(cfg["cells"], cfg["extent"], lay.l2g)
    sh = Shard(lay, X_local, device=device, tile_elems=tile_elems, interface=interface)
    sh.U_local = syn.jittered_state_rows(cfg["cells"], cfg["extent"], lay.l2g, sigma=sigma)
    sh.t_total, sh.n_total = lay.t_total, lay.n_total
    sh.nnz_total = None
    if os.environ.get("SKB_NATIVE_NCCL", "") in ("1", "2"):      # the textbook loop driven from C++ (2: graph replay)
        sh.solver = "native_graph" if os.environ["SKB_NATIVE_NCCL"] == "2" else "native"
    return sh
