"""Element sharding: host-side layout logic on CPU.

* the slab generator agrees with the general partitioner;
* per-rank assembly (CPU replay of the kernel phase functions with pattern-only elements) followed by the
  interface exchange reproduces the owned rows of the global oracle matrix and gradient;
* the same exchange through torch.distributed (gloo, world_size 2) -- the N > 1 data path without a GPU.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe
from simkit_b200 import sharding as sh
from simkit_b200 import synthetic as syn
import hostsim

MAT = "stable_neo_hookean"
MAT_ID = 0


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("cells", [(5, 3, 4), (7, 5)])
@pytest.mark.parametrize("world", [2, 3])
def test_slab_layout_matches_general_partition(cells, world):
    X, T = syn.make_mesh(cells)
    per_plane = int(np.prod(cells[1:])) * (6 if len(cells) == 3 else 2)
    owned = np.zeros(X.shape[0], dtype=int)
    for r in range(world):
        a, _ = sh.layout_from_global(T, X.shape[0], len(cells), r, world, align=per_plane)
        b = sh.layout_grid_slab(cells, r, world)
        assert np.array_equal(a.l2g, b.l2g) and np.array_equal(a.T_local, b.T_local)
        assert a.t_own == b.t_own and a.t_pattern == b.t_pattern
        for q in a.send:
            for u, v in zip(a.send[q], b.send[q]):
                assert np.array_equal(u, v)
        owned[a.v_lo:a.v_hi] += 1
        assert b.n_total == X.shape[0] and b.t_total == T.shape[0]
    assert np.all(owned == 1)                       # every vertex row has exactly one owner


def _rank_assembly(lay, X, U, mu, lam, reorder=False):
    """CPU replay of one rank: plan with pattern-only elements, assembly of the own elements only."""
    Xl, Ul = X[lay.l2g], U[lay.l2g]
    out = hostsim.run(Xl, lay.T_local, MAT_ID, 1, Ul, mu, lam, None, tile_elems=32, t_active=lay.t_own, reorder=reorder)
    return out


def _apply_exchange(lays, outs):
    """numpy stand-in for pack -> send/recv -> scatter-add, in rank order."""
    maps = [lay.exchange_maps(o["bptr"], o["bcol"]) for lay, o in zip(lays, outs)]
    for q, lay in enumerate(lays):
        for p in sorted(lay.recv):
            gs, hs = maps[p][0][q]
            gr, hr = maps[q][1][p]
            assert gs.size == gr.size and hs.size == hr.size
            outs[q]["g"][gr] += outs[p]["g"][gs]
            outs[q]["vals"][hr] += outs[p]["vals"][hs]


@pytest.mark.parametrize("cells,world", [((5, 3, 4), 2), ((6, 3, 3), 3), ((9, 6), 2), ((9, 6), 3)])
def test_sharded_assembly_equals_global_oracle(cells, world):
    X, T = syn.make_mesh(cells)
    dim = len(cells)
    n = X.shape[0]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.lame()
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    lays = [sh.layout_grid_slab(cells, r, world) for r in range(world)]
    outs = [_rank_assembly(lay, X, U, mu, lam) for lay in lays]
    _apply_exchange(lays, outs)
    for lay, o in zip(lays, outs):
        nl = lay.n_local
        Ql = hostsim.csr_from_blocks(o["bptr"], o["bcol"], o["vals"], nl, dim).tocsr()
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        # owned rows of the local matrix, columns mapped back to global numbering
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        ref_rows = Q_ref[gdof[own]]
        assert rel(rows_g.toarray(), ref_rows.toarray()) < 1e-10
        assert rel(o["g"][own], g_ref[gdof[own]]) < 1e-10


@pytest.mark.parametrize("cells,world", [((5, 3, 4), 2), ((9, 6), 3)])
def test_sharding_a_shuffled_element_list(cells, world):
    """Elements listed in random order: the partitioner orders them by smallest vertex id itself instead
    of refusing, per-element materials follow through ``own_elements``, every rank's plan applies its own spatial
    element order on top, and the exchanged result is still the global oracle's."""
    X, T = syn.make_mesh(cells)
    dim, n = len(cells), X.shape[0]
    rng = np.random.default_rng(3)
    T = T[rng.permutation(T.shape[0])]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    lays = [sh.layout_from_global(T, n, dim, r, world)[0] for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate([l.own_elements for l in lays])), np.arange(T.shape[0]))
    outs = [_rank_assembly(l, X, U, mu[l.own_elements], lam[l.own_elements], reorder=True) for l in lays]
    _apply_exchange(lays, outs)
    owned = np.zeros(n, dtype=int)
    for lay, o in zip(lays, outs):
        Ql = hostsim.csr_from_blocks(o["bptr"], o["bcol"], o["vals"], lay.n_local, dim).tocsr()
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        assert rel(rows_g.toarray(), Q_ref[gdof[own]].toarray()) < 1e-10
        assert rel(o["g"][own], g_ref[gdof[own]]) < 1e-10
        owned[lay.v_lo:lay.v_hi] += 1
    assert np.all(owned == 1)


@pytest.mark.parametrize("cells,world", [((5, 3, 4), 2), ((6, 3, 3), 3), ((9, 6), 3)])
def test_recomputed_interface_needs_no_exchange(cells, world):
    """``Shard(interface="recompute")``: a rank that also evaluates its lower neighbour's interface elements (the
    pattern-only elements made active) has complete owned rows without any exchange."""
    X, T = syn.make_mesh(cells)
    dim, n = len(cells), X.shape[0]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.lame()
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    for r in range(world):
        lay = sh.layout_grid_slab(cells, r, world)
        o = hostsim.run(X[lay.l2g], lay.T_local, MAT_ID, 1, U[lay.l2g], mu, lam, None, tile_elems=32, reorder=True)
        Ql = hostsim.csr_from_blocks(o["bptr"], o["bcol"], o["vals"], lay.n_local, dim).tocsr()
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        assert rel(rows_g.toarray(), Q_ref[gdof[own]].toarray()) < 1e-10
        assert rel(o["g"][own], g_ref[gdof[own]]) < 1e-10


# ------------------------------------------------------------------------------ gloo, world size 2
def _gloo_worker(rank, world, port, cells, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, T = syn.make_mesh(cells)
    dim = len(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.lame()
    lay = sh.layout_grid_slab(cells, rank, world)
    o = _rank_assembly(lay, X, U, mu, lam)
    send, recv = lay.exchange_maps(o["bptr"], o["bcol"])
    ops, sbuf, rbuf = [], {}, {}
    for q, (gi, hi) in sorted(send.items()):
        sbuf[q] = torch.from_numpy(np.concatenate([o["g"][gi], o["vals"][hi]]))
        ops.append(dist.P2POp(dist.isend, sbuf[q], q))
    for p, (gi, hi) in sorted(recv.items()):
        rbuf[p] = torch.empty(gi.size + hi.size, dtype=torch.float64)
        ops.append(dist.P2POp(dist.irecv, rbuf[p], p))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for p, (gi, hi) in sorted(recv.items()):
        b = rbuf[p].numpy()
        o["g"][gi] += b[:gi.size]
        o["vals"][hi] += b[gi.size:]
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), g=o["g"], vals=o["vals"], bptr=o["bptr"], bcol=o["bcol"])
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_over_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    cells, world = (5, 3, 4), 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, cells, str(tmp_path)), nprocs=world, join=True)
    X, T = syn.make_mesh(cells)
    dim, n = 3, X.shape[0]
    U = syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.3)
    mu, lam = syn.lame()
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    for r in range(world):
        lay = sh.layout_grid_slab(cells, r, world)
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        Ql = hostsim.csr_from_blocks(d["bptr"], d["bcol"], d["vals"], lay.n_local, dim).tocsr()
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        assert rel(rows_g.toarray(), Q_ref[gdof[own]].toarray()) < 1e-10
        assert rel(d["g"][own], g_ref[gdof[own]]) < 1e-10
