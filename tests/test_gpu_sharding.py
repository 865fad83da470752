"""Sharded assembly on the GPU: per-rank plans with pattern-only elements (skb_plan_create_sharded), the device
pack / scatter-add kernels of the interface exchange, checked against the global oracle.  The ranks run one after
the other on cuda:0 (the transport itself -- NCCL send/recv -- is exercised by bench.py --gpus N and, on CPU,
by the gloo test in test_sharding.py)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe
from simkit_b200 import sharding as sh
from simkit_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
MAT = "stable_neo_hookean"


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("cells,world", [((6, 4, 5), 2), ((9, 4, 4), 4), ((12, 7), 3)])
def test_sharded_device_assembly(cells, world):
    import torch
    import simkit_b200 as sk
    from simkit_b200._lib import MATERIAL_IDS, PSD_AFTER_VOL, check, load
    lib = load()
    dev = torch.device("cuda", 0)
    X, T = syn.make_mesh(cells)
    dim, n = len(cells), X.shape[0]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    per_plane = int(np.prod(cells[1:])) * (6 if dim == 3 else 2)
    ecuts = sh.element_cuts(T.shape[0], world, per_plane)
    st = torch.cuda.current_stream().cuda_stream
    ranks = []
    for r in range(world):
        lay = sh.layout_grid_slab(cells, r, world)
        plan = sk.MeshPlan(X=X[lay.l2g], T=lay.T_local, tile_elems=32, t_active=lay.t_own)
        assert plan.t == lay.t_own
        sl = slice(ecuts[r], ecuts[r + 1])
        plan.set_materials(mu[sl], lam[sl], plan.volume())
        x_d = torch.from_numpy(U[lay.l2g].reshape(-1).copy()).to(dev)
        g_d = torch.empty(plan.ndof, dtype=torch.float64, device=dev)
        v_d = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
        check(lib.skb_gradient_hessian_dev(plan._h, MATERIAL_IDS[MAT], PSD_AFTER_VOL, x_d.data_ptr(), None,
                                           g_d.data_ptr(), v_d.data_ptr(), st))
        bptr, bcol = plan.block_pattern()
        send, recv = lay.exchange_maps(bptr, bcol)
        ranks.append(dict(lay=lay, plan=plan, g=g_d, v=v_d, send=send, recv=recv))
    torch.cuda.synchronize()
    # interface exchange through the device pack / scatter-add kernels, in rank order
    for q, R in enumerate(ranks):
        for p in sorted(R["recv"]):
            gs, hs = ranks[p]["send"][q]
            gr, hr = R["recv"][p]
            for src, dst, si, di in ((ranks[p]["g"], R["g"], gs, gr), (ranks[p]["v"], R["v"], hs, hr)):
                si_d, di_d = torch.from_numpy(si).to(dev), torch.from_numpy(di).to(dev)
                buf = torch.empty(si.size, dtype=torch.float64, device=dev)
                check(lib.skb_gather_dev(src.data_ptr(), si_d.data_ptr(), si.size, buf.data_ptr(), st))
                check(lib.skb_scatter_add_dev(dst.data_ptr(), di_d.data_ptr(), di.size, buf.data_ptr(), st))
    torch.cuda.synchronize()
    for R in ranks:
        lay, plan = R["lay"], R["plan"]
        indptr, indices = plan.csr_pattern()
        Ql = sps.csr_matrix((R["v"].cpu().numpy(), indices, indptr), shape=(plan.ndof, plan.ndof))
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        assert rel(rows_g.toarray(), Q_ref[gdof[own]].toarray()) < 1e-10
        assert rel(R["g"].cpu().numpy()[own], g_ref[gdof[own]]) < 1e-10
        # owned rows carry the full global pattern: structural nnz of those rows match the global structural pattern
        ip, ix, _, _ = oe.structural_pattern(T, n, dim)
        ref_nnz = ip[gdof[own] + 1] - ip[gdof[own]]
        assert np.array_equal(np.diff(indptr)[own], ref_nnz)
