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


@pytest.mark.parametrize("cells,world", [((6, 4, 5), 2), ((9, 4, 4), 4), ((12, 7), 3)])
def test_sharded_device_assembly_recompute(cells, world):
    """``Shard(interface="recompute")`` on the device: every rank evaluates its lower neighbour's interface elements
    itself (plan with ``t_energy`` = own elements), so its owned rows are complete without any exchange, and the
    per-rank energies -- own elements only -- add up to the global energy."""
    import ctypes
    import torch
    import simkit_b200 as sk
    from simkit_b200._lib import MATERIAL_IDS, PSD_AFTER_VOL, check, load
    lib = load()
    dev = torch.device("cuda", 0)
    X, T = syn.make_mesh(cells)
    dim, n = len(cells), X.shape[0]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.3)
    mu, lam = syn.lame()
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    g_ref = oe.gradient_x(MAT, U, J, mu, lam, vol).ravel()
    Q_ref = sps.csr_matrix(oe.hessian_x(MAT, U, J, mu, lam, vol, psd=True))
    E_ref = oe.energy_x(MAT, U, J, mu, lam, vol)
    st = torch.cuda.current_stream().cuda_stream
    E_sum = 0.0
    for r in range(world):
        lay = sh.layout_grid_slab(cells, r, world)
        plan = sk.MeshPlan(X=X[lay.l2g], T=lay.T_local, tile_elems=32, t_energy=lay.t_own)
        assert plan.t == lay.t_own + lay.t_pattern
        plan.set_materials(mu, lam, plan.volume())
        x_d = torch.from_numpy(U[lay.l2g].reshape(-1).copy()).to(dev)
        g_d = torch.empty(plan.ndof, dtype=torch.float64, device=dev)
        v_d = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
        e_d = torch.zeros(1, dtype=torch.float64, device=dev)
        check(lib.skb_gradient_hessian_dev(plan._h, MATERIAL_IDS[MAT], PSD_AFTER_VOL, x_d.data_ptr(), None,
                                           g_d.data_ptr(), v_d.data_ptr(), st))
        check(lib.skb_energy_dev(plan._h, MATERIAL_IDS[MAT], x_d.data_ptr(), None, e_d.data_ptr(), st))
        torch.cuda.synchronize()
        E_sum += float(e_d.item())
        indptr, indices = plan.csr_pattern()
        Ql = sps.csr_matrix((v_d.cpu().numpy(), indices, indptr), shape=(plan.ndof, plan.ndof))
        own = np.arange(lay.own_lo * dim, lay.own_hi * dim)
        gdof = (lay.l2g[:, None] * dim + np.arange(dim)[None, :]).ravel()
        rows = Ql[own]
        rows_g = sps.csr_matrix((rows.data, gdof[rows.indices], rows.indptr), shape=(own.size, n * dim))
        assert rel(rows_g.toarray(), Q_ref[gdof[own]].toarray()) < 1e-10
        assert rel(g_d.cpu().numpy()[own], g_ref[gdof[own]]) < 1e-10
    assert abs(E_sum - E_ref) <= 1e-12 * abs(E_ref)


# ------------------------------------------------------------------ 2 GPUs: NCCL exchange + distributed Newton
def _nccl_worker(rank, world, port, cells, tmp, interface="recompute"):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dim = len(cells)
    extent = tuple(1.0 for _ in cells)
    lay = sh.layout_grid_slab(cells, rank, world)
    shard = sh.Shard(lay, syn.grid_vertices(cells, extent, lay.l2g), device=rank, tile_elems=32, interface=interface)
    U = syn.jittered_state_rows(cells, extent, lay.l2g, sigma=0.2)
    mu, lam = syn.lame()
    shard.set_materials(mu, lam)
    rho, h = 1e3, 1e-2
    x_d = torch.from_numpy(U.reshape(-1).copy()).to(dev)
    mass_d = shard.lumped_mass_dofs(rho)
    fext_d = torch.zeros(lay.n_local, dim, dtype=torch.float64, device=dev)
    fext_d[:, 1] = -9.8
    fext_d = fext_d.reshape(-1) * mass_d
    xs = x_d.clone()
    info = shard.newton_step(MAT, xs, x_tilde_d=x_d, mass_d=mass_d, kin_scale=1.0 / h ** 2, fext_d=fext_d, max_iter=3,
                             pcg_rtol=1e-12)
    o0, o1 = lay.own_lo * dim, lay.own_hi * dim
    # same step with the two-level preconditioner (global aggregates, all-reduced coarse matrix and restriction)
    n_agg = shard.set_coarse_space(12)
    xs2 = x_d.clone()
    info2 = shard.newton_step(MAT, xs2, x_tilde_d=x_d, mass_d=mass_d, kin_scale=1.0 / h ** 2, fext_d=fext_d, max_iter=3,
                              pcg_rtol=1e-12)
    shard.set_coarse_space(0)
    # per-vertex terms on a sharded mesh: face x = 0 pinned with a penalty, plane and sphere contact
    Xl = syn.grid_vertices(cells, extent, lay.l2g)
    pin_k = np.zeros((lay.n_local, dim))
    pin_k[Xl[:, 0] == 0.0] = 1e6
    pin_k_d = torch.from_numpy(pin_k.reshape(-1)).to(dev)
    pin_t_d = torch.from_numpy(Xl.reshape(-1).copy()).to(dev)
    wv_d = mass_d.reshape(-1, dim)[:, 0].contiguous()
    xs3 = x_d.clone()
    info3 = shard.newton_step(MAT, xs3, x_tilde_d=x_d, mass_d=mass_d, kin_scale=1.0 / h ** 2, fext_d=fext_d, max_iter=3,
                              pcg_rtol=1e-12, pin_k_d=pin_k_d, pin_target_d=pin_t_d,
                              contact_plane=dict(k=1e5, p=[0.0, 0.2, 0.0], n=[0.0, 1.0, 0.0], w_d=wv_d),
                              contact_sphere=dict(k=1e5, p=[0.5, 0.5, 0.5], r=0.3, w_d=wv_d))
    np.savez(os.path.join(tmp, "newton%d.npz" % rank), x=xs.cpu().numpy()[o0:o1], lo=lay.v_lo, hi=lay.v_hi,
             alphas=np.asarray(info["alphas"]), mass=mass_d.cpu().numpy()[o0:o1], x2=xs2.cpu().numpy()[o0:o1],
             alphas2=np.asarray(info2["alphas"]), its=info["pcg_iters"], its2=info2["pcg_iters"], n_agg=n_agg,
             x3=xs3.cpu().numpy()[o0:o1], alphas3=np.asarray(info3["alphas"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("interface", ["recompute", "exchange"])
def test_distributed_newton_matches_single_gpu(tmp_path, interface):
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import simkit_b200 as sk
    cells, world = (8, 5, 5), 2
    port = 29600 + ((os.getpid() + (0 if interface == "recompute" else 31)) % 1000)
    mp.spawn(_nccl_worker, args=(world, port, cells, str(tmp_path), interface), nprocs=world, join=True)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.2)
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    plan = sk.MeshPlan(X=X, T=T, tile_elems=32)
    plan.set_materials(mu, lam, plan.volume())
    mass = np.repeat(plan.vertex_masses(rho), 3)
    fext = np.zeros_like(X)
    fext[:, 1] = -9.8
    fext = fext.reshape(-1) * mass
    x1, info = plan.newton(MAT, U.reshape(-1), x_tilde=U.reshape(-1), mass=mass, kin_scale=1.0 / h ** 2, f_ext=fext,
                           max_iter=3, pcg_rtol=1e-12)
    pin_k = np.zeros_like(X)
    pin_k[X[:, 0] == 0.0] = 1e6
    plan.set_contact_plane(1e5, [0.0, 0.2, 0.0], [0.0, 1.0, 0.0], mass[::3])
    plan.set_contact_sphere(1e5, [0.5, 0.5, 0.5], 0.3, mass[::3])
    x3, info3 = plan.newton(MAT, U.reshape(-1), x_tilde=U.reshape(-1), mass=mass, kin_scale=1.0 / h ** 2, f_ext=fext,
                            pin_k=pin_k.reshape(-1), pin_target=X.reshape(-1), max_iter=3, pcg_rtol=1e-12)
    plan.set_contact_plane(0.0)
    plan.set_contact_sphere(0.0)
    assert np.abs(x3 - x1).max() > 1e-6 * np.abs(x1).max()          # the extra terms do change the step
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "newton%d.npz" % r))
        lo, hi = int(d["lo"]) * 3, int(d["hi"]) * 3
        assert rel(d["mass"], mass[lo:hi]) < 1e-13
        assert list(d["alphas3"]) == list(info3["alphas"]) and rel(d["x3"], x3.ravel()[lo:hi]) < 1e-8
        assert list(d["alphas"]) == list(info["alphas"])
        assert rel(d["x"], x1.ravel()[lo:hi]) < 1e-8
        # two-level preconditioner: same iterate, fewer CG iterations
        assert rel(d["x2"], x1.ravel()[lo:hi]) < 1e-8 and list(d["alphas2"]) == list(info["alphas"])
        assert 1 < int(d["n_agg"]) <= 16 and int(d["its2"]) < int(d["its"])


# ------------------------------------------------------------------ 2 GPUs: the same PCG driven natively (capi_nccl.cu)
def _nccl_native_worker(rank, world, port, cells, tmp):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dim = len(cells)
    extent = tuple(1.0 for _ in cells)
    lay = sh.layout_grid_slab(cells, rank, world)
    shard = sh.Shard(lay, syn.grid_vertices(cells, extent, lay.l2g), device=rank, tile_elems=32)
    U = syn.jittered_state_rows(cells, extent, lay.l2g, sigma=0.2)
    mu, lam = syn.lame()
    shard.set_materials(mu, lam)
    rho, h = 1e3, 1e-2
    x_d = torch.from_numpy(U.reshape(-1).copy()).to(dev)
    mass_d = shard.lumped_mass_dofs(rho)
    fext_d = torch.zeros(lay.n_local, dim, dtype=torch.float64, device=dev)
    fext_d[:, 1] = -9.8
    fext_d = fext_d.reshape(-1) * mass_d
    kw = dict(x_tilde_d=x_d, mass_d=mass_d, kin_scale=1.0 / h ** 2, fext_d=fext_d, max_iter=3, pcg_rtol=1e-12)
    out = {}
    for tag, solver, n_agg in (("py", "python", 0), ("py_c", "python", 12), ("nat", "native", 0), ("nat_c", "native", 12),
                               ("natg", "native_graph", 0), ("natg_c", "native_graph", 12), ("p2e", "pcg2_eager", 0),
                               ("p2e_c", "pcg2_eager", 12), ("p2", "pcg2", 0), ("p2_c", "pcg2", 12),
                               ("pe", "peer_eager", 0), ("pe_c", "peer_eager", 12), ("pp", "peer", 0), ("pp_c", "peer", 12),
                               ("pp2", "peer", 0)):
        shard.solver = solver
        shard.set_coarse_space(n_agg)
        xs = x_d.clone()
        info = shard.newton_step(MAT, xs, **kw)
        o0, o1 = lay.own_lo * dim, lay.own_hi * dim
        out["x_" + tag] = xs.cpu().numpy()[o0:o1]
        out["its_" + tag] = info["pcg_iters"]
        out["alphas_" + tag] = np.asarray(info["alphas"])
    out["peer_ok"] = int(getattr(shard, "_peer", False))
    np.savez(os.path.join(tmp, "native%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_native_nccl_pcg_matches_python_loop(tmp_path):
    """The C++-driven distributed PCG (own NCCL communicator, no Python between the steps) runs the same kernels in the
    same order as Shard.pcg: same iteration counts and line-search steps, iterates equal to rounding."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    cells, world = (8, 5, 5), 2
    port = 29600 + ((os.getpid() + 17) % 1000)
    mp.spawn(_nccl_native_worker, args=(world, port, cells, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "native%d.npz" % r))
        for a, b in (("py", "nat"), ("py_c", "nat_c"), ("py", "natg"), ("py_c", "natg_c")):
            assert int(d["its_" + a]) == int(d["its_" + b]) and list(d["alphas_" + a]) == list(d["alphas_" + b])
            assert rel(d["x_" + b], d["x_" + a]) < 1e-12
        # the single-reduction solve (capi_pcg2.cu): other recurrences, same Krylov iterates to rounding -> same
        # line-search steps, iteration counts within a few, Newton iterates equal to the solver tolerance
        for a, b in (("py", "p2e"), ("py_c", "p2e_c"), ("py", "p2"), ("py_c", "p2_c")):
            assert list(d["alphas_" + a]) == list(d["alphas_" + b])
            assert abs(int(d["its_" + a]) - int(d["its_" + b])) <= 3 + int(d["its_" + a]) // 8
            assert rel(d["x_" + b], d["x_" + a]) < 1e-9
        assert np.array_equal(d["x_p2"], d["x_p2e"]) and np.array_equal(d["x_p2_c"], d["x_p2e_c"])   # graph replay = eager
        # peer-memory transport (NVLink stores + flags instead of NCCL calls): the same recurrences; the reduction sums
        # the ranks' partials in rank order instead of NCCL's order -> equal to rounding; reproducible run to run
        assert int(d["peer_ok"]) == 1
        for a, b in (("p2", "pe"), ("p2_c", "pe_c"), ("p2", "pp"), ("p2_c", "pp_c")):
            assert list(d["alphas_" + a]) == list(d["alphas_" + b])
            assert abs(int(d["its_" + a]) - int(d["its_" + b])) <= 2
            assert rel(d["x_" + b], d["x_" + a]) < 1e-10
        assert np.array_equal(d["x_pp"], d["x_pe"]) and np.array_equal(d["x_pp"], d["x_pp2"])


# ------------------------------------------------------------------ 2 GPUs: sharded reduced Hessian + r x r all-reduce
def _reduced_worker(rank, world, port, cells, tmp, interface):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    extent = tuple(1.0 for _ in cells)
    lay = sh.layout_grid_slab(cells, rank, world)
    Xl = syn.grid_vertices(cells, extent, lay.l2g)
    shard = sh.Shard(lay, Xl, device=rank, tile_elems=32, interface=interface)
    mu, lam = syn.lame()
    shard.set_materials(mu, lam)
    r = 24
    Bl = syn.cos_modes(Xl, r, seed=9, lo=np.zeros(len(cells)), hi=np.ones(len(cells)), n_total=lay.n_total)
    z = 0.05 * np.random.default_rng(5).standard_normal(r)
    E, g, H = shard.reduced(MAT, Bl, z, x0_local=Xl.reshape(-1))
    if rank == 0:
        np.savez(os.path.join(tmp, "reduced_%s.npz" % interface), E=E, g=g, H=H, nbytes=shard.reduced_allreduce_bytes)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("interface", ["recompute", "exchange"])
def test_sharded_reduced_hessian_matches_single_gpu(tmp_path, interface):
    """north_star: "... and allreduces the small r x r reduced Hessian" -- every rank contracts its own elements, one
    all-reduce of 1 + r + r^2 doubles; equal to the single-GPU reduced tier on the whole mesh and to B^T Q B of the oracle."""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import simkit_b200 as sk
    cells, world = (8, 5, 5), 2
    port = 29600 + ((os.getpid() + (53 if interface == "recompute" else 71)) % 1000)
    mp.spawn(_reduced_worker, args=(world, port, cells, str(tmp_path), interface), nprocs=world, join=True)
    d = np.load(os.path.join(str(tmp_path), "reduced_%s.npz" % interface))
    X, T = syn.make_mesh(cells)
    mu, lam = syn.lame()
    r = 24
    B = syn.cos_modes(X, r, seed=9, lo=np.zeros(3), hi=np.ones(3), n_total=X.shape[0])
    z = 0.05 * np.random.default_rng(5).standard_normal(r)
    plan = sk.MeshPlan(X=X, T=T, tile_elems=32)
    plan.set_materials(mu, lam, plan.volume())
    E1, g1, H1 = plan.reduced(MAT, B, z, x0=X.reshape(-1))
    assert abs(float(d["E"]) - E1) <= 1e-12 * abs(E1)
    assert rel(d["g"], g1) < 1e-10 and rel(d["H"], H1) < 1e-10
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    x = (B @ z).reshape(-1, 3) + X
    Qo = oe.hessian_x(MAT, x, Jo, mu, lam, volo)
    assert rel(d["H"], B.T @ (Qo @ B)) < 1e-10
    assert int(d["nbytes"]) == 8 * (1 + r + r * r)
