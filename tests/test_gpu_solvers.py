"""GPU parity tests of the linear solve, the Newton loop / integrators, the reduced tier and FST."""
import os

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import simkit_b200 as sk
from oracle import elasticity as oe
from simkit_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-10
ITER_TOL = 1e-8


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


def _system(cells, sigma=0.1, material="stable_neo_hookean"):
    X, T = syn.make_mesh(cells)
    dim = X.shape[1]
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=sigma)
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    g, vals = plan.gradient_hessian(material, U, mu, lam, vol, 1)
    mass = np.repeat(plan.vertex_masses(1e3), dim)
    return X, T, U, plan, g, vals, mass


@pytest.mark.parametrize("cells", [(8, 8, 8), (30, 30)])
def test_pcg_matches_direct_solve(cells):
    X, T, U, plan, g, vals, mass = _system(cells)
    h = 1e-2
    dadd = mass / h ** 2
    A = plan.csr_matrix(vals) + sps.diags(dadd)
    ref = spla.spsolve(A.tocsc(), -g.ravel())
    x, iters, relres = plan.pcg(vals, -g, diag_add=dadd, rtol=1e-12, max_iter=5000)
    assert relres <= 1e-12 and 0 < iters < 5000
    assert rel(x, ref) < 1e-9
    # generic CSR entry point (any scipy matrix), unsorted input
    Au = A.tocsc().tocsr()
    x2, it2, rr2 = sk.solve_sparse(Au, -g, rtol=1e-12, return_info=True)
    assert rel(x2, ref) < 1e-9
    # scalar Jacobi variant
    x3 = sk.solve_sparse(Au, -g, rtol=1e-12, block=1)
    assert rel(x3, ref) < 1e-9


def test_pcg_zero_rhs_and_spd_quadratic():
    X, T, U, plan, g, vals, mass = _system((4, 4, 4))
    x, iters, relres = plan.pcg(vals, np.zeros(plan.ndof), diag_add=mass * 1e4)
    assert iters == 0 and np.all(x == 0.0)


def test_dense_solve():
    rng = np.random.default_rng(0)
    for n in (1, 7, 200):
        A = rng.standard_normal((n, n))
        A = A @ A.T + n * np.eye(n)
        A[0, 0] = 1e-8 if n > 1 else A[0, 0]      # forces pivoting
        b = rng.standard_normal((n, 1))
        assert rel(sk.solve_dense(A, b), np.linalg.solve(A, b).ravel()) < 1e-10
    with pytest.raises(ValueError):
        sk.solve_dense(np.zeros((3, 3)), np.ones(3))


def test_newton_exact_on_spd_quadratic():
    """reference tests/test_solvers.py:32-54: Newton is exact in one step on a quadratic."""
    rng = np.random.default_rng(1)
    n = 30
    A = rng.standard_normal((n, n))
    A = A @ A.T + n * np.eye(n)
    b = rng.standard_normal((n, 1))
    E = lambda x: float(0.5 * x.T @ A @ x - b.T @ x)
    G = lambda x: A @ x - b
    for Hf in (lambda x: A, lambda x: sps.csr_matrix(A)):
        x, info = sk.newton_solver(np.zeros((n, 1)), E, G, Hf, max_iter=3, return_info=True)
        assert np.allclose(x, np.linalg.solve(A, b), atol=1e-8)
        assert info["iters"] <= 1 and set(info) == {"g", "dx", "alphas", "iters"}
    x0 = np.ones((n, 1))
    sk.newton_solver(x0, E, G, lambda x: A)
    assert np.all(x0 == 1.0)                                  # x0 is not mutated (newton.py:42)


@pytest.mark.parametrize("tag", ["step_tet", "step_tri"])
def test_backward_euler_against_golden(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T = g["X"], g["T"]
    dim = int(g["dim"])
    mu, lam, h = float(g["mu"]), float(g["lam"]), float(g["h"])
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Mv = sps.diags(g["mass_diag"]).tocsc()
    fg = g["fg"]
    for m in oe.MATERIALS:
        a = (mu,) if m == "arap" else (mu, lam)
        e_x, g_x, h_x = (getattr(sk, f"{m}_{k}_x") for k in ("energy", "gradient", "hessian"))

        def E(x):
            return e_x(x.reshape(-1, dim), J, *a, vol) - float((fg.T @ x).item())

        def G(x):
            return g_x(x.reshape(-1, dim), J, *a, vol) - fg

        def H(x):
            return h_x(x.reshape(-1, dim), J, *a, vol)

        # (1) generic closures through the drop-in integrator: GPU energies + GPU PCG
        x, info = sk.backward_euler(g[f"{m}_be_x_curr"], g[f"{m}_be_x_prev"], E, G, H, Mv, h, max_iter=3,
                                    return_info=True)
        assert np.array_equal(np.array(info["alphas"]), g[f"{m}_be_alphas"]), m
        assert info["iters"] == int(g[f"{m}_be_iters"])
        assert rel(x, g[f"{m}_be_x"]) < ITER_TOL, m
        assert rel(info["dx"][0], g[f"{m}_be_dx0"]) < ITER_TOL
        assert rel(info["g"][0], g[f"{m}_be_g0"]) < VAL_TOL
        # (2) device-resident step
        pot = sk.ElasticPotential(m, mu, lam, vol, J=J, dim=dim, f_ext=fg)
        x2, info2 = sk.backward_euler(g[f"{m}_be_x_curr"], g[f"{m}_be_x_prev"], pot.energy, pot.gradient,
                                      pot.hessian, Mv, h, max_iter=3, return_info=True, pcg_rtol=1e-12)
        assert list(info2["alphas"]) == list(g[f"{m}_be_alphas"]), m
        assert info2["iters"] == int(g[f"{m}_be_iters"])
        assert rel(x2, g[f"{m}_be_x"]) < ITER_TOL, m


def test_bdf2_and_pins_against_oracle():
    cells = (5, 4, 4)
    X, T = syn.make_mesh(cells)
    dim = 3
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    M = sps.kron(oe.massmatrix(X, T, rho), sps.identity(dim)).tocsc()
    fg = oe.gravity_force(X, T, -9.8, rho).reshape(-1, 1)
    pinned = np.where(X[:, 0] == 0.0)[0]
    pin_k = np.zeros((X.shape[0], dim))
    pin_k[pinned] = 1e8
    pin_k = pin_k.reshape(-1, 1)
    pin_t = X.reshape(-1, 1).copy()
    rng = np.random.default_rng(4)
    xs = [X.reshape(-1, 1) + 1e-3 * rng.standard_normal((X.size, 1)) for _ in range(4)]
    m = "stable_neo_hookean"

    def Eo(x):
        d = x - pin_t
        return oe.energy_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) - float((fg.T @ x).item()) + 0.5 * float((pin_k * d * d).sum())

    def Go(x):
        return oe.gradient_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) - fg + pin_k * (x - pin_t)

    def Ho(x):
        return oe.hessian_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) + sps.diags(pin_k.ravel())

    xo, io = oe.bdf2(xs[0], xs[1], xs[2], xs[3], Eo, Go, Ho, M, h, max_iter=3, return_info=True)
    pot = sk.ElasticPotential(m, mu, lam, vol, J=J, dim=dim, f_ext=fg, pin_k=pin_k, pin_target=pin_t)
    x1, i1 = sk.bdf2(xs[0], xs[1], xs[2], xs[3], pot.energy, pot.gradient, pot.hessian, M, h, max_iter=3,
                     return_info=True, pcg_rtol=1e-13)
    assert list(i1["alphas"]) == list(io["alphas"])
    assert rel(x1, xo) < ITER_TOL
    # the same potential used as plain closures by the drop-in (host loop + GPU PCG)
    x2 = sk.bdf2(xs[0], xs[1], xs[2], xs[3], lambda x: pot.energy(x), lambda x: pot.gradient(x),
                 lambda x: pot.hessian(x), M, h, max_iter=3, pcg_rtol=1e-13)
    assert rel(x2, xo) < ITER_TOL


# --------------------------------------------------------------------------- reduced tier / FST
@pytest.mark.parametrize("tag", ["reduced_tet", "reduced_tri"])
def test_reduced_against_golden(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T, B, z = g["X"], g["T"], g["B"], g["z"]
    dim = int(g["dim"])
    mu, lam, vol = float(g["mu"]), float(g["lam"]), g["vol"]
    Jo = oe.deformation_jacobian(X, T)
    JB = np.asarray(Jo @ B)
    Jx0 = np.asarray(Jo @ X.reshape(-1, 1))
    u = z.reshape(-1, dim)
    plan = sk.MeshPlan(X=X, T=T)
    for m in ("stable_neo_hookean", "arap"):
        a = (mu,) if m == "arap" else (mu, lam)
        # `_u` tier with a dense operator
        Hr = getattr(sk, f"{m}_hessian_u")(u, JB, Jx0, *a, vol)
        assert Hr.shape == g[f"{m}_Hr"].shape and rel(Hr, g[f"{m}_Hr"]) < VAL_TOL
        assert rel(getattr(sk, f"{m}_gradient_u")(u, JB, Jx0, *a, vol), g[f"{m}_gr"]) < VAL_TOL
        Er = getattr(sk, f"{m}_energy_u")(u, JB, Jx0, *a, vol)
        assert abs(Er - float(g[f"{m}_Er"])) <= 1e-12 * abs(float(g[f"{m}_Er"]))
        # basis form: JB rows formed on the fly from (plan, B)
        plan.set_materials(mu, lam, vol)
        E2, g2, H2 = plan.reduced(m, B, z, x0=X.reshape(-1))
        assert rel(H2, g[f"{m}_Hr"]) < VAL_TOL and rel(g2, g[f"{m}_gr"]) < VAL_TOL
        assert abs(E2 - float(g[f"{m}_Er"])) <= 1e-12 * abs(float(g[f"{m}_Er"]))
    # dispatcher `_z` tier (floor before vol) against the oracle
    pre = sk.ElasticEnergyZPrecomp(B, X.reshape(-1, 1), sps.identity(Jo.shape[0]), Jo, dim)
    Hz = sk.elastic_hessian_z(z, mu, lam, vol, "arap", pre)
    Ho = oe.hessian_x("arap", u, JB, mu, lam, vol, Jx_bar=Jx0, psd_before_vol=True)
    assert rel(Hz, Ho) < VAL_TOL
    # FST
    f = sk.fast_sandwich_transform_clustered(g["fst_A"], sps.csr_matrix(g["fst_B"]), g["fst_l"], dim=dim)
    assert rel(f.ARBs, g["fst_ARBs"]) < 1e-12
    assert rel(f(g["fst_r"]), g["fst_out"]) < 1e-12
    assert rel(f.eval(2 * g["fst_r"]), 2 * g["fst_out"]) < 1e-12      # linearity in r
    # sparse operators densified in chunks of elements (tiny chunk budget forces several chunks): same ARBs
    cls = sk.fast_sandwich_transform_clustered
    old = cls.CHUNK_BYTES
    cls.CHUNK_BYTES = 8 * dim * dim * (g["fst_A"].shape[0] + g["fst_B"].shape[1]) * 3     # three elements per chunk
    try:
        f2 = cls(sps.csc_matrix(g["fst_A"]), sps.csr_matrix(g["fst_B"]), g["fst_l"], dim=dim)
    finally:
        cls.CHUNK_BYTES = old
    assert rel(f2.ARBs, g["fst_ARBs"]) < 1e-12


def test_reduced_r200_against_oracle():
    """r = 200 modes (BASELINE config 4's reduced dimension) on a mesh the oracle can handle."""
    cells = (10, 10, 10)
    X, T = syn.make_mesh(cells)
    mu, lam = syn.lame()
    B = syn.smooth_modes(X, 200)
    z = 0.01 * np.random.default_rng(7).standard_normal(200)
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    plan.set_materials(mu, lam, vol)
    E, gr, Hr = plan.reduced("stable_neo_hookean", B, z, x0=X.reshape(-1))
    Jo = oe.deformation_jacobian(X, T)
    x = (B @ z).reshape(-1, 3) + X
    Qo = oe.hessian_x("stable_neo_hookean", x, Jo, mu, lam, vol)
    assert rel(Hr, B.T @ (Qo @ B)) < VAL_TOL
    assert rel(gr, B.T @ oe.gradient_x("stable_neo_hookean", x, Jo, mu, lam, vol)) < VAL_TOL
    assert np.abs(Hr - Hr.T).max() <= 1e-12 * np.abs(Hr).max()
    # basis kept resident on the device: same bits, no upload
    with pytest.raises(ValueError):
        plan.reduced("stable_neo_hookean", None, z, x0=X.reshape(-1))
    plan.set_basis(B)
    E2, g2, H2 = plan.reduced("stable_neo_hookean", None, z, x0=X.reshape(-1))
    assert E2 == E and np.array_equal(g2, gr) and np.array_equal(H2, Hr)
    plan.set_basis(None)
    with pytest.raises(ValueError):
        plan.reduced("stable_neo_hookean", None, z, x0=X.reshape(-1))


@pytest.mark.parametrize("r", [33, 64, 250])
def test_reduced_other_dimensions_against_oracle(r):
    """odd r (rows are not 16-byte multiples: the simple staging kernel), r = 64 (block pairs with unused
    tiles) and r = 250 (two passes of block pairs) through the basis form."""
    cells = (6, 5, 7)
    X, T = syn.make_mesh(cells)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    B = syn.smooth_modes(X, r, seed=5)
    z = 0.02 * np.random.default_rng(8).standard_normal(r)
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    plan.set_materials(mu, lam, vol)
    E, gr, Hr = plan.reduced("stable_neo_hookean", B, z, x0=X.reshape(-1))
    Jo = oe.deformation_jacobian(X, T)
    x = (B @ z).reshape(-1, 3) + X
    Qo = oe.hessian_x("stable_neo_hookean", x, Jo, mu, lam, vol)
    assert rel(Hr, B.T @ (Qo @ B)) < VAL_TOL
    assert rel(gr, B.T @ oe.gradient_x("stable_neo_hookean", x, Jo, mu, lam, vol)) < VAL_TOL
    Eo = oe.energy_x("stable_neo_hookean", x, Jo, mu, lam, vol)
    assert abs(E - Eo) <= 1e-12 * abs(Eo)


@pytest.mark.parametrize("cells", [(14, 12, 10), (40, 30)])
def test_two_level_preconditioner(cells):
    """Block-Jacobi + rigid-mode coarse correction: same solution as the direct solve, fewer iterations, and the
    Newton step it drives equals the block-Jacobi one to the iterate tolerance."""
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.2)
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    plan.set_materials(mu, lam, vol)
    g, vals = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    Q = plan.csr_matrix(vals)
    mass = np.repeat(plan.vertex_masses(1e3), dim)
    dadd = mass / 1e-2 ** 2
    rhs = -g.reshape(-1)
    x1, it1, rr1 = plan.pcg(vals, rhs, diag_add=dadd, rtol=1e-12)
    n_agg = plan.set_coarse_space(X, 27 if dim == 3 else 16)
    assert 1 < n_agg <= 64
    x2, it2, rr2 = plan.pcg(vals, rhs, diag_add=dadd, rtol=1e-12)
    import scipy.sparse.linalg as spla
    xo = spla.spsolve((Q + sps.diags(dadd)).tocsc(), rhs)
    assert rel(x1, xo) < 1e-9 and rel(x2, xo) < 1e-9
    assert it2 < it1
    x3, it3, _ = plan.pcg(vals, rhs, diag_add=dadd, rtol=1e-12)
    assert it3 == it2 and np.array_equal(x3, x2)               # deterministic
    xc = U.reshape(-1)
    fext = np.zeros((plan.n, dim))
    fext[:, 1] = -9.8
    fext = fext.reshape(-1) * mass
    xa, ia = plan.newton("stable_neo_hookean", xc, x_tilde=xc, mass=mass, kin_scale=1e4, f_ext=fext, max_iter=2, pcg_rtol=1e-12)
    plan.set_coarse_space(None)
    xb, ib = plan.newton("stable_neo_hookean", xc, x_tilde=xc, mass=mass, kin_scale=1e4, f_ext=fext, max_iter=2, pcg_rtol=1e-12)
    assert rel(xa, xb) < ITER_TOL and ia["alphas"] == ib["alphas"]
    assert ia["pcg_iters"] < ib["pcg_iters"]
    # through the drop-in surface: backward_euler with an ElasticPotential that carries the rest positions
    M = sps.diags(mass)
    pot = sk.ElasticPotential("stable_neo_hookean", mu, lam, vol, X=X, T=T, f_ext=fext, coarse=n_agg)
    xe = sk.backward_euler(xc.reshape(-1, 1), xc.reshape(-1, 1), pot.energy, pot.gradient, pot.hessian, M, 1e-2, max_iter=2,
                           pcg_rtol=1e-12)
    assert pot.plan.n_agg == n_agg and rel(xe, xb) < ITER_TOL
    # "auto": block-Jacobi first, coarse correction switched on once a step needed many CG iterations
    pot2 = sk.ElasticPotential("stable_neo_hookean", mu, lam, vol, X=X, T=T, f_ext=fext)
    old = sk.MeshPlan.COARSE_MIN_ITERS
    try:
        sk.MeshPlan.COARSE_MIN_ITERS = 5
        xe2 = sk.backward_euler(xc.reshape(-1, 1), xc.reshape(-1, 1), pot2.energy, pot2.gradient, pot2.hessian, M, 1e-2,
                                max_iter=2, pcg_rtol=1e-12)
        assert getattr(pot2.plan, "n_agg", 0) > 0 and rel(xe2, xb) < ITER_TOL
    finally:
        sk.MeshPlan.COARSE_MIN_ITERS = old


def test_two_level_degenerate_aggregates_fall_back():
    """Aggregates of collinear vertices make the rigid modes dependent (singular coarse matrix): the solve falls
    back to block-Jacobi instead of failing."""
    cells = (6, 6, 6)
    X, T = syn.make_mesh(cells)
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    vol = plan.volume()
    g, vals = plan.gradient_hessian("stable_neo_hookean", syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.1), mu, lam, vol, 1)
    dadd = np.full(plan.ndof, 10.0)
    x1, it1, _ = plan.pcg(vals, -g.reshape(-1), diag_add=dadd, rtol=1e-12)
    n_agg = plan.set_coarse_space(X, 343)          # one vertex per box: 6 modes on 3 dofs
    assert n_agg == 343
    x2, it2, _ = plan.pcg(vals, -g.reshape(-1), diag_add=dadd, rtol=1e-12)
    assert it2 == it1 and np.array_equal(x1, x2)
