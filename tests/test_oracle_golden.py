"""The oracle replays the frozen reference outputs (tests/golden, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe

MESHES = ["tet_s01", "tet_s04", "tri_s01", "tri_s04"]
TOL = 1e-12


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


def mats(g):
    return [m for m in oe.MATERIALS if f"{m}_E" in g.files]


@pytest.mark.parametrize("tag", MESHES)
def test_operators(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T = g["X"], g["T"]
    J = oe.canonical_csr(oe.deformation_jacobian(X, T))
    assert np.array_equal(J.indptr, g["J_indptr"]) and np.array_equal(J.indices, g["J_indices"])
    assert rel(J.data, g["J_data"]) < 1e-13
    assert rel(oe.volume(X, T), g["vol"]) < 1e-14
    R, S = oe.polar_svd(g["F"])
    assert rel(R, g["polar_R"]) < TOL and rel(S, g["polar_S"]) < TOL
    assert rel(oe.rotation_gradient_F(g["F"]), g["rotgrad"]) < 1e-11
    assert rel(oe.psd_project(g["psd_in"]), g["psd_proj"]) < TOL
    assert rel(oe.psd_project(g["psd_in"], "abs"), g["psd_abs"]) < TOL


@pytest.mark.parametrize("tag", MESHES)
def test_element_and_global_tiers(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T, U, mu, lam, vol, F = (g[k] for k in ("X", "T", "U", "mu", "lam", "vol", "F"))
    dim = int(g["dim"])
    n = X.shape[0]
    J = oe.deformation_jacobian(X, T)
    indptr, indices, _, _ = oe.structural_pattern(T, n, dim)
    for m in mats(g):
        assert rel(oe.energy_element_F(m, F, mu, lam), g[f"{m}_psi"]) < TOL
        assert rel(oe.gradient_element_F(m, F, mu, lam), g[f"{m}_P"]) < TOL
        assert rel(oe.hessian_element_F(m, F, mu, lam), g[f"{m}_He"]) < 1e-11
        assert abs(oe.energy_x(m, U, J, mu, lam, vol) - float(g[f"{m}_E"])) <= 1e-13 * abs(float(g[f"{m}_E"]))
        assert rel(oe.gradient_x(m, U, J, mu, lam, vol), g[f"{m}_g"]) < TOL
        xb = g["x_bar"]
        Jxb = J @ xb.reshape(-1, 1)
        assert abs(oe.energy_x(m, U - xb, J, mu, lam, vol, Jx_bar=Jxb) - float(g[f"{m}_E_u"])) <= 1e-12 * abs(float(g[f"{m}_E_u"]))
        assert rel(oe.gradient_x(m, U - xb, J, mu, lam, vol, Jx_bar=Jxb), g[f"{m}_g_u"]) < 1e-11
        for psd in (1, 0):
            k = f"{m}_Q_psd{psd}"
            Qref = sps.csr_matrix((g[k + "_data"], g[k + "_indices"], g[k + "_indptr"]), shape=(n * dim, n * dim))
            Q = oe.hessian_x(m, U, J, mu, lam, vol, psd=bool(psd))
            assert rel(Q.toarray(), Qref.toarray()) < 1e-11
            # reference pattern is a subset of the structural pattern
            S = sps.csr_matrix((np.ones(indices.shape[0]), indices, indptr), shape=Qref.shape)
            assert (abs(Qref) > 0).multiply(S).nnz == (abs(Qref) > 0).nnz
    # elastic dispatcher routes (elastic.py:75): floor before vol, linear elasticity projected too
    for m in ("arap", "linear_elasticity", "fcr", "macklin_mueller_neo_hookean"):
        k = f"{m}_Qdisp"
        Qref = sps.csr_matrix((g[k + "_data"], g[k + "_indices"], g[k + "_indptr"]), shape=(n * dim, n * dim))
        Q = oe.hessian_x(m, U, J, mu, lam, vol, psd=True, psd_before_vol=True)
        assert rel(Q.toarray(), Qref.toarray()) < 1e-11
        assert abs(oe.energy_x(m, U, J, mu, lam, vol) - float(g[f"{m}_Edisp"])) <= 1e-13 * abs(float(g[f"{m}_Edisp"]))
        assert rel(oe.gradient_x(m, U, J, mu, lam, vol), g[f"{m}_gdisp"]) < TOL
        assert rel(oe.psd_project(oe.hessian_element_F(m, F, mu, lam)), g[f"{m}_Hedisp"]) < 1e-11


def test_slot_map_definition(golden_dir):
    g = load(golden_dir, "tet_s01")
    T = g["T"]
    n = g["X"].shape[0]
    indptr, indices, bptr, bcol = oe.structural_pattern(T, n, 3)
    slot = oe.slot_map(T[:10], indptr, indices, 3)
    for e in range(10):
        for a in range(4):
            for b in range(4):
                r, c = T[e, a] * 3 + 1, T[e, b] * 3 + 2
                assert indices[slot[e, a, 1, b, 2]] == c
                assert indptr[r] <= slot[e, a, 1, b, 2] < indptr[r + 1]


@pytest.mark.parametrize("tag", ["step_tet", "step_tri"])
def test_backward_euler(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T = g["X"], g["T"]
    dim = int(g["dim"])
    mu, lam, h = float(g["mu"]), float(g["lam"]), float(g["h"])
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    Mv = sps.diags(g["mass_diag"]).tocsc()
    fg = g["fg"]
    for m in oe.MATERIALS:
        def E(x):
            return oe.energy_x(m, x.reshape(-1, dim), J, mu, lam, vol) - float((fg.T @ x).item())

        def G(x):
            return oe.gradient_x(m, x.reshape(-1, dim), J, mu, lam, vol) - fg

        def H(x):
            return oe.hessian_x(m, x.reshape(-1, dim), J, mu, lam, vol)

        x, info = oe.backward_euler(g[f"{m}_be_x_curr"], g[f"{m}_be_x_prev"], E, G, H, Mv, h, max_iter=3, return_info=True)
        assert np.array_equal(np.array(info["alphas"]), g[f"{m}_be_alphas"])
        assert info["iters"] == int(g[f"{m}_be_iters"])
        assert rel(x, g[f"{m}_be_x"]) < 1e-10
        assert rel(info["dx"][0], g[f"{m}_be_dx0"]) < 1e-9


@pytest.mark.parametrize("tag", ["reduced_tet", "reduced_tri"])
def test_reduced_and_fst(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T, B, z = g["X"], g["T"], g["B"], g["z"]
    dim = int(g["dim"])
    mu, lam, vol = float(g["mu"]), float(g["lam"]), g["vol"]
    J = oe.deformation_jacobian(X, T)
    JB = np.asarray(J @ B)
    Jx0 = np.asarray(J @ X.reshape(-1, 1))
    u = z.reshape(-1, dim)
    for m in ("stable_neo_hookean", "arap"):
        Hr = oe.hessian_x(m, u, JB, mu, lam, vol, Jx_bar=Jx0)
        assert rel(Hr, g[f"{m}_Hr"]) < 1e-11
        assert rel(oe.gradient_x(m, u, JB, mu, lam, vol, Jx_bar=Jx0), g[f"{m}_gr"]) < 1e-11
    ARBs = oe.fst_precompute(g["fst_A"], g["fst_B"], g["fst_l"], dim)
    assert rel(ARBs, g["fst_ARBs"]) < TOL
    assert rel(oe.fst_eval(ARBs, g["fst_r"], dim), g["fst_out"]) < TOL


@pytest.mark.parametrize("cells,h", [((6, 5, 4), 1e-2), ((6, 5, 4), 1.0), ((20, 16), 1e-2)])
def test_single_reduction_pcg_prototype(cells, h):
    """The Chronopoulos-Gear (one reduction per iteration) PCG planned for the distributed solve: same iterates as
    the textbook loop to rounding -- same iteration count (+-2) and the same solution on a backward-Euler system."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    from simkit_b200 import synthetic as syn
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.2)
    mu, lam = syn.lame()
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    H = oe.hessian_x("stable_neo_hookean", U, J, mu, lam, vol) + sps.kron(oe.massmatrix(X, T, 1e3), sps.identity(dim)) / h ** 2
    g = oe.gradient_x("stable_neo_hookean", U, J, mu, lam, vol).ravel()
    x1, i1 = oe.block_jacobi_cg(H, -g, dim, rtol=1e-10)
    x2, i2 = oe.block_jacobi_cg_single_reduction(H, -g, dim, rtol=1e-10)
    xd = spla.spsolve(H.tocsc(), -g)
    scale = np.abs(xd).max()
    assert abs(i1 - i2) <= 2
    assert np.abs(x2 - xd).max() <= 10 * max(np.abs(x1 - xd).max(), 1e-12 * scale)


@pytest.mark.parametrize("cells,h", [((8, 8, 8), 1e-2), ((20, 14), 1e-2)])
def test_two_level_single_reduction_pcg_prototype(cells, h):
    """numpy statement-for-statement prototype of csrc/capi_pcg2.cu (bootstrap pass, Chronopoulos-Gear recurrences, the
    restricted residual carried by its own recurrence instead of a second all-reduce): same solution as the direct
    solve, fewer iterations than block-Jacobi alone, and block-Jacobi-only equals the earlier prototype."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    from simkit_b200 import synthetic as syn
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, tuple(1.0 for _ in cells), sigma=0.2)
    mu, lam = syn.lame()
    J, vol = oe.deformation_jacobian(X, T), oe.volume(X, T)
    H = oe.hessian_x("stable_neo_hookean", U, J, mu, lam, vol) + sps.kron(oe.massmatrix(X, T, 1e3), sps.identity(dim)) / h ** 2
    g = oe.gradient_x("stable_neo_hookean", U, J, mu, lam, vol).ravel()
    xd = spla.spsolve(H.tocsc(), -g)
    scale = np.abs(xd).max()
    x0, i0 = oe.block_jacobi_cg_single_reduction(H, -g, dim, rtol=1e-10)
    x1, i1 = oe.two_level_cg_single_reduction(H, -g, dim, None, rtol=1e-10)
    assert i1 == i0 and np.abs(x1 - x0).max() <= 1e-12 * scale
    nb = 3
    ib = np.minimum((X / X.max(0) * nb).astype(int), nb - 1)
    agg = ib[:, 0]
    for a in range(1, dim):
        agg = agg * nb + ib[:, a]
    P = oe.rigid_mode_prolongator(X, agg)
    x2, i2 = oe.two_level_cg_single_reduction(H, -g, dim, P, rtol=1e-10)
    assert i2 < i1
    assert np.abs(x2 - xd).max() <= 1e-8 * scale
    # the recurrence of the restricted residual is exact: P^T (b - H x) at the end equals what it carried (to rounding)
    r_true = -g - H @ x2
    assert np.linalg.norm(r_true) <= 2e-10 * np.linalg.norm(g)
