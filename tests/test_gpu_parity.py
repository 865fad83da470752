"""GPU parity tests: the CUDA path (through the Python drop-in surface -> ctypes -> C ABI) against
the numpy oracle on seeded inputs and against the frozen reference outputs in tests/golden.

Tolerances (BASELINE.json north_star): CSR row pointers / column indices / slot map bit-exact;
energies, gradients and Hessian values 1e-10 relative (max-norm); Newton iterates 1e-8 relative.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sps

import simkit_b200 as sk
from oracle import elasticity as oe
from simkit_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-10
ENERGY_TOL = 1e-12
ITER_TOL = 1e-8
MESHES = ["tet_s01", "tet_s04", "tri_s01", "tri_s04"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    return np.load(os.path.join(golden_dir, tag + ".npz"))


def margs(m, mu, lam):
    return (mu,) if m == "arap" else (mu, lam)


def fn(m, kind, tier):
    return getattr(sk, f"{m}_{kind}{tier}")


# --------------------------------------------------------------------------- structure
@pytest.mark.parametrize("cells", [(4, 3, 5), (9, 7)])
def test_pattern_and_slot_map_bit_exact(cells):
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(5)
    perm = rng.permutation(X.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    X, T = X[perm], inv[T]
    dim = X.shape[1]
    plan = sk.MeshPlan(X=X, T=T, tile_elems=32)
    indptr, indices, bptr, bcol = oe.structural_pattern(T, X.shape[0], dim)
    ip, ix = plan.csr_pattern()
    assert ip.dtype == np.int32 and ix.dtype == np.int32
    assert np.array_equal(ip, indptr) and np.array_equal(ix, indices)
    bp, bc = plan.block_pattern()
    assert np.array_equal(bp, bptr) and np.array_equal(bc, bcol)
    slot = plan.slot_map()
    assert slot.dtype == np.int32
    assert np.array_equal(slot, oe.slot_map(T, indptr, indices, dim))
    assert rel(plan.element_D(), oe.element_D(X, T)) < 1e-13
    assert rel(plan.volume(), oe.volume(X, T)) < 1e-14
    assert rel(plan.vertex_masses(1e3), oe.vertex_masses(X, T, 1e3)) < 1e-13


# --------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("tag", MESHES)
def test_golden_operators(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T, F = g["X"], g["T"], g["F"]
    dim = int(g["dim"])
    J = sk.deformation_jacobian(X, T)
    assert sps.isspmatrix_csc(J) and J.shape == (T.shape[0] * dim * dim, X.shape[0] * dim)
    Jc = oe.canonical_csr(J)
    assert np.array_equal(Jc.indptr, g["J_indptr"]) and np.array_equal(Jc.indices, g["J_indices"])
    assert rel(Jc.data, g["J_data"]) < 1e-12
    assert rel(sk.volume(X, T), g["vol"]) < 1e-13
    R, S = sk.polar_svd(F)
    assert rel(R, g["polar_R"]) < VAL_TOL and rel(S, g["polar_S"]) < VAL_TOL
    assert rel(sk.rotation_gradient_F(F), g["rotgrad"]) < VAL_TOL
    P = sk.psd_project(g["psd_in"])
    assert P.shape == g["psd_proj"].shape and rel(P, g["psd_proj"]) < VAL_TOL
    assert rel(sk.psd_project(g["psd_in"], "abs"), g["psd_abs"]) < VAL_TOL
    P1 = sk.psd_project(g["psd_in"][0])
    assert P1.ndim == 3                                   # promoted and stays 3-D (psd_project.py:28-29)
    with pytest.raises(NameError):
        sk.polar_svd(F, flip=False)                       # reference quirk kept (polar_svd.py:80-84)


@pytest.mark.parametrize("tag", MESHES)
def test_golden_energies(golden_dir, tag):
    g = load(golden_dir, tag)
    X, T, U, mu, lam, vol, F = (g[k] for k in ("X", "T", "U", "mu", "lam", "vol", "F"))
    dim = int(g["dim"])
    n = X.shape[0]
    J = sk.deformation_jacobian(X, T)
    Jplain = oe.deformation_jacobian(X, T)                # a scipy J that carries no plan
    xb = g["x_bar"]
    Jxb = Jplain @ xb.reshape(-1, 1)
    for m in oe.MATERIALS:
        if f"{m}_E" not in g.files:
            continue
        a = margs(m, mu, lam)
        assert rel(fn(m, "energy", "_element_F")(F, *a), g[f"{m}_psi"]) < ENERGY_TOL
        assert rel(fn(m, "gradient", "_element_F")(F, *a), g[f"{m}_P"]) < VAL_TOL
        assert rel(fn(m, "hessian", "_element_F")(F, *a), g[f"{m}_He"]) < VAL_TOL
        E = fn(m, "energy", "_x")(U, J, *a, vol)
        assert isinstance(E, float) and abs(E - float(g[f"{m}_E"])) <= ENERGY_TOL * abs(float(g[f"{m}_E"]))
        gr = fn(m, "gradient", "_x")(U, J, *a, vol)
        assert gr.shape == (n * dim, 1) and rel(gr, g[f"{m}_g"]) < VAL_TOL
        for psd in (1, 0):
            k = f"{m}_Q_psd{psd}"
            Qref = sps.csr_matrix((g[k + "_data"], g[k + "_indices"], g[k + "_indptr"]), shape=(n * dim, n * dim))
            Q = fn(m, "hessian", "_x")(U, J, *a, vol, psd=bool(psd))
            assert sps.isspmatrix_csr(Q) and Q.indices.dtype == np.int32
            assert rel(Q.toarray(), Qref.toarray()) < VAL_TOL
            # pattern(ref) subset of pattern(ours); our extra slots are exactly 0.0
            mask = sps.csr_matrix((np.ones(Qref.nnz), Qref.indices, Qref.indptr), shape=Q.shape)
            extra = Q - Q.multiply(mask)
            assert extra.nnz == 0 or abs(extra).max() == 0.0
            if Qref.nnz == Q.nnz:
                assert np.array_equal(Q.indptr, Qref.indptr) and np.array_equal(Q.indices, Qref.indices)
        # plain scipy operator -> plan recovered from J; _u tier
        Eu = fn(m, "energy", "_u")(U - xb, Jplain, Jxb, *a, vol)
        assert abs(Eu - float(g[f"{m}_E_u"])) <= 1e-11 * abs(float(g[f"{m}_E_u"]))
        assert rel(fn(m, "gradient", "_u")(U - xb, Jplain, Jxb, *a, vol), g[f"{m}_g_u"]) < VAL_TOL
        # self-contained tier
        Es = fn(m, "energy", "")(X, T, *a, U)
        assert abs(Es - float(g[f"{m}_E"])) <= ENERGY_TOL * abs(float(g[f"{m}_E"]))
    with pytest.raises(ValueError):
        sk.elastic_hessian_x(U, J, mu, lam, vol, "no-such-material")


DISPATCH = (("arap", "arap"), ("linear_elasticity", "linear-elasticity"), ("fcr", "fcr"),
            ("macklin_mueller_neo_hookean", "macklin-mueller-neo-hookean"))


@pytest.mark.parametrize("tag", MESHES)
def test_golden_elastic_dispatcher(golden_dir, tag):
    """energies/elastic.py: every routed material through every tier of the string dispatcher."""
    g = load(golden_dir, tag)
    X, T, U, mu, lam, vol, F = (g[k] for k in ("X", "T", "U", "mu", "lam", "vol", "F"))
    dim = int(g["dim"])
    n = X.shape[0]
    J = sk.deformation_jacobian(X, T)
    xb = g["x_bar"]
    Jxb = J @ xb.reshape(-1, 1)
    for m, name in DISPATCH:
        k = f"{m}_Qdisp"
        Qref = sps.csr_matrix((g[k + "_data"], g[k + "_indices"], g[k + "_indptr"]), shape=(n * dim, n * dim))
        Q = sk.elastic_hessian_x(U, J, mu, lam, vol, name, psd=True)
        assert rel(Q.toarray(), Qref.toarray()) < VAL_TOL
        Qs = sk.elastic_hessian(X, T, mu, lam, name, U=U, psd=True)          # self-contained: same route
        assert rel(Qs.toarray(), Qref.toarray()) < VAL_TOL
        Eref = float(g[f"{m}_Edisp"])
        assert abs(sk.elastic_energy_x(U, J, mu, lam, vol, name) - Eref) <= ENERGY_TOL * abs(Eref)
        assert abs(sk.elastic_energy(X, T, mu, lam, name, U=U) - Eref) <= ENERGY_TOL * abs(Eref)
        assert rel(sk.elastic_gradient_x(U, J, mu, lam, vol, name), g[f"{m}_gdisp"]) < VAL_TOL
        assert rel(sk.elastic_gradient(X, T, mu, lam, name, U=U), g[f"{m}_gdisp"]) < VAL_TOL
        assert rel(sk.elastic_hessian_element_F(F, mu, lam, name, psd=True), g[f"{m}_Hedisp"]) < VAL_TOL
        assert rel(sk.elastic_energy_element_F(F, mu, lam, name), g[f"{m}_psi"]) < ENERGY_TOL
        assert rel(sk.elastic_gradient_element_F(F, mu, lam, name), g[f"{m}_P"]) < VAL_TOL
        # _u tier dispatches to the per-material modules (floor after vol, LE unprojected)
        Eu = sk.elastic_energy_u(U - xb, J, Jxb, mu, lam, vol, name)
        assert abs(Eu - float(g[f"{m}_E_u"])) <= 1e-11 * abs(float(g[f"{m}_E_u"]))
        assert rel(sk.elastic_gradient_u(U - xb, J, Jxb, mu, lam, vol, name), g[f"{m}_g_u"]) < VAL_TOL
    # stretch (S) tier: ARAP and Macklin-Mueller, full and compact stretches
    for name, ts in (("arap", "arap"), ("macklin-mueller-neo-hookean", "mm")):
        for form, S in (("full", g["polar_S"]), ("compact", g["S_compact"])):
            Eref = float(g[f"{ts}_S_{form}_E"])
            assert abs(sk.elastic_energy_S(S, mu, lam, vol, name) - Eref) <= ENERGY_TOL * max(abs(Eref), 1e-300)
            gs = sk.elastic_gradient_S(S, mu, lam, vol, name)
            assert gs.shape == g[f"{ts}_S_{form}_g"].shape and rel(gs, g[f"{ts}_S_{form}_g"]) < VAL_TOL
            Hs = sk.elastic_hessian_S(S, mu, lam, vol, name)
            assert Hs.shape == g[f"{ts}_S_{form}_H"].shape and rel(Hs, g[f"{ts}_S_{form}_H"]) < VAL_TOL
    with pytest.raises(ValueError):
        sk.elastic_energy_S(g["polar_S"], mu, lam, vol, "fcr")


# --------------------------------------------------------------------------- oracle, config C1 size
@pytest.mark.parametrize("material", oe.MATERIALS)
def test_c1_cube_against_oracle(material):
    """BASELINE config 1 mesh (20^3 cells, 48k tets), heterogeneous material, sigma 0.1."""
    cells = (20, 20, 20)
    X, T = syn.make_mesh("C1")
    U = syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.1)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo = oe.deformation_jacobian(X, T)
    volo = oe.volume(X, T)
    a = margs(material, mu, lam)
    E = fn(material, "energy", "_x")(U, J, *a, vol)
    Eo = oe.energy_x(material, U, Jo, mu, lam, volo)
    assert abs(E - Eo) <= ENERGY_TOL * abs(Eo)
    assert rel(fn(material, "gradient", "_x")(U, J, *a, vol), oe.gradient_x(material, U, Jo, mu, lam, volo)) < VAL_TOL
    Q = fn(material, "hessian", "_x")(U, J, *a, vol)
    Qo = oe.canonical_csr(oe.hessian_x(material, U, Jo, mu, lam, volo))
    d = (Q - Qo)
    assert abs(d).max() / abs(Qo).max() < VAL_TOL
    if Qo.nnz == Q.nnz:
        assert np.array_equal(Q.indptr, Qo.indptr) and np.array_equal(Q.indices, Qo.indices)


def test_c2_square_against_oracle():
    """BASELINE config 2 mesh (316^2 cells, ~200k triangles), neo-Hookean."""
    X, T = syn.make_mesh("C2")
    U = syn.jittered_state(X, (316, 316), (1.0, 1.0), sigma=0.1)
    mu, lam = syn.lame()
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    assert rel(sk.neo_hookean_gradient_x(U, J, mu, lam, vol), oe.gradient_x("neo_hookean", U, Jo, mu, lam, volo)) < VAL_TOL
    Q = sk.neo_hookean_hessian_x(U, J, mu, lam, vol)
    Qo = oe.canonical_csr(oe.hessian_x("neo_hookean", U, Jo, mu, lam, volo))
    assert abs(Q - Qo).max() / abs(Qo).max() < VAL_TOL


def test_deterministic_bitwise():
    X, T = syn.make_mesh((12, 11, 10))
    U = syn.jittered_state(X, (12, 11, 10), (1.0, 1.0, 1.0), sigma=0.3)
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    g1, v1 = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    g2, v2 = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    assert np.array_equal(g1, g2) and np.array_equal(v1, v2)
    # a different tile size changes the summation tree only: equal to rounding, not bits
    plan2 = sk.MeshPlan(X=X, T=T, tile_elems=64)
    g3, v3 = plan2.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    assert rel(v3, v1) < 1e-13 and rel(g3, g1) < 1e-12


def test_pipelined_many_rounds():
    """A mesh that takes the persistent pipelined kernel several rounds (staging buffers handed between the
    groups, chunks of phase 2 pulled dynamically): same matrix as the one-CTA-per-tile kernel (other tile
    size => other summation tree: equal to rounding), bitwise reproducible, and equal to the oracle on a
    closed sub-mesh."""
    cells = (36, 36, 36)
    X, T = syn.make_mesh(cells)
    U = syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.3)
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    plan = sk.MeshPlan(X=X, T=T)                       # 2187 tiles of 128: ~5 rounds of 148 x 3 tiles
    g1, v1 = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    g2, v2 = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    assert np.array_equal(g1, g2) and np.array_equal(v1, v2)
    plan64 = sk.MeshPlan(X=X, T=T, tile_elems=64)      # not the pipelined shape: one CTA per tile
    g3, v3 = plan64.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    assert rel(v1, v3) < 1e-13 and rel(g1, g3) < 1e-12
    Q = plan.csr_matrix(v1)
    assert abs(Q - Q.T).max() == 0.0
    nl = 3                                             # first 3 layers of cells: closed sub-mesh
    tsub = 6 * nl * 36 * 36
    Ts = T[:tsub]
    nsub = int(Ts.max()) + 1
    Jo, volo = oe.deformation_jacobian(X[:nsub], Ts), oe.volume(X[:nsub], Ts)
    Qo = oe.canonical_csr(oe.hessian_x("stable_neo_hookean", U[:nsub], Jo, mu[:tsub], lam[:tsub], volo))
    nfull = nl * 37 * 37
    d = Q[: nfull * 3][:, : nsub * 3] - Qo[: nfull * 3]
    assert abs(d).max() / abs(Qo).max() < VAL_TOL


def test_argument_errors():
    X, T = syn.make_mesh((3, 3, 3))
    plan = sk.MeshPlan(X=X, T=T)
    with pytest.raises(ValueError):
        plan.energy("stable_neo_hookean", X, np.ones(5), 1.0, None)          # ragged material array
    with pytest.raises(ValueError):
        plan.energy("stable_neo_hookean", X[:-1], 1.0, 1.0, None)            # wrong x size
    with pytest.raises(ValueError):
        sk.MeshPlan(X=X, T=np.array([[0, 1, 2, X.shape[0]]]))                # index out of range
    with pytest.raises(ValueError):
        sk.MeshPlan(X=X[:, :2], T=T)                                         # 4 corners in 2D


# --------------------------------------------------------------------------- full-size properties
def test_c5_size_properties():
    """BASELINE config 5 (139^3 cells, 16.1M tets): size-independent properties of the assembled
    operator, where the oracle cannot run (8 GB RSS per 1M tets)."""
    cells = (139, 139, 139)
    X, T = syn.make_mesh("C5")
    T = T.astype(np.int32)
    U = syn.jittered_state(X, cells, (1.0, 1.0, 1.0), sigma=0.1)
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    assert plan.t == 16113714 and plan.n == 2744000
    g, vals = plan.gradient_hessian("stable_neo_hookean", U, mu, lam, None, 1)
    assert np.isfinite(g).all() and np.isfinite(vals).all()
    # translation invariance: internal forces sum to zero per axis; H annihilates rigid translations
    gs = g.reshape(-1, 3).sum(axis=0)
    assert np.abs(gs).max() <= 1e-9 * np.abs(g).sum()
    Q = plan.csr_matrix(vals)
    for ax in range(3):
        tvec = np.zeros((plan.n, 3))
        tvec[:, ax] = 1.0
        r = Q @ tvec.reshape(-1)
        assert np.abs(r).max() <= 1e-9 * np.abs(vals).max()
    # symmetry, on a sampled set of rows
    rows = np.random.default_rng(0).choice(plan.ndof, 2000, replace=False)
    sub = Q[rows][:, rows]
    assert abs(sub - sub.T).max() <= 1e-12 * abs(vals).max()
    # PSD projection: every element block is PSD => x^T Q x >= 0
    for seed in range(3):
        v = np.random.default_rng(seed).standard_normal(plan.ndof)
        assert v @ (Q @ v) >= 0.0
    # sampled sub-block parity with the oracle: the first 3 layers of cells form a closed sub-mesh
    nsub_cells = 3
    tsub = 6 * nsub_cells * 139 * 139
    Ts = T[:tsub].astype(np.int64)
    nsub = int(Ts.max()) + 1
    Xs, Us = X[:nsub], U[:nsub]
    Jo, volo = oe.deformation_jacobian(Xs, Ts), oe.volume(Xs, Ts)
    go = oe.gradient_x("stable_neo_hookean", Us, Jo, mu, lam, volo)
    # rows of vertices in the first (nsub_cells-1) complete layers see only sub-mesh elements
    nfull = (nsub_cells) * 140 * 140
    assert rel(g[: nfull * 3], go[: nfull * 3]) < VAL_TOL
    Qo = oe.canonical_csr(oe.hessian_x("stable_neo_hookean", Us, Jo, mu, lam, volo))
    d = Q[: nfull * 3][:, : nsub * 3] - Qo[: nfull * 3]
    assert abs(d).max() / abs(Qo).max() < VAL_TOL
