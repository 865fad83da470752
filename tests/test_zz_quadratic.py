"""General sparse quadratic term and the Dirichlet penalty: oracle and host logic vs the frozen
reference outputs on the CPU; on the GPU the drop-in ``quadratic_*`` functions and a backward-Euler step of stable
neo-Hookean + gravity + quadratic term, host-callable and device-resident."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe

TAGS = ["quadratic_tet", "quadratic_tri"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    nd = g["X"].shape[0] * int(g["dim"])
    Q = sps.csr_matrix((g["Q_data"], g["Q_indices"], g["Q_indptr"]), shape=(nd, nd))
    return g, Q


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_quadratic(golden_dir, tag):
    g, Q = load(golden_dir, tag)
    x = g["U"].reshape(-1, 1)
    assert abs(oe.quadratic_energy(x, Q, g["b"]) - float(g["E"])) <= 1e-13 * abs(float(g["E"]))
    assert rel(oe.quadratic_gradient(x, Q, g["b"]), g["g"]) < 1e-13
    assert oe.quadratic_hessian(Q) is Q
    Qd, bd = oe.dirichlet_penalty(g["bI"], g["y"], g["X"].shape[0], g["gamma"])
    assert rel(Qd.toarray(), g["Qd"]) == 0.0 and rel(bd, g["bd"]) == 0.0
    Qs, bs = oe.dirichlet_penalty(g["bI"], g["y"], g["X"].shape[0], 1e6)
    assert rel(Qs.toarray(), g["Qs"]) == 0.0 and rel(bs, g["bs"]) == 0.0


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_step_with_quadratic(golden_dir, tag):
    """The oracle's backward Euler with the quadratic term reproduces the reference iterate."""
    g, Q = load(golden_dir, tag)
    X, T, dim = g["X"], g["T"], int(g["dim"])
    mu, lam, h, fg, b = float(g["mu"]), float(g["lam"]), float(g["h"]), g["fg"], g["b"]
    J = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T)
    Md = sps.kron(sps.diags(g["mass"]), sps.identity(dim)).tocsc()

    def En(x):
        return oe.energy_x("stable_neo_hookean", x.reshape(-1, dim), J, mu, lam, vol) - float((fg.T @ x).item()) + oe.quadratic_energy(x, Q, b)

    def Gr(x):
        return oe.gradient_x("stable_neo_hookean", x.reshape(-1, dim), J, mu, lam, vol) - fg + oe.quadratic_gradient(x, Q, b)

    def He(x):
        return oe.hessian_x("stable_neo_hookean", x.reshape(-1, dim), J, mu, lam, vol, psd=True) + Q

    x1, info = oe.backward_euler(g["U"].reshape(-1, 1), X.reshape(-1, 1), En, Gr, He, Md, h, max_iter=3, return_info=True)
    assert list(info["alphas"]) == list(g["be_alphas"]) and rel(x1, g["be_x"]) < 1e-10


@pytest.mark.parametrize("tag", TAGS)
def test_dirichlet_penalty_host(golden_dir, tag):
    """``simkit_b200.dirichlet_penalty`` is set-up code on the host (no GPU call): same tuples as the reference."""
    from simkit_b200.dirichlet_penalty import dirichlet_penalty
    g, _ = load(golden_dir, tag)
    nv = g["X"].shape[0]
    Qd, bd = dirichlet_penalty(g["bI"], g["y"], nv, g["gamma"])
    assert sps.issparse(Qd) and rel(Qd.toarray(), g["Qd"]) == 0.0 and bd.shape == g["bd"].shape and rel(bd, g["bd"]) == 0.0
    Qs, bs, SG = dirichlet_penalty(g["bI"].reshape(-1, 1), g["y"], nv, 1e6, return_SGamma=True)
    assert rel(Qs.toarray(), g["Qs"]) == 0.0 and rel(bs, g["bs"]) == 0.0
    (b_only,) = dirichlet_penalty(g["bI"], g["y"] + 1.0, nv, 1e6, only_b=True, SGamma=SG)
    assert rel(b_only, -(SG @ (g["y"] + 1.0).reshape(-1, 1))) == 0.0
    with pytest.raises(AssertionError):
        dirichlet_penalty(g["bI"], g["y"].reshape(-1), nv, 1e6)


@pytest.mark.parametrize("tag", TAGS)
def test_value_positions_host_replay(golden_dir, tag):
    """csr_value_position (the map skb_newton_set_quadratic builds on the device) replayed on the host: every entry of
    Q lands on the entry of the canonical CSR pattern with the same (row, col); entries outside give -1."""
    import hostsim
    g, Q = load(golden_dir, tag)
    X, T, dim = g["X"], g["T"], int(g["dim"])
    r = hostsim.run(X, T, 0, 1, g["U"], float(g["mu"]), float(g["lam"]))
    H = hostsim.csr_from_blocks(r["bptr"], r["bcol"], r["vals"], X.shape[0], dim)
    Qc = Q.tocoo()
    pos = hostsim.value_positions(r["bptr"], r["bcol"], Qc.row, Qc.col, dim)
    assert (pos >= 0).all() and np.unique(pos).size == pos.size
    rows_of = np.repeat(np.arange(H.shape[0]), np.diff(H.indptr))
    assert np.array_equal(rows_of[pos], Qc.row) and np.array_equal(H.indices[pos], Qc.col)
    # adding Q through the map == adding the matrices
    vals = r["vals"].copy()
    vals[pos] += Qc.data
    H2 = sps.csr_matrix((vals, H.indices, H.indptr), shape=H.shape)
    assert rel(H2.toarray(), (H + Q).toarray()) < 1e-15
    # an entry between two vertices that share no element is outside the pattern
    Hp = sps.csr_matrix((np.ones_like(H.data), H.indices, H.indptr), shape=H.shape).toarray() > 0
    out_r, out_c = np.where(~Hp)
    if out_r.size:
        assert (hostsim.value_positions(r["bptr"], r["bcol"], out_r[:50], out_c[:50], dim) == -1).all()


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_dirichlet_laplacian(golden_dir, tag):
    g, _ = load(golden_dir, tag)
    X, T = g["X"], g["T"]
    assert rel(oe.dirichlet_laplacian(X, T, g["mu_het"]).toarray(), g["L_het"]) < 1e-13
    assert rel(oe.dirichlet_laplacian(X, T, 2.5, vector=True).toarray(), g["Lv_scalar"]) < 1e-13


@pytest.mark.parametrize("tag", TAGS)
def test_dirichlet_laplacian_host_replay(golden_dir, tag):
    """The route ``simkit_b200.dirichlet_laplacian`` takes on the GPU -- the constant linear-elasticity block with
    ``lam = -mu``, no projection, coordinate blocks averaged -- replayed with the kernel phase functions on the host."""
    import hostsim
    g, _ = load(golden_dir, tag)
    X, T, dim = g["X"], g["T"], int(g["dim"])
    n = X.shape[0]
    mu = g["mu_het"]
    r = hostsim.run(X, T, 4, 0, X, mu, -mu)
    H = hostsim.csr_from_blocks(r["bptr"], r["bcol"], r["vals"], n, dim).tocsc()
    L = sum(H[np.arange(n) * dim + i, :][:, np.arange(n) * dim + i] for i in range(dim)) / dim
    assert rel(L.toarray(), g["L_het"]) < 1e-13
    # every coordinate block is the Laplacian on its own
    for i in range(dim):
        Ii = np.arange(n) * dim + i
        assert rel(H[Ii, :][:, Ii].toarray(), g["L_het"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_dirichlet_laplacian(golden_dir, tag):
    import simkit_b200 as sk
    g, _ = load(golden_dir, tag)
    X, T = g["X"], g["T"]
    L = sk.dirichlet_laplacian(X, T, g["mu_het"])
    assert sps.issparse(L) and L.format == "csc" and L.shape == g["L_het"].shape and rel(L.toarray(), g["L_het"]) < 1e-12
    Lv = sk.dirichlet_laplacian(X, T, 2.5, vector=True)
    assert Lv.shape == g["Lv_scalar"].shape and rel(Lv.toarray(), g["Lv_scalar"]) < 1e-12
    with pytest.raises(AssertionError):
        sk.dirichlet_laplacian(X, T, np.ones(3))


def test_dirichlet_penalty_pinned_state_is_stationary():
    """tests/test_dirichlet_penalty.py:10-32 of the reference: shapes, zero gradient and zero energy (up to the dropped
    constant) at the pinned state -- here for the host drop-in and the oracle."""
    from simkit_b200.dirichlet_penalty import dirichlet_penalty
    nv, dim = 4, 2
    bI = np.array([0, 2])
    y = np.array([[1.0, 0.0], [0.0, 1.0]])
    for fn in (dirichlet_penalty, oe.dirichlet_penalty):
        Q, b = fn(bI, y, nv, 10.0)
        assert Q.shape == (nv * dim, nv * dim) and b.shape == (nv * dim, 1)
        x = np.zeros((nv * dim, 1))
        for k, vi in enumerate(bI):
            x[vi * dim:(vi + 1) * dim, 0] = y[k]
        assert np.allclose(Q @ x + b, 0.0, atol=1e-12)
        assert abs(oe.quadratic_energy(x, Q, b) + 0.5 * 10.0 * (y * y).sum()) < 1e-12    # minus the dropped constant


def _isolated(call, timeout=240):
    """Runs ``test_zz_quadratic.<call>`` in a child process with a time limit: inputs of a size no GPU run has seen yet
    (one or two elements) must not be able to leave a stuck kernel in the context the rest of the suite uses."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = "import sys; sys.path.insert(0, %r); import test_zz_quadratic as t; t.%s" % (here, call)
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(here), capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def _laplacian_on_one_element(mesh):
    import simkit_b200 as sk
    if mesh == "triangle":
        X, T = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]])
    else:
        X, T = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), np.array([[0, 1, 2, 3]])
    n, dim = X.shape
    L = sk.dirichlet_laplacian(X, T, mu=1.0, vector=False)
    assert L.shape == (n, n) and sps.isspmatrix_csc(L)
    D = L.toarray()
    assert np.allclose(D, D.T, atol=1e-12) and np.linalg.eigvalsh(D).min() >= -1e-10
    assert np.allclose(L @ np.ones(n), 0.0, atol=1e-10)
    assert rel(D, oe.dirichlet_laplacian(X, T, 1.0).toarray()) < 1e-12
    Lv = sk.dirichlet_laplacian(X, T, mu=1.0, vector=True)
    Dv = Lv.toarray()
    assert Lv.shape == (n * dim, n * dim) and np.allclose(Dv, Dv.T, atol=1e-12) and np.linalg.eigvalsh(Dv).min() >= -1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("mesh", ["triangle", "tet"])
def test_gpu_laplacian_on_one_element(mesh):
    """tests/test_dirichlet_laplacian.py:27-60 of the reference on its one-element meshes: csc, symmetric, positive
    semi-definite, constants in the null space; the vector form has the same properties."""
    _isolated("_laplacian_on_one_element(%r)" % mesh)


@pytest.mark.parametrize("impl", [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="simkit_b200", marks=pytest.mark.gpu)])
def test_quadratic_minimiser_and_central_differences(impl):
    """tests/test_quadratic.py of the reference: for a positive definite dense Q the energy grows away from
    x* = -Q^-1 b, the gradient matches central differences of the energy (1e-6) and vanishes at x*."""
    if impl == "gpu":
        import simkit_b200 as sk
        qe, qg, qh = sk.quadratic_energy, sk.quadratic_gradient, sk.quadratic_hessian
    else:
        qe, qg, qh = oe.quadratic_energy, oe.quadratic_gradient, oe.quadratic_hessian
    rng = np.random.default_rng(0)
    n = 6
    A = rng.standard_normal((n, n))
    Q = A.T @ A + n * np.eye(n)
    b = rng.standard_normal((n, 1))
    x_min = -np.linalg.solve(Q, b)
    assert qe(x_min + 0.1 * rng.standard_normal((n, 1)), Q, b) > qe(x_min, Q, b)
    assert np.abs(qg(x_min, Q, b)).max() < 1e-12
    x = rng.standard_normal((n, 1))
    g_fd = np.zeros(n)
    for k in range(n):
        e = np.zeros((n, 1))
        e[k] = 1e-6
        g_fd[k] = (qe(x + e, Q, b) - qe(x - e, Q, b)) / 2e-6
    g = qg(x, Q, b)
    assert g.shape == (n, 1) and np.allclose(g.ravel(), g_fd, atol=1e-6)
    assert qh(Q) is Q


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_quadratic(golden_dir, tag):
    import simkit_b200 as sk
    g, Q = load(golden_dir, tag)
    X, T, U, dim = g["X"], g["T"], g["U"], int(g["dim"])
    b = g["b"]
    x = U.reshape(-1, 1)
    E = sk.quadratic_energy(x, Q, b)
    assert isinstance(E, float) and abs(E - float(g["E"])) <= 1e-12 * abs(float(g["E"]))
    gr = sk.quadratic_gradient(x, Q, b)
    assert gr.shape == g["g"].shape and rel(gr, g["g"]) < 1e-12
    assert rel(sk.quadratic_gradient(x, Q.toarray(), b), g["g"]) < 1e-12      # dense Q, as the reference accepts
    assert sk.quadratic_hessian(Q) is Q
    # backward Euler with the term: device-resident step (ElasticPotential) and the host-callable path
    mu, lam, h = float(g["mu"]), float(g["lam"]), float(g["h"])
    Md = sps.kron(sps.diags(g["mass"]), sps.identity(dim)).tocsc()
    pot = sk.ElasticPotential("stable_neo_hookean", mu, lam, X=X, T=T, f_ext=g["fg"], quadratic=(Q, b))
    x_prev = X.reshape(-1, 1)
    x1, info = sk.backward_euler(x, x_prev, pot.energy, pot.gradient, pot.hessian, Md, h, max_iter=3, return_info=True,
                                 pcg_rtol=1e-13)
    assert list(info["alphas"]) == list(g["be_alphas"]) and rel(x1, g["be_x"]) < 1e-8
    x2 = sk.backward_euler(x, x_prev, lambda v: pot.energy(v), lambda v: pot.gradient(v), lambda v: pot.hessian(v), Md, h,
                           max_iter=3, pcg_rtol=1e-13)
    assert rel(x2, g["be_x"]) < 1e-8
    # a second step reuses the uploaded term
    x3 = sk.backward_euler(x, x_prev, pot.energy, pot.gradient, pot.hessian, Md, h, max_iter=3, pcg_rtol=1e-13)
    assert rel(x3, g["be_x"]) < 1e-8
    # entries outside the mesh's CSR pattern are refused
    nd = X.shape[0] * dim
    far = int(np.argmax(np.linalg.norm(X - X[0], axis=1)))
    Qbad = sps.csr_matrix(([1.0, 1.0], ([0, far * dim], [far * dim, 0])), shape=(nd, nd))
    with pytest.raises(ValueError):
        pot.plan.set_quadratic(Qbad, None)
    pot.plan.set_quadratic(None)


def _edge_case(case, dim):
    import simkit_b200 as sk
    from simkit_b200 import synthetic as syn
    from test_hostsim import _odd_meshes
    X, T, U = _odd_meshes(dim)[case]
    mu, lam = syn.heterogeneous_lame(T.shape[0])
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    e = oe.energy_x("stable_neo_hookean", U, Jo, mu, lam, volo)
    g = oe.gradient_x("stable_neo_hookean", U, Jo, mu, lam, volo)
    H = oe.hessian_x("stable_neo_hookean", U, Jo, mu, lam, volo, psd=True)
    J, vol = sk.deformation_jacobian(X, T), sk.volume(X, T)
    assert J.shape == Jo.shape and abs(J - Jo).max() <= 1e-12 * abs(Jo).max() and rel(vol, volo) < 1e-13
    assert abs(sk.stable_neo_hookean_energy_x(U, J, mu, lam, vol) - e) <= 1e-12 * abs(e)
    assert rel(sk.stable_neo_hookean_gradient_x(U, J, mu, lam, vol), g) < 1e-10
    Hs = sk.stable_neo_hookean_hessian_x(U, J, mu, lam, vol)
    assert Hs.shape == H.shape and abs(Hs - H).max() <= 1e-10 * abs(H).max()
    # self-contained tier and a plain scipy J (plan recovered from the operator)
    assert abs(sk.stable_neo_hookean_hessian(X, T, mu, lam, U) - H).max() <= 1e-10 * abs(H).max()
    assert abs(sk.stable_neo_hookean_hessian_x(U, sps.csc_matrix(Jo), mu, lam, volo) - H).max() <= 1e-10 * abs(H).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("case", ["unreferenced_end", "unreferenced_middle", "one_element", "two_disjoint", "listed_twice",
                                  "shuffled"])
def test_gpu_edge_case_meshes(case, dim):
    """Edge-case inputs through the drop-in API (the host replay of the same cases is tests/test_hostsim.py): vertices
    no element references, one element, disjoint elements, an element listed twice, shuffled elements."""
    _isolated("_edge_case(%r, %d)" % (case, dim))
