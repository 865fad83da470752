"""GPU parity at the sizes BASELINE.json's configs name: C3 (1 M-tet beam, Newton with pins, PCG
and line search) and C4 (4.09 M tets, reduced Hessian with r = 200) against the oracle, plus the plan's internal element
order on a shuffled mesh through every per-element input of the boundary.

The oracle runs where it finishes in about a minute: the whole C3 mesh once (assembly ~40 s, CG ~10 s on the host),
closed element sub-blocks elsewhere."""
import numpy as np
import pytest
import scipy.sparse as sps

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


def _closed_layers(cells, nl):
    """Elements of the first ``nl`` cell layers along the slowest axis (a closed sub-mesh of the generator's order) and
    the number of vertices whose rows those elements complete (the first ``nl`` vertex planes)."""
    per = 6 if len(cells) == 3 else 2
    tsub = per * nl * int(np.prod(cells[1:]))
    nfull = nl * int(np.prod([c + 1 for c in cells[1:]]))
    return tsub, nfull


# ------------------------------------------------------------------------------------------ shuffled mesh, internal order
@pytest.mark.parametrize("dim", [2, 3])
def test_shuffled_mesh_through_every_per_element_input(dim):
    """Elements AND vertices listed in random order, heterogeneous mu / lam / vol arrays, the `_u` tier's per-element
    ``Jx_bar``: the plan reorders its elements internally (``MeshPlan.element_order`` is a non-trivial permutation) and
    every result still matches the oracle in the caller's numbering."""
    cells = (12, 10, 9) if dim == 3 else (40, 31)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(21)
    vp = rng.permutation(X.shape[0])
    inv = np.empty_like(vp)
    inv[vp] = np.arange(vp.size)
    X, T = X[vp], inv[T][rng.permutation(T.shape[0])]
    n, t = X.shape[0], T.shape[0]
    U = X + 0.1 * syn.cell_size(cells, tuple(1.0 for _ in cells)) * rng.standard_normal(X.shape)
    mu, lam = syn.heterogeneous_lame(t)
    plan = sk.MeshPlan(X=X, T=T)
    order = plan.element_order()
    assert np.array_equal(np.sort(order), np.arange(t)) and not np.array_equal(order, np.arange(t))
    indptr, indices, bptr, bcol = oe.structural_pattern(T, n, dim)
    ip, ix = plan.csr_pattern()
    assert np.array_equal(ip, indptr) and np.array_equal(ix, indices)
    assert np.array_equal(plan.slot_map(), oe.slot_map(T, indptr, indices, dim))
    assert rel(plan.element_D(), oe.element_D(X, T)) < 1e-13
    assert rel(plan.volume(), oe.volume(X, T)) < 1e-14
    rho = 1e3 * (1.0 + rng.random(t))
    mo = np.zeros(n)
    np.add.at(mo, T.ravel(), np.repeat(oe.volume(X, T).ravel() * rho / (dim + 1), dim + 1))
    assert rel(plan.vertex_masses(rho), mo) < 1e-13
    Jo = oe.deformation_jacobian(X, T)
    vol = oe.volume(X, T) * (1.0 + 0.3 * rng.random((t, 1)))           # a per-element weight that is NOT the plan's own
    J = sk.deformation_jacobian(X, T)
    m = "stable_neo_hookean"
    E = sk.stable_neo_hookean_energy_x(U, J, mu, lam, vol)
    Eo = oe.energy_x(m, U, Jo, mu, lam, vol)
    assert abs(E - Eo) <= 1e-12 * abs(Eo)
    assert rel(sk.stable_neo_hookean_gradient_x(U, J, mu, lam, vol), oe.gradient_x(m, U, Jo, mu, lam, vol)) < VAL_TOL
    Q = sk.stable_neo_hookean_hessian_x(U, J, mu, lam, vol)
    Qo = oe.canonical_csr(oe.hessian_x(m, U, Jo, mu, lam, vol))
    assert abs(Q - Qo).max() / abs(Qo).max() < VAL_TOL
    # `_u` tier: displacement + per-element offset Jx_bar, listed in the caller's element order
    xb = X + 0.05 * rng.standard_normal(X.shape)
    Jxb = Jo @ xb.reshape(-1, 1)
    u = (U - xb)
    gu = sk.stable_neo_hookean_gradient_u(u, J, Jxb, mu, lam, vol)
    assert rel(gu, oe.gradient_x(m, u, Jo, mu, lam, vol, Jx_bar=Jxb)) < VAL_TOL
    Qu = sk.stable_neo_hookean_hessian_u(u, J, Jxb, mu, lam, vol)
    Quo = oe.canonical_csr(oe.hessian_x(m, u, Jo, mu, lam, vol, Jx_bar=Jxb))
    assert abs(Qu - Quo).max() / abs(Quo).max() < VAL_TOL
    # SKB_ELEMENT_ORDER=input keeps the caller's order: same values to rounding (another summation tree)
    import os
    os.environ["SKB_ELEMENT_ORDER"] = "input"
    try:
        plan_in = sk.MeshPlan(X=X, T=T)
    finally:
        del os.environ["SKB_ELEMENT_ORDER"]
    assert np.array_equal(plan_in.element_order(), np.arange(t))
    g1, v1 = plan.gradient_hessian(m, U, mu, lam, vol, 1)
    g2, v2 = plan_in.gradient_hessian(m, U, mu, lam, vol, 1)
    assert rel(v1, v2) < 1e-13 and rel(g1, g2) < 1e-12
    assert plan.n_block_partials < plan_in.n_block_partials


# ------------------------------------------------------------------------------------------ C3: 1 M-tet beam
def _c3():
    cfg = syn.CONFIGS["C3"]
    X, T = syn.make_mesh("C3")
    U = syn.jittered_state(X, cfg["cells"], cfg["extent"], sigma=0.1)
    return cfg, X, T, U


@pytest.mark.parametrize("material", ["arap", "stable_neo_hookean"])
def test_c3_beam_gradient_hessian_subblock(material):
    """BASELINE config 3 mesh (32 x 32 x 163 cells, 1,001,472 tets): gradient rows and Hessian block rows of the first
    two vertex planes against the oracle on the closed sub-mesh of the first two cell layers."""
    cfg, X, T, U = _c3()
    mu, lam = syn.lame()
    plan = sk.MeshPlan(X=X, T=T)
    assert plan.t == 1001472
    g, vals = plan.gradient_hessian(material, U, mu, lam, None, 1)
    Q = plan.csr_matrix(vals)
    tsub, nfull = _closed_layers(cfg["cells"], 2)
    Ts = T[:tsub]
    nsub = int(Ts.max()) + 1
    Jo, volo = oe.deformation_jacobian(X[:nsub], Ts), oe.volume(X[:nsub], Ts)
    go = oe.gradient_x(material, U[:nsub], Jo, mu, lam, volo)
    assert rel(g[: nfull * 3], go[: nfull * 3]) < VAL_TOL
    Qo = oe.canonical_csr(oe.hessian_x(material, U[:nsub], Jo, mu, lam, volo))
    d = Q[: nfull * 3][:, : nsub * 3] - Qo[: nfull * 3]
    assert abs(d).max() / abs(Qo).max() < VAL_TOL


def test_c3_beam_newton_iterate_against_oracle():
    """One backward-Euler Newton iteration of the C3 beam (stable neo-Hookean, gravity, face x = 0 pinned with a
    ``dirichlet_penalty``-style stiffness 1e8, line search) on the device against the oracle's Newton loop with scipy
    CG + the same 3 x 3 block-Jacobi in place of SuperLU (spsolve is infeasible at this size).
    Newton iterate within 1e-8 relative, same line-search step."""
    cfg, X, T, U = _c3()
    dim = 3
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    m = "stable_neo_hookean"
    J = sk.deformation_jacobian(X, T)
    vol = sk.volume(X, T)
    Jo, volo = oe.deformation_jacobian(X, T), oe.volume(X, T)
    M = sps.kron(oe.massmatrix(X, T, rho), sps.identity(dim)).tocsc()
    fg = oe.gravity_force(X, T, -9.8, rho).reshape(-1, 1)
    pin_k = np.zeros((X.shape[0], dim))
    pin_k[X[:, 0] == 0.0] = 1e8
    pin_k = pin_k.reshape(-1, 1)
    pin_t = X.reshape(-1, 1).copy()
    x_curr = U.reshape(-1, 1)

    def Eo(x):
        d = x - pin_t
        return oe.energy_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) - float((fg.T @ x).item()) + 0.5 * float((pin_k * d * d).sum())

    def Go(x):
        return oe.gradient_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) - fg + pin_k * (x - pin_t)

    def Ho(x):
        return oe.hessian_x(m, x.reshape(-1, dim), Jo, mu, lam, volo) + sps.diags(pin_k.ravel())

    its = []

    def solver(Hm, rhs):
        x, it = oe.block_jacobi_cg(Hm, rhs, dim, rtol=1e-12)
        its.append(it)
        return x

    xo, io = oe.backward_euler(x_curr, x_curr, Eo, Go, Ho, M, h, max_iter=1, return_info=True, linear_solver=solver)
    pot = sk.ElasticPotential(m, mu, lam, vol, J=J, dim=dim, f_ext=fg, pin_k=pin_k, pin_target=pin_t)
    x1, i1 = sk.backward_euler(x_curr, x_curr, pot.energy, pot.gradient, pot.hessian, M, h, max_iter=1, return_info=True,
                               pcg_rtol=1e-12)
    assert list(i1["alphas"]) == list(io["alphas"])
    assert rel(x1 - x_curr, xo - x_curr) < 1e-6        # the step itself, relative to the step
    assert rel(x1, xo) < ITER_TOL
    assert abs(i1["pcg_iters"] - its[0]) <= max(3, its[0] // 20)   # same preconditioner: same iteration count (+-5 %)


# ------------------------------------------------------------------------------------------ C4: reduced Hessian, r = 200
def _cos_modes(X, r, seed=2):
    """Smooth modes like ``synthetic.smooth_modes`` without the QR (170 GFLOP on the host at C4); scaled to unit columns."""
    rng = np.random.default_rng(seed)
    n, dim = X.shape
    Xn = (X - X.min(0)) / (X.max(0) - X.min(0))
    B = np.empty((n * dim, r))
    for j in range(r):
        k = rng.integers(0, 4, size=dim)
        phase = rng.uniform(0, np.pi, size=dim)
        f = np.prod(np.cos(np.pi * k[None, :] * Xn + phase[None, :]), axis=1)
        w = rng.standard_normal(dim)
        B[:, j] = (f[:, None] * w[None, :]).reshape(-1) / np.sqrt(n)
    return B


def test_c4_reduced_hessian_r200_subblock():
    """BASELINE config 4 (88^3 cells, 4,088,832 tets, r = 200): the reduced energy / gradient / Hessian of the element
    sub-block made of the first three cell layers -- selected on the device by a per-element weight that is zero
    elsewhere, with the floor applied before the weight (the `_z` tier's order, elastic.py:663-664) -- against
    ``B^T Q B`` of the oracle on that closed sub-mesh."""
    cfg = syn.CONFIGS["C4"]
    X, T = syn.make_mesh("C4")
    r = 200
    mu, lam = syn.lame()
    B = _cos_modes(X, r)
    z = 0.02 * np.random.default_rng(3).standard_normal(r)
    plan = sk.MeshPlan(X=X, T=T)
    assert plan.t == 4088832
    tsub, _ = _closed_layers(cfg["cells"], 3)
    vol = plan.volume()
    w = np.zeros_like(vol)
    w[:tsub] = vol[:tsub]
    plan.set_materials(mu, lam, w)
    E, gr, Hr = plan.reduced("stable_neo_hookean", B, z, x0=X.reshape(-1), psd_mode=2)
    Ts = T[:tsub]
    nsub = int(Ts.max()) + 1
    Bs = B[: nsub * 3]
    Jo, volo = oe.deformation_jacobian(X[:nsub], Ts), oe.volume(X[:nsub], Ts)
    x = (Bs @ z).reshape(-1, 3) + X[:nsub]
    Qo = oe.hessian_x("stable_neo_hookean", x, Jo, mu, lam, volo, psd_before_vol=True)
    Ho = Bs.T @ (Qo @ Bs)
    assert rel(Hr, Ho) < VAL_TOL
    assert rel(gr, Bs.T @ oe.gradient_x("stable_neo_hookean", x, Jo, mu, lam, volo)) < VAL_TOL
    Eo = oe.energy_x("stable_neo_hookean", x, Jo, mu, lam, volo)
    assert abs(E - Eo) <= 1e-11 * abs(Eo)
    assert np.abs(Hr - Hr.T).max() <= 1e-12 * np.abs(Hr).max()
    # the whole mesh: symmetric, positive semi-definite, and at least the sub-block in the Loewner order
    plan.set_materials(mu, lam, vol)
    E2, g2, H2 = plan.reduced("stable_neo_hookean", B, z, x0=X.reshape(-1), psd_mode=2)
    assert np.abs(H2 - H2.T).max() <= 1e-12 * np.abs(H2).max()
    ev = np.linalg.eigvalsh(H2 - Hr)
    assert ev.min() >= -1e-9 * np.abs(ev).max()
