"""The small full-space MFEM problem of the reference's tests/test_mfem_blocks.py:60-155 (TEST INFRASTRUCTURE).

``mfem_problem(sk_mod, ...)`` builds the block closures from whatever module ``sk_mod`` provides the simkit
surface: the reference itself (oracle/make_golden.py, to freeze the golden fixture), the oracle namespace
(``oracle.elasticity.MfemSurface``, CPU test) or ``simkit_b200`` (GPU test).
"""

import numpy as np


def mfem_problem(sk_mod, X, T, rho_aug, material="macklin-mueller-neo-hookean"):
    """The small full-space MFEM problem of the reference's tests/test_mfem_blocks.py:60-155, built from the
    functions of ``sk_mod`` (the reference here; tests/test_gpu_mfem.py rebuilds it from simkit_b200)."""
    import scipy as sp
    n, dim = X.shape
    q = X.reshape(-1, 1)
    nz = n * dim
    t = T.shape[0]
    vol = sk_mod.volume(X, T).reshape(-1, 1)
    mu_l, lam_l = sk_mod.ympr_to_lame(1e4, 0.45)
    mu = np.full((t, 1), mu_l)
    lam = np.full((t, 1), lam_l)
    Mv = sp.sparse.kron(sk_mod.massmatrix(X, T, 1e3), sp.sparse.identity(dim)).tocsc()
    J = sk_mod.deformation_jacobian(X, T)
    GJB = sp.sparse.csc_matrix(J)
    GJq = J @ q
    C, Ci = sk_mod.symmetric_stretch_map(t, dim)
    k = dim * (dim + 1) // 2
    wv = np.array([[1.0] * dim + [2.0] * (k - dim)]).T
    w = np.kron(vol, wv)
    W = sp.sparse.diags(w.flatten())
    Wi = sp.sparse.diags(1.0 / w.flatten())
    na = Ci.shape[0]
    h = 1e-2
    Qm = (1e3 * sp.sparse.identity(nz)).tocsc()
    b = np.ones((nz, 1)) * 0.1
    rng = np.random.default_rng(0)
    z_curr = 0.01 * rng.standard_normal((nz, 1))
    z_prev = 0.01 * rng.standard_normal((nz, 1))
    y = 2.0 * z_curr - z_prev                      # backward-Euler target (energies/kinetic.py:107-119)
    E = sk_mod.energies

    def split(p):
        return p[:nz], p[nz:nz + na], p[nz + na:]

    def energy(p):
        u, a, ll = split(p)
        A = a.reshape(-1, k)
        F = np.asarray(GJB @ u + GJq).reshape(-1, dim, dim)
        c = Ci @ sk_mod.stretch(F) - a
        wc = w * c
        el = E.elastic_energy_S(A, mu, lam, vol, material)
        d = u - y
        kin = 0.5 / h ** 2 * float((d.T @ (Mv @ d)).item())
        quad = 0.5 * float((u.T @ (Qm @ u)).item()) + float((b.T @ u).item())
        return el + kin + quad + float((ll.T @ wc).item()) + 0.5 * rho_aug * float((c.T @ wc).item())

    def grad_blocks(p):
        u, a, ll = split(p)
        A = a.reshape(-1, k)
        F = np.asarray(GJB @ u + GJq).reshape(-1, dim, dim)
        c = Ci @ sk_mod.stretch(F) - a
        wc = w * c
        G_u = sk_mod.stretch_gradient_dz(u, GJB, Ci=Ci, dim=dim, GJq=GJq) @ W
        f_u = Mv @ (u - y) / h ** 2 + Qm @ u + b + rho_aug * (G_u @ c)
        f_z = E.elastic_gradient_S(A, mu, lam, vol, material).reshape(-1, 1) - rho_aug * wc
        return [f_u, f_z, wc]

    def hess_blocks(p):
        u, a, ll = split(p)
        A = a.reshape(-1, k)
        G_u = sk_mod.stretch_gradient_dz(u, GJB, Ci=Ci, dim=dim, GJq=GJq) @ W
        H_u = Mv / h ** 2 + Qm + rho_aug * (G_u @ Wi @ G_u.T)
        H_z = sp.sparse.block_diag([hh for hh in E.elastic_hessian_S(A, mu, lam, vol, material)]) + rho_aug * W
        G_z = -W
        G_zi = sp.sparse.diags(1.0 / G_z.diagonal())
        return [H_u, H_z, G_u, G_z, G_zi]

    rng = np.random.default_rng(1)
    u0 = 0.03 * rng.standard_normal((nz, 1))
    a0 = np.tile(np.array([[1.0] * dim + [0.0] * (k - dim)]), (t, 1)).reshape(-1, 1) + 0.03 * rng.standard_normal((na, 1))
    ll0 = 0.05 * rng.standard_normal((na, 1))
    return dict(energy=energy, grad_blocks=grad_blocks, hess_blocks=hess_blocks, p0=np.vstack([u0, a0, ll0]), nz=nz, na=na,
                GJB=GJB, GJq=GJq, Ci=Ci)
