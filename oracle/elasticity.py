"""numpy/scipy oracle for the hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function restates, in its own words, what the cited reference code
computes (paths relative to ``/root/reference/simkit``).  The algorithmic
*steps* are kept the same as the reference's (SpMV ``J@x`` -> batched element
formulas -> LAPACK ``eigh`` per block -> ``block_diag`` -> two SpGEMMs), so that
timing this module is a fair stand-in for the reference CPU path; the element
formulas are written from the constitutive models with ``einsum`` rather than
the reference's generated scalar code.
"""

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import scipy.linalg

MATERIALS = ("stable_neo_hookean", "neo_hookean", "arap", "stvk", "linear_elasticity",
             "fcr", "macklin_mueller_neo_hookean")

PSD_FLOOR = 1e-6  # psd_project.py:37


# --------------------------------------------------------------------------- #
# mesh -> operators                                                           #
# --------------------------------------------------------------------------- #
def _ref_grad(dim):
    # reference-element shape function gradients, deformation_jacobian.py:43-54
    return np.vstack([-np.ones((1, dim)), np.eye(dim)])


def element_D(X, T):
    """Per-element ``D (t, dim, dim+1)`` with ``F_ij = sum_a D[j,a] x[a,i]``.

    deformation_jacobian.py:58-61: ``D = (H (X_e^T H)^-1)^T``.
    """
    X = np.asarray(X, dtype=np.float64)
    dim = X.shape[1]
    Hs = _ref_grad(dim)
    Xe = X[T]                                   # (t, dim+1, dim)
    edge = np.einsum("tad,ak->tdk", Xe, Hs)     # X_e^T H   (t, dim, dim)
    inv = np.linalg.inv(edge)
    return np.einsum("ak,tkj->tja", Hs, inv)    # (H inv)^T  (t, dim, dim+1)


def deformation_jacobian(X, T):
    """Sparse ``J`` with ``vec_rowmajor(F_e) = (J x)[e*d*d:(e+1)*d*d]``.

    deformation_jacobian.py:9-87.  Row ``e*d*d + i*d + j``, column ``v*d + i``,
    value ``D[e, j, a]`` for ``v = T[e, a]``.  The reference forms it as a
    SpGEMM, which prunes exact zeros; mirrored with ``eliminate_zeros``.
    """
    X = np.asarray(X, dtype=np.float64)
    T = np.asarray(T)
    dim = X.shape[1]
    dt = T.shape[1]
    t = T.shape[0]
    n = X.shape[0]
    D = element_D(X, T)
    e = np.arange(t)[:, None, None, None]
    i = np.arange(dim)[None, :, None, None]
    j = np.arange(dim)[None, None, :, None]
    rows = np.broadcast_to(e * dim * dim + i * dim + j, (t, dim, dim, dt))
    cols = np.broadcast_to(T[:, None, None, :] * dim + i, (t, dim, dim, dt))
    vals = np.broadcast_to(D[:, None, :, :], (t, dim, dim, dt))
    J = sps.csc_matrix(
        (vals.ravel(), (rows.ravel(), cols.ravel())), shape=(t * dim * dim, n * dim)
    )
    J.sum_duplicates()
    J.eliminate_zeros()
    return J


def volume(X, T):
    """(t,1) signed tet volume / unsigned triangle area.

    volume.py:14-41, tetrahedron_volumes.py:26-27, triangle_areas.py:60-79.
    """
    X = np.asarray(X, dtype=np.float64)
    if T.shape[1] == 4:
        e = X[T[:, 1:]] - X[T[:, [0]]]
        return (np.linalg.det(e) / 6.0).reshape(-1, 1)
    if T.shape[1] == 3:
        if X.shape[1] == 2:
            X = np.hstack([X, np.zeros((X.shape[0], 1))])
        a = X[T[:, 1]] - X[T[:, 0]]
        b = X[T[:, 2]] - X[T[:, 0]]
        return (0.5 * np.linalg.norm(np.cross(a, b), axis=1)).reshape(-1, 1)
    raise ValueError("volume: only triangles and tets")


def vertex_masses(X, T, rho=1.0):
    """Lumped vertex masses: element mass split equally over corners (massmatrix.py:41-49)."""
    m = volume(X, T) * rho
    vv = np.zeros(X.shape[0])
    np.add.at(vv, T.ravel(), np.repeat(m.ravel(), T.shape[1]))
    return vv / T.shape[1]


def massmatrix(X, T, rho=1.0):
    return sps.diags(vertex_masses(X, T, rho))


def gravity_force(X, T, a=-9.8, rho=1.0):
    """``M [0, a, 0]`` per vertex: always axis 1 (gravity_force.py:36-41)."""
    g = np.zeros(X.shape)
    g[:, 1] = a
    return massmatrix(X, T, rho) @ g


# --------------------------------------------------------------------------- #
# small batched matrix helpers                                                #
# --------------------------------------------------------------------------- #
def psd_project(H, method="proj"):
    """psd_project.py:12-47: eigh; floor eigenvalues at 1e-6 ('proj') or abs; rebuild."""
    H = np.asarray(H)
    if H.ndim == 2:
        H = H[None]
    s, U = np.linalg.eigh(H)
    if method == "abs":
        s = np.abs(s)
    elif method == "proj":
        s = np.where(s < PSD_FLOOR, PSD_FLOOR, s)
    with np.errstate(all="ignore"):
        return np.einsum("tik,tk,tjk->tij", U, s, U)


def svd_rv(F):
    """Rotation-variant SVD (svd_rv.py:8-53): ``F = U S V^T`` with the reflection
    pushed onto the last singular value when exactly one of U, V is improper."""
    F = np.asarray(F, dtype=np.float64)
    if F.ndim == 2:
        F = F[None]
    d = F.shape[-1]
    U, s, Vt = np.linalg.svd(F)
    V = np.swapaxes(Vt, -1, -2)
    detU = np.linalg.det(U)
    detV = np.linalg.det(V)
    flipU = (detU < 0) & (detV > 0)
    flipV = (detV < 0) & (detU > 0)
    sgn = detU * detV                      # det(U V^T)
    U = U.copy()
    V = V.copy()
    s = s.copy()
    U[flipU, :, d - 1] *= sgn[flipU, None]
    V[flipV, :, d - 1] *= sgn[flipV, None]
    s[:, d - 1] *= sgn
    S = np.zeros_like(F)
    idx = np.arange(d)
    S[:, idx, idx] = s
    return U, S, V


def polar_svd(F, flip=True):
    """polar_svd.py:59-89 (only ``flip=True`` is live in the reference)."""
    if not flip:
        raise NameError("polar_svd(flip=False) is broken in the reference (polar_svd.py:80-84)")
    U, S, V = svd_rv(F)
    Vt = np.swapaxes(V, -1, -2)
    return U @ Vt, V @ S @ Vt


def rotation_gradient_F(F):
    """dR/dF from twist modes (rotation_gradient.py:12-75), with its clamps."""
    dim = F.shape[-1]
    n = F.shape[0]
    U, S, V = svd_rv(F)
    Vt = np.swapaxes(V, -1, -2)
    K = np.zeros((n, dim * dim, dim * dim))
    if dim == 2:
        pairs = [((0, 1), 1e-12)]
    else:
        pairs = [((0, 1), 1e-8), ((1, 2), 1e-8), ((0, 2), 1e-8)]
    for (p, q), clamp in pairs:
        E = np.zeros((dim, dim))
        E[p, q] = -1.0
        E[q, p] = 1.0
        tw = (U @ E @ Vt / np.sqrt(2.0)).reshape(n, dim * dim)
        den = np.maximum(S[:, p, p] + S[:, q, q], clamp)
        K += (2.0 / den)[:, None, None] * tw[:, :, None] * tw[:, None, :]
    return K


def stretch_gradient_dF(F):
    """dS/dF of S = R^T F, (t,d,d,d,d) indexed [m,n,i,j] = dS_ij/dF_mn (stretch_gradient.py:28-54):
    sum_k dR_ki/dF_mn F_kj + R_mi delta_nj, with dR/dF the symmetric matrix of rotation_gradient_F."""
    dim = F.shape[-1]
    R, _ = polar_svd(F)
    K = rotation_gradient_F(F).reshape(-1, dim, dim, dim, dim)
    eye = np.eye(dim)
    return np.einsum("tmnki,tkj->tmnij", K, F) + np.einsum("tmi,nj->tmnij", R, eye)


def symmetric_stretch_map(t, dim):
    """symmetric_stretch_map.py:46-67: diagonal entries first, then the upper triangle row by row; the inverse map
    averages the two copies of an off-diagonal."""
    k = dim * (dim + 1) // 2
    col = {}
    c = 0
    for i in range(dim):
        col[(i, i)] = c
        c += 1
    for i in range(dim):
        for j in range(i + 1, dim):
            col[(i, j)] = col[(j, i)] = c
            c += 1
    S = np.zeros((dim * dim, k))
    Si = np.zeros((k, dim * dim))
    for i in range(dim):
        for j in range(dim):
            S[i * dim + j, col[(i, j)]] = 1.0
            Si[col[(i, j)], i * dim + j] = 1.0 if i == j else 0.5
    return sps.kron(sps.identity(t), sps.csc_matrix(S)), sps.kron(sps.identity(t), sps.csc_matrix(Si))


def dirichlet_laplacian(X, T, mu=1, vector=False):
    """dirichlet_laplacian.py:16-76: H = J^T diag(vol * mu) J; vector=False averages its dim per-coordinate blocks
    into the (n, n) vertex Laplacian."""
    X = np.asarray(X, dtype=np.float64)
    t = np.asarray(T).shape[0]
    mu = np.ones((t, 1)) * mu if np.isscalar(mu) else np.asarray(mu, dtype=np.float64).reshape(-1, 1)
    assert mu.shape[0] == t
    n, dim = X.shape
    a = volume(X, T).reshape(-1, 1) * mu
    J = deformation_jacobian(X, T)
    H = (J.T @ sps.diags(np.repeat(a.ravel(), dim * dim)) @ J).tocsc()
    if vector:
        return H
    L = sps.csc_matrix((n, n))
    for i in range(dim):
        Ii = np.arange(n) * dim + i
        L = L + H[Ii, :][:, Ii]
    return sps.csc_matrix(L / dim)


def quadratic_energy(x, Q, b):
    """energies/quadratic.py:15-34: float(1/2 x^T Q x + b^T x)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 1)
    return float(np.asarray(0.5 * x.T @ (Q @ x) + b.T @ x).item())


def quadratic_gradient(x, Q, b):
    """energies/quadratic.py:37-54: Q x + b, (n, 1)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    return np.asarray(Q @ x) + np.asarray(b, dtype=np.float64).reshape(-1, 1)


def quadratic_hessian(Q):
    """energies/quadratic.py:57-70: Q itself."""
    return Q


def dirichlet_penalty(bI, y, nv, gamma):
    """dirichlet_penalty.py:65-142 (default options): Q = S Gamma S^T, b = -S Gamma y with S the (nv*d, cn*d)
    selection of the pinned vertices' dofs and Gamma = diag(gamma) (x) I_d."""
    y = np.asarray(y, dtype=np.float64)
    assert y.ndim == 2
    bI = np.asarray(bI).reshape(-1)
    cn, d = bI.shape[0], y.shape[1]
    gam = np.ones(cn) * gamma if np.isscalar(gamma) else np.asarray(gamma, dtype=np.float64).reshape(-1)
    dofs = (bI[:, None] * d + np.arange(d)[None, :]).ravel()
    w = np.repeat(gam, d)
    Q = sps.csc_matrix((w, (dofs, dofs)), (nv * d, nv * d))
    b = np.zeros((nv * d, 1))
    np.add.at(b[:, 0], dofs, -w * y.reshape(-1))
    return Q, b


def contact_springs_plane(X, k, p, n, M=None):
    """energies/contact_springs_plane.py:245-388: (energy, gradient (n*d,1), Hessian csc, contacting indices) of
    k/2 sum_{v: n.(x_v - p) < 0} m_v (n.(x_v - p))^2, m = diag(M) (identity by default)."""
    X = np.asarray(X, dtype=np.float64)
    nv, dim = X.shape
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    n = np.asarray(n, dtype=np.float64).reshape(-1)
    m = np.ones(nv) if M is None else np.asarray(sps.csr_matrix(M).diagonal())
    off = (X - p[None, :]) @ n
    under = off < 0
    km = np.where(under, k * m, 0.0)
    E = float((0.5 * km * off * off).sum())
    g = (km * off)[:, None] * n[None, :]
    H = sps.block_diag([kmv * np.outer(n, n) for kmv in km], format="csc")
    return E, g.reshape(-1, 1), H, np.where(under)[0]


def contact_springs_sphere(X, k, p, r, M=None):
    """energies/contact_springs_sphere.py:242-360: vertices with |x_v - p| < r, n_v = (x_v - p)/|x_v - p| held fixed:
    (energy, gradient (n*d,1), Hessian csc) of k/2 sum m_v (|x_v - p| - r)^2."""
    X = np.asarray(X, dtype=np.float64)
    nv, dim = X.shape
    p = np.asarray(p, dtype=np.float64).reshape(-1)
    m = np.ones(nv) if M is None else np.asarray(sps.csr_matrix(M).diagonal())
    d = X - p[None, :]
    ln = np.linalg.norm(d, axis=1)
    inside = ln < r
    with np.errstate(all="ignore"):
        n = d / ln[:, None]
    off = ln - r
    km = np.where(inside, k * m, 0.0)
    E = float((0.5 * km * off * off).sum())
    g = np.where(inside[:, None], (km * off)[:, None] * n, 0.0)
    H = sps.block_diag([kmv * np.outer(nn, nn) if kmv != 0.0 else np.zeros((dim, dim)) for kmv, nn in zip(km, n)], format="csc")
    return E, g.reshape(-1, 1), H


def stretch(F):
    """stretch.py:9-27: S of F = R S, stacked as a column."""
    _, S = polar_svd(F)
    return S.reshape(-1, 1)


def stretch_gradient_dz(z, GJB, dim, Ci=None, GJq=None):
    """stretch_gradient.py:57-131: J^T blockdiag(dS/dF) [Ci^T]."""
    f = GJB @ np.asarray(z).reshape(-1, 1)
    if GJq is not None:
        f = f + GJq
    F = np.asarray(f).reshape(-1, dim, dim)
    blocks = stretch_gradient_dF(F).reshape(-1, dim * dim, dim * dim)
    out = GJB.T @ sps.block_diag(blocks)
    return out if Ci is None else out @ Ci.T


def _compact_embedding(dim):
    S, _ = symmetric_stretch_map(1, dim)
    return np.asarray(S.todense())


def elastic_S(kind, S, mu, lam, vol, material):
    """Stretch tier of the dispatcher for Macklin-Mueller NH (energies/elastic.py:327-357, 563-593, 816-846;
    macklin_mueller_neo_hookean.py:395-478): the F tier at the symmetric stretch, mapped to the compact components
    by the embedding C0; the Hessian is PSD-projected before the vol weighting."""
    assert material == "macklin-mueller-neo-hookean"
    m = "macklin_mueller_neo_hookean"
    S = np.asarray(S)
    t, k = S.shape
    dim = 2 if k == 3 else 3
    C0 = _compact_embedding(dim)
    Sf = (S @ C0.T).reshape(t, dim, dim)
    w = np.asarray(vol, dtype=np.float64).reshape(-1, 1)
    if kind == "energy":
        return float((w * energy_element_F(m, Sf, mu, lam)).sum())
    if kind == "gradient":
        return (gradient_element_F(m, Sf, mu, lam).reshape(t, dim * dim) @ C0) * w
    H = np.einsum("ji,tjk,kl->til", C0, hessian_element_F(m, Sf, mu, lam), C0)
    return psd_project(H) * w.reshape(-1, 1, 1)


class _MfemEnergies:
    @staticmethod
    def elastic_energy_S(S, mu, lam, vol, material):
        return elastic_S("energy", S, mu, lam, vol, material)

    @staticmethod
    def elastic_gradient_S(S, mu, lam, vol, material):
        return elastic_S("gradient", S, mu, lam, vol, material)

    @staticmethod
    def elastic_hessian_S(S, mu, lam, vol, material):
        return elastic_S("hessian", S, mu, lam, vol, material)


class MfemSurface:
    """The part of the simkit surface oracle/mfem_problem.py needs, backed by this oracle."""
    energies = _MfemEnergies
    volume = staticmethod(volume)
    massmatrix = staticmethod(massmatrix)
    deformation_jacobian = staticmethod(deformation_jacobian)
    symmetric_stretch_map = staticmethod(symmetric_stretch_map)
    stretch = staticmethod(stretch)
    stretch_gradient_dz = staticmethod(stretch_gradient_dz)

    @staticmethod
    def ympr_to_lame(ym, pr):
        return ym / (2.0 * (1.0 + pr)), ym * pr / ((1.0 + pr) * (1.0 - 2.0 * pr))


def sqp_mfem(p0, energy_func, hess_blocks_func, grad_blocks_func, tolerance=1e-4, max_iter=100, do_line_search=True):
    """solvers/sqpmfem.py:56-89 with direct solves."""
    p = p0.copy()
    for _ in range(max_iter):
        H_u, H_z, G_u, G_z, G_zi = hess_blocks_func(p)
        f_u, f_z, f_mu = grad_blocks_func(p)
        Q = H_u + G_u @ G_zi @ H_z @ G_zi @ G_u.T
        g_u = -f_u + G_u @ G_zi @ (f_z - H_z @ G_zi @ f_mu)
        du = spla.spsolve(sps.csc_matrix(Q), g_u) if sps.issparse(Q) else scipy.linalg.solve(Q, g_u)
        du = np.asarray(du).reshape(-1, 1)
        dz = G_zi @ (-(f_mu + G_u.T @ du))
        mu = -G_zi @ (f_z + H_z @ dz)
        g = np.vstack([f_u + G_u @ mu, f_z + G_z @ mu])
        dp = np.vstack([du, dz])
        if do_line_search:
            alpha, _, _ = backtracking_line_search(lambda q: energy_func(np.vstack([q, mu])), p[:-mu.shape[0]], g, dp)
        else:
            alpha = 1.0
        p[:-mu.shape[0]] += alpha * dp
        p[-mu.shape[0]:] = mu
        if float((np.asarray(g_u).T @ du).item()) < tolerance:
            break
    return p


# --------------------------------------------------------------------------- #
# element tier                                                                #
# --------------------------------------------------------------------------- #
def _levi(dim):
    if dim == 2:
        return np.array([[0.0, 1.0], [-1.0, 0.0]])
    e = np.zeros((3, 3, 3))
    e[0, 1, 2] = e[1, 2, 0] = e[2, 0, 1] = 1.0
    e[0, 2, 1] = e[2, 1, 0] = e[1, 0, 2] = -1.0
    return e


def _det_cof_dcof(F):
    """det F, cofactor c = dJ/dF (t,d,d) and dc/dF (t,d,d,d,d)."""
    dim = F.shape[-1]
    t = F.shape[0]
    eps = _levi(dim)
    if dim == 2:
        c = np.einsum("ik,jl,tkl->tij", eps, eps, F)
        dc = np.broadcast_to(np.einsum("ik,jl->ijkl", eps, eps), (t, 2, 2, 2, 2))
        Jd = F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
    else:
        dc = np.einsum("ikm,jln,tmn->tijkl", eps, eps, F)
        c = 0.5 * np.einsum("tijkl,tkl->tij", dc, F)
        Jd = np.einsum("tj,tj->t", F[:, 0, :], c[:, 0, :])
    return Jd, c, dc


def _bc(a, nd):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return a.reshape((-1,) + (1,) * nd)


def energy_element_F(material, F, mu, lam=None):
    """psi (t,1).  sNH stable_neo_hookean.py:65-129; NH neo_hookean.py:64-97;
    ARAP arap.py:71-91; StVK stvk.py:61-93; LE linear_elasticity.py:43-70;
    FCR fcr.py:41-62; Macklin-Mueller NH macklin_mueller_neo_hookean.py:66-122."""
    F = np.asarray(F, dtype=np.float64)
    dim = F.shape[-1]
    F = F.reshape(-1, dim, dim)
    mu = _bc(mu, 0)
    lam = _bc(lam, 0) if lam is not None else None
    I_C = np.einsum("tij,tij->t", F, F)
    if material == "stable_neo_hookean":
        Jd = np.linalg.det(F) if dim == 3 else F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
        alpha = 1.0 + dim * mu / ((dim + 1) * lam)
        psi = 0.5 * mu * (I_C - dim) - 0.5 * mu * np.log(I_C + 1.0) + 0.5 * lam * (Jd - alpha) ** 2
    elif material == "neo_hookean":
        with np.errstate(all="ignore"):
            lJ = np.log(np.linalg.det(F))
            psi = 0.5 * mu * (I_C - dim) - mu * lJ + 0.5 * lam * lJ ** 2
    elif material == "arap":
        R, _ = polar_svd(F)
        psi = 0.5 * mu * np.einsum("tij,tij->t", F - R, F - R)
    elif material == "stvk":
        E = 0.5 * (np.einsum("tki,tkj->tij", F, F) - np.eye(dim))
        psi = mu * np.einsum("tij,tij->t", E, E) + 0.5 * lam * np.trace(E, axis1=1, axis2=2) ** 2
    elif material == "linear_elasticity":
        eps = 0.5 * (F + np.swapaxes(F, 1, 2)) - np.eye(dim)
        psi = mu * np.einsum("tij,tij->t", eps, eps) + 0.5 * lam * np.trace(eps, axis1=1, axis2=2) ** 2
    elif material == "fcr":
        # fcr.py:41-62: twice the ARAP shear term plus lam/2 (det F - 1)^2
        R, _ = polar_svd(F)
        psi = mu * np.einsum("tij,tij->t", F - R, F - R) + 0.5 * lam * (np.linalg.det(F) - 1.0) ** 2
    elif material == "macklin_mueller_neo_hookean":
        # macklin_mueller_neo_hookean.py:66-122: mu (1-J) + lam/2 (1-J)^2 + mu/2 (I_C - d)
        Jd, _, _ = _det_cof_dcof(F)
        psi = mu * (1.0 - Jd) + 0.5 * lam * (1.0 - Jd) ** 2 + 0.5 * mu * (I_C - dim)
    else:
        raise ValueError("unknown material " + str(material))
    return psi.reshape(-1, 1)


def gradient_element_F(material, F, mu, lam=None):
    """PK1 P (t,d,d).  sNH :132-218; NH :100-131; ARAP :94-114; StVK :96-127; LE :73-100."""
    F = np.asarray(F, dtype=np.float64)
    dim = F.shape[-1]
    F = F.reshape(-1, dim, dim)
    mu = _bc(mu, 2)
    lam = _bc(lam, 2) if lam is not None else None
    eye = np.eye(dim)
    if material == "stable_neo_hookean":
        Jd, c, _ = _det_cof_dcof(F)
        I_C = np.einsum("tij,tij->t", F, F)[:, None, None]
        alpha = 1.0 + dim * mu / ((dim + 1) * lam)
        return mu * (1.0 - 1.0 / (I_C + 1.0)) * F + lam * (Jd[:, None, None] - alpha) * c
    if material == "neo_hookean":
        with np.errstate(all="ignore"):
            lJ = np.log(np.linalg.det(F))[:, None, None]
            FinvT = np.swapaxes(np.linalg.inv(F), 1, 2)
            return mu * F + (lam * lJ - mu) * FinvT
    if material == "arap":
        R, _ = polar_svd(F)
        return mu * (F - R)
    if material == "stvk":
        E = 0.5 * (np.einsum("tki,tkj->tij", F, F) - eye)
        S2 = 2.0 * mu * E + lam * np.trace(E, axis1=1, axis2=2)[:, None, None] * eye
        return F @ S2
    if material == "linear_elasticity":
        return mu * (F + np.swapaxes(F, 1, 2) - 2 * eye) + lam * np.trace(F - eye, axis1=1, axis2=2)[:, None, None] * eye
    if material == "fcr":
        # fcr.py:65-125: 2 mu (F - R) + lam (J - 1) cof F
        Jd, c, _ = _det_cof_dcof(F)
        R, _ = polar_svd(F)
        return 2.0 * mu * (F - R) + lam * (Jd[:, None, None] - 1.0) * c
    if material == "macklin_mueller_neo_hookean":
        # macklin_mueller_neo_hookean.py:125-185: mu F + (lam (J - 1) - mu) cof F
        Jd, c, _ = _det_cof_dcof(F)
        return mu * F + (lam * (Jd[:, None, None] - 1.0) - mu) * c
    raise ValueError("unknown material " + str(material))


def hessian_element_F(material, F, mu, lam=None):
    """d2psi/dF2 (t,d*d,d*d), row-major F layout, unweighted and unprojected.
    sNH :221-443; NH :134-178; ARAP :117-145; StVK :130-179; LE :103-136."""
    F = np.asarray(F, dtype=np.float64)
    dim = F.shape[-1]
    F = F.reshape(-1, dim, dim)
    t = F.shape[0]
    b = dim * dim
    mu4 = _bc(mu, 4)
    lam4 = _bc(lam, 4) if lam is not None else None
    eye = np.eye(dim)
    II = np.einsum("ik,jl->ijkl", eye, eye)[None]
    if material == "stable_neo_hookean":
        Jd, c, dc = _det_cof_dcof(F)
        I_C = np.einsum("tij,tij->t", F, F).reshape(-1, 1, 1, 1, 1)
        alpha = 1.0 + dim * mu4 / ((dim + 1) * lam4)
        H5 = (
            mu4 * (1.0 - 1.0 / (I_C + 1.0)) * II
            + (2.0 * mu4 / (I_C + 1.0) ** 2) * np.einsum("tij,tkl->tijkl", F, F)
            + lam4 * np.einsum("tij,tkl->tijkl", c, c)
            + lam4 * (Jd.reshape(-1, 1, 1, 1, 1) - alpha) * dc
        )
    elif material == "neo_hookean":
        with np.errstate(all="ignore"):
            lJ = np.log(np.linalg.det(F)).reshape(-1, 1, 1, 1, 1)
            G = np.swapaxes(np.linalg.inv(F), 1, 2)
            H5 = (
                mu4 * II
                + lam4 * np.einsum("tij,tkl->tijkl", G, G)
                + (mu4 - lam4 * lJ) * np.einsum("til,tkj->tijkl", G, G)
            )
    elif material == "arap":
        H = mu4.reshape(-1, 1, 1) * (np.eye(b)[None] - rotation_gradient_F(F))
        return H
    elif material == "stvk":
        E = 0.5 * (np.einsum("tki,tkj->tij", F, F) - eye)
        FFt = np.einsum("tik,tjk->tij", F, F)
        trE = np.trace(E, axis1=1, axis2=2).reshape(-1, 1, 1, 1, 1)
        H5 = (
            2.0 * mu4 * np.einsum("ik,tlj->tijkl", eye, E)
            + mu4 * np.einsum("til,tkj->tijkl", F, F)
            + mu4 * np.einsum("tik,jl->tijkl", FFt, eye)
            + lam4 * np.einsum("tij,tkl->tijkl", F, F)
            + lam4 * trE * II
        )
    elif material == "fcr":
        # fcr.py:128-302: 2 * ARAP Hessian + lam (c c^T + (J - 1) dc/dF)
        Jd, c, dc = _det_cof_dcof(F)
        H5 = lam4 * np.einsum("tij,tkl->tijkl", c, c) + lam4 * (Jd.reshape(-1, 1, 1, 1, 1) - 1.0) * dc
        Hs = 2.0 * mu4.reshape(-1, 1, 1) * (np.eye(b)[None] - rotation_gradient_F(F))
        return np.ascontiguousarray(H5).reshape(t, b, b) + Hs
    elif material == "macklin_mueller_neo_hookean":
        # macklin_mueller_neo_hookean.py:188-370: mu I + lam c c^T + (lam (J - 1) - mu) dc/dF
        Jd, c, dc = _det_cof_dcof(F)
        H5 = (
            mu4 * II
            + lam4 * np.einsum("tij,tkl->tijkl", c, c)
            + (lam4 * (Jd.reshape(-1, 1, 1, 1, 1) - 1.0) - mu4) * dc
        )
    elif material == "linear_elasticity":
        TT = np.einsum("il,jk->ijkl", eye, eye)[None]
        tr = np.einsum("ij,kl->ijkl", eye, eye)[None]
        H5 = np.broadcast_to(mu4 * (II + TT) + lam4 * tr, (t, dim, dim, dim, dim))
    else:
        raise ValueError("unknown material " + str(material))
    return np.ascontiguousarray(H5).reshape(t, b, b)


# --------------------------------------------------------------------------- #
# global tiers (_x / _u)                                                      #
# --------------------------------------------------------------------------- #
def _F_of(x, J, Jx_bar, dim):
    f = J @ np.asarray(x, dtype=np.float64).reshape(-1, 1)
    if Jx_bar is not None:
        f = f + np.asarray(Jx_bar).reshape(-1, 1)
    return np.asarray(f).reshape(-1, dim, dim)


def energy_x(material, x, J, mu, lam, vol, Jx_bar=None):
    """``float(sum(vol*psi))`` e.g. stable_neo_hookean.py:470-474 / :573-576."""
    dim = x.shape[1]
    F = _F_of(x, J, Jx_bar, dim)
    psi = energy_element_F(material, F, mu, lam)
    return float((np.asarray(vol).reshape(-1, 1) * psi).sum())


def gradient_x(material, x, J, mu, lam, vol, Jx_bar=None):
    """``J^T vec(vol*P)`` -> (n*d,1), e.g. stable_neo_hookean.py:498-503."""
    dim = x.shape[1]
    F = _F_of(x, J, Jx_bar, dim)
    P = gradient_element_F(material, F, mu, lam) * np.asarray(vol).reshape(-1, 1, 1)
    return np.asarray(J.T @ P.reshape(-1, 1))


def weighted_element_hessians(material, F, mu, lam, vol, psd=True, psd_before_vol=False):
    """Element blocks as they enter assembly.

    Per-material tiers floor eigenvalues *after* the vol weighting
    (stable_neo_hookean.py:533-535); the ``elastic`` dispatcher / ``_z`` tier
    floors *before* it (elastic.py:663-664, 776-780).  Linear elasticity's own
    module ignores ``psd`` (linear_elasticity.py:199-230).
    """
    He = hessian_element_F(material, F, mu, lam)
    w = np.asarray(vol, dtype=np.float64).reshape(-1, 1, 1)
    if psd_before_vol:
        if psd:
            He = psd_project(He)
        return He * w
    He = He * w
    if psd and material != "linear_elasticity":
        He = psd_project(He)
    return He


def hessian_x(material, x, J, mu, lam, vol, psd=True, Jx_bar=None, psd_before_vol=False):
    """``J^T blockdiag(He) J`` (csr), e.g. stable_neo_hookean.py:530-538."""
    dim = x.shape[1]
    F = _F_of(x, J, Jx_bar, dim)
    He = weighted_element_hessians(material, F, mu, lam, vol, psd, psd_before_vol)
    if sps.issparse(J):
        H = sps.block_diag(He)
        return J.T @ H @ J
    # dense J (reduced path)
    b = dim * dim
    Jr = np.asarray(J).reshape(-1, b, J.shape[1])
    return np.einsum("tbi,tbc,tcj->ij", Jr, He, Jr)


# --------------------------------------------------------------------------- #
# canonical pattern / slot map                                                #
# --------------------------------------------------------------------------- #
def canonical_csr(Q):
    """``tocsr(); sum_duplicates(); sort_indices()`` (SURVEY §7 hard part 1)."""
    Q = sps.csr_matrix(Q)
    Q.sum_duplicates()
    Q.sort_indices()
    return Q


def structural_pattern(T, n, dim):
    """Canonical structural CSR pattern = vertex adjacency (incl. self) (x) dim x dim,
    sorted; int32 ``(indptr, indices)`` plus block view ``(bptr, bcol)``."""
    T = np.asarray(T)
    k = T.shape[1]
    a = np.repeat(T, k, axis=1).ravel()
    b = np.tile(T, (1, k)).ravel()
    A = sps.csr_matrix((np.ones(a.shape[0], dtype=np.int8), (a, b)), shape=(n, n))
    A.sum_duplicates()
    A.sort_indices()
    bptr = A.indptr.astype(np.int32)
    bcol = A.indices.astype(np.int32)
    nb = np.diff(bptr)
    # scalar rows v*dim+i each hold nb[v]*dim entries
    rows_nnz = np.repeat(nb * dim, dim)
    indptr = np.zeros(n * dim + 1, dtype=np.int64)
    np.cumsum(rows_nnz, out=indptr[1:])
    blk_cols = (bcol[:, None].astype(np.int64) * dim + np.arange(dim)[None, :])  # (nnzb, dim)
    # for each vertex row v, its scalar rows repeat the same column list
    indices = np.empty(int(indptr[-1]), dtype=np.int32)
    for v in range(n):
        cols = blk_cols[bptr[v]:bptr[v + 1]].ravel()
        for i in range(dim):
            r = v * dim + i
            indices[indptr[r]:indptr[r + 1]] = cols
    return indptr.astype(np.int32), indices, bptr, bcol


def slot_map(T, indptr, indices, dim):
    """``slot[e,a,i,b,k] = indptr[r] + searchsorted(indices[row r], c)`` with
    ``r=T[e,a]*dim+i``, ``c=T[e,b]*dim+k`` (SURVEY §7 hard part 2).  int32."""
    T = np.asarray(T)
    t, k = T.shape
    slot = np.empty((t, k, dim, k, dim), dtype=np.int32)
    for e in range(t):
        for a in range(k):
            for i in range(dim):
                r = T[e, a] * dim + i
                row = indices[indptr[r]:indptr[r + 1]]
                for b in range(k):
                    for kk in range(dim):
                        c = T[e, b] * dim + kk
                        slot[e, a, i, b, kk] = indptr[r] + np.searchsorted(row, c)
    return slot


# --------------------------------------------------------------------------- #
# kinetic terms, line search, Newton, integrators                             #
# --------------------------------------------------------------------------- #
def be_target(x_curr, x_prev, h):
    """kinetic.py:87-90."""
    return x_curr + h * ((x_curr - x_prev) / h)


def _vel_bdf2(a, b, c, h):
    return (3.0 * a - 4.0 * b + c) / (2.0 * h)


def bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h):
    """kinetic.py:92-101."""
    v_curr = _vel_bdf2(x_curr, x_prev, x_prev2, h)
    v_prev = _vel_bdf2(x_prev, x_prev2, x_prev3, h)
    return (4.0 / 3.0) * x_curr - (1.0 / 3.0) * x_prev + (8.0 * h / 9.0) * v_curr - (2.0 * h / 9.0) * v_prev


def kinetic_energy(d, M, h, c):
    """kinetic.py:107-109."""
    return float((0.5 * c * (d.T @ M @ d) * (1 / (h ** 2))).item())


def kinetic_gradient(d, M, h, c):
    return c * (M @ d) * (1 / (h ** 2))


def kinetic_hessian(M, h, c):
    return M * (c / (h ** 2))


def backtracking_line_search(f, x0, g, dx, alpha=0.01, beta=0.5, max_iter=100, threshold=1e-12):
    """Armijo backtracking (backtracking_line_search.py:52-66)."""
    assert alpha > 0 and alpha <= 0.5
    assert beta > 0 and beta < 1
    assert np.ndim(x0) == np.ndim(dx)
    step = 1.0
    f0 = f(x0)
    slope = (g.T @ dx)
    for _ in range(max_iter):
        x = x0 + step * dx
        fx = f(x)
        if fx <= f0 + alpha * step * slope + threshold:
            return step, x, fx
        step = beta * step
    return 0.0, x0, f0


def newton_solver(x0, energy_func, gradient_func, hessian_func, tolerance=1e-6, max_iter=1,
                  do_line_search=True, return_info=False, linear_solver=None):
    """solvers/newton.py:42-75.  ``linear_solver(H, rhs)`` overrides the reference's
    direct solve (SuperLU / LAPACK) for sizes where those are infeasible."""
    x = x0.copy()
    info = {"g": [], "dx": [], "alphas": [], "iters": -1}
    for it in range(max_iter):
        g = gradient_func(x)
        H = hessian_func(x)
        if linear_solver is not None:
            dx = np.asarray(linear_solver(H, -g)).reshape(-1, 1)
        elif sps.issparse(H):
            dx = spla.spsolve(H.tocsc(), -g).reshape(-1, 1)
        else:
            dx = scipy.linalg.solve(H, -g).reshape(-1, 1)
        if do_line_search:
            a, _, _ = backtracking_line_search(energy_func, x, g, dx)
        else:
            a = 1.0
        x += a * dx
        info["g"].append(g)
        info["dx"].append(dx)
        info["alphas"].append(a)
        info["iters"] = it
        if np.linalg.norm(a * dx) < tolerance:
            break
    return (x, info) if return_info else x


def backward_euler(x_curr, x_prev, energy_func, gradient_func, hessian_func, M, h, tolerance=1e-6,
                   max_iter=1, do_line_search=True, return_info=False, linear_solver=None):
    """integrators/backward_euler.py:27-91 (c = 1)."""
    xt = be_target(x_curr, x_prev, h)

    def E(x):
        return energy_func(x) + kinetic_energy(x - xt, M, h, 1.0)

    def G(x):
        return gradient_func(x) + kinetic_gradient(x - xt, M, h, 1.0)

    def Hf(x):
        return hessian_func(x) + kinetic_hessian(M, h, 1.0)

    return newton_solver(xt, E, G, Hf, tolerance, max_iter, do_line_search, return_info, linear_solver)


def bdf2(x_curr, x_prev, x_prev2, x_prev3, energy_func, gradient_func, hessian_func, M, h,
         tolerance=1e-6, max_iter=1, do_line_search=True, return_info=False, linear_solver=None):
    """integrators/bdf2.py:31-101 (c = 9/4)."""
    xt = bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h)
    c = 9.0 / 4.0

    def E(x):
        return energy_func(x) + kinetic_energy(x - xt, M, h, c)

    def G(x):
        return gradient_func(x) + kinetic_gradient(x - xt, M, h, c)

    def Hf(x):
        return hessian_func(x) + kinetic_hessian(M, h, c)

    return newton_solver(xt, E, G, Hf, tolerance, max_iter, do_line_search, return_info, linear_solver)


def block_jacobi_cg(H, rhs, dim, rtol=1e-12, maxiter=20000):
    """CPU stand-in for ``spsolve`` at sizes where SuperLU is infeasible
    (BASELINE.md §3): scipy CG with dim x dim block-Jacobi.  Not reference code."""
    H = sps.csr_matrix(H)
    n = H.shape[0] // dim
    Hb = sps.bsr_matrix(H, blocksize=(dim, dim))
    Hb.sort_indices()
    diag = np.zeros((n, dim, dim))
    rows = np.repeat(np.arange(n), np.diff(Hb.indptr))
    sel = Hb.indices == rows
    diag[rows[sel]] = Hb.data[sel]
    inv = np.linalg.inv(diag)

    def prec(v):
        return np.einsum("nij,nj->ni", inv, v.reshape(n, dim)).reshape(-1)

    Mop = spla.LinearOperator(H.shape, matvec=prec)
    its = [0]

    def cb(_):
        its[0] += 1

    x, _ = spla.cg(H, np.asarray(rhs).reshape(-1), rtol=rtol, atol=0.0, maxiter=maxiter, M=Mop, callback=cb)
    return x, its[0]


def block_jacobi_cg_single_reduction(H, rhs, dim, rtol=1e-10, maxiter=20000):
    """Prototype (not reference code) of the single-reduction PCG planned for the distributed solve (DESIGN.md
    section 8, item 5): Chronopoulos-Gear recurrences -- u = M^-1 r, w = A u, then ONE reduction for
    (r.u, w.u, r.r) per iteration instead of the two of the textbook loop (p.q, then r.z and r.r).  Same
    block-Jacobi preconditioner as ``block_jacobi_cg``; returns (x, iterations).  In exact arithmetic the iterates are
    those of standard PCG."""
    H = sps.csr_matrix(H)
    n = H.shape[0] // dim
    Hb = sps.bsr_matrix(H, blocksize=(dim, dim))
    Hb.sort_indices()
    diag = np.zeros((n, dim, dim))
    rows = np.repeat(np.arange(n), np.diff(Hb.indptr))
    sel = Hb.indices == rows
    diag[rows[sel]] = Hb.data[sel]
    inv = np.linalg.inv(diag)
    prec = lambda v: np.einsum("nij,nj->ni", inv, v.reshape(n, dim)).reshape(-1)  # noqa: E731
    b = np.asarray(rhs, dtype=np.float64).reshape(-1)
    x = np.zeros_like(b)
    r = b.copy()
    bb = float(b @ b)
    if not bb > 0.0:
        return x, 0
    p = np.zeros_like(b)
    s = np.zeros_like(b)
    gamma_old, alpha = 1.0, 1.0
    for it in range(maxiter):
        u = prec(r)
        w = H @ u
        gamma, delta, rr = float(r @ u), float(w @ u), float(r @ r)     # the one reduction of the iteration
        if not rr > rtol * rtol * bb:
            return x, it
        if it == 0:
            beta, alpha = 0.0, gamma / delta
        else:
            beta = gamma / gamma_old
            alpha = gamma / (delta - beta * gamma / alpha)
        p = u + beta * p
        s = w + beta * s                                                # s = A p without a second SpMV
        x = x + alpha * p
        r = r - alpha * s
        gamma_old = gamma
    return x, maxiter


def orthonormalize(B, M=None, threshold=1e-16):
    """orthonormalize.py:9-49."""
    if M is None:
        M = sps.identity(B.shape[0])
    msqrt = np.sqrt(M.diagonal())
    Bm = sps.diags(msqrt, 0) @ B
    Q, R = np.linalg.qr(Bm.toarray() if sps.issparse(Bm) else Bm)
    nonsing = np.abs(R).sum(axis=1) > threshold
    return sps.diags(1.0 / msqrt, 0) @ Q[:, nonsing]


def project_into_subspace(y, B, M=None, BMB=None, BMy=None):
    """project_into_subspace.py:9-59."""
    if M is None:
        M = sps.identity(y.shape[0])
    if BMy is None:
        BMy = B.T @ M @ y
    if BMB is None:
        BMB = B.T @ M @ B
    if sps.issparse(BMB):
        return spla.spsolve(BMB, BMy).reshape(-1, 1)
    return np.linalg.solve(BMB, BMy).reshape(-1, 1)


def average_onto_simplex(A, T):
    """average_onto_simplex.py:8-37."""
    A = np.asarray(A, dtype=np.float64)
    return sum(A[T[:, c], :] / T.shape[1] for c in range(T.shape[1]))


def spectral_clustering(W, k, D=None, seed=0):
    """spectral_clustering.py:9-48: scipy's kmeans2 with k-means++ seeding on the weighted rows (scipy IS the
    reference's algorithm here; simkit_b200 restates it on the GPU)."""
    from scipy.cluster.vq import kmeans2
    B = np.asarray(W, dtype=np.float64) * (np.ones((W.shape[0], 1)) if D is None else D)
    c, l = kmeans2(B, k, seed=seed, minit="++")
    return l, c


def spectral_cubature(X, T, W, k):
    """spectral_cubature.py:18-74 -> (lI, mc, labels, centroids)."""
    Wt = average_onto_simplex(W, T)
    labels, centroids = spectral_clustering(Wt, k)
    D = np.linalg.norm(centroids[:, None, :] - Wt[None, :, :], axis=2)
    return np.argmin(D, axis=1), np.bincount(labels, volume(X, T).flatten()), labels, centroids


def lbs_jacobian(V, W):
    """lbs_jacobian.py:12-47."""
    n, d = V.shape
    k = W.shape[1]
    V1 = np.hstack((V, np.ones((n, 1))))
    Wexp = np.kron(W, np.ones((1, d + 1)))
    V1exp = np.kron(np.ones((1, k)), V1)
    return np.kron(Wexp * V1exp, np.identity(d))


def skinning_eigenmodes(X, T, k, mu=1, bI=None):
    """skinning_eigenmodes.py:19-84 with the shift-invert ARPACK call of eigs.py:96-103.  The reference's inverse
    operator is a UMFPACK LU (cvxopt, absent here: PARITY UNPINNED for this function -- the reference itself cannot be
    run); this restatement lets scipy factor ``L + eps M`` with SuperLU, the same tiny regularisation the product uses."""
    M = sps.csc_matrix(massmatrix(X, T))
    L = sps.csc_matrix(dirichlet_laplacian(X, T, mu))
    Ii = np.arange(X.shape[0]) if bI is None else np.setdiff1d(np.arange(X.shape[0]), bI)
    Lf = L[Ii, :][:, Ii]
    Mf = sps.diags(M.diagonal()[Ii]).tocsc()
    eps = 1e-10 * Lf.diagonal().sum() / Mf.diagonal().sum()
    lu = spla.splu((Lf + eps * Mf).tocsc())
    op = spla.LinearOperator(Lf.shape, matvec=lu.solve, dtype=np.float64)
    E, Wi = spla.eigsh(Lf, M=Mf, k=k, sigma=0, which="LM", OPinv=op, v0=np.ones(Lf.shape[0]))
    W = np.zeros((X.shape[0], k))
    W[Ii] = Wi
    return W, E - eps, lbs_jacobian(X, W)


def rigid_mode_prolongator(X, agg):
    """P (n*dim, NC*n_agg) of the two-level preconditioner of simkit_b200/csrc/coarse.cuh (not reference code): the
    rigid-body modes of vertex aggregates, P_v = [I | -[x_v - c_I]_x] (3 x 6; 2 x 3 in 2D), c_I the aggregate's centroid."""
    X = np.asarray(X, dtype=np.float64)
    n, dim = X.shape
    agg = np.asarray(agg)
    n_agg = int(agg.max()) + 1
    NC = 6 if dim == 3 else 3
    cnt = np.bincount(agg, minlength=n_agg).astype(np.float64)
    cen = np.stack([np.bincount(agg, weights=X[:, a], minlength=n_agg) / cnt for a in range(dim)], axis=1)
    xr = X - cen[agg]
    rows, cols, vals = [], [], []
    for v in range(n):
        I = agg[v]
        for i in range(dim):
            rows.append(v * dim + i); cols.append(I * NC + i); vals.append(1.0)
        x = xr[v]
        if dim == 3:
            # u + w x x :  out = c[:3] + cross(c[3:], x)
            ent = [(0, 4, x[2]), (0, 5, -x[1]), (1, 5, x[0]), (1, 3, -x[2]), (2, 3, x[1]), (2, 4, -x[0])]
        else:
            ent = [(0, 2, -x[1]), (1, 2, x[0])]
        for i, a, val in ent:
            rows.append(v * dim + i); cols.append(I * NC + a); vals.append(val)
    return sps.csr_matrix((vals, (rows, cols)), shape=(n * dim, NC * n_agg))


def two_level_cg_single_reduction(H, rhs, dim, P=None, rtol=1e-10, maxiter=20000):
    """Prototype (not reference code) of csrc/capi_pcg2.cu, statement for statement: Chronopoulos-Gear PCG with the
    additive two-level preconditioner M^-1 = D^-1 + P (P^T H P)^-1 P^T, whose restricted residual is carried by the
    recurrence  P^T r_{k+1} = P^T r_k - alpha (P^T w_k + beta P^T s_{k-1})  so that the one reduction of an iteration
    (gamma = r.u, delta = w.u, r.r and P^T w) is the only collective.  A bootstrap pass with alpha = beta = 0 produces
    u_0, w_0 and the first reduction.  ``P=None``: block-Jacobi only.  Returns (x, iterations)."""
    H = sps.csr_matrix(H)
    n = H.shape[0] // dim
    Hb = sps.bsr_matrix(H, blocksize=(dim, dim))
    Hb.sort_indices()
    diag = np.zeros((n, dim, dim))
    rows = np.repeat(np.arange(n), np.diff(Hb.indptr))
    sel = Hb.indices == rows
    diag[rows[sel]] = Hb.data[sel]
    inv = np.linalg.inv(diag)
    dinv = lambda v: np.einsum("nij,nj->ni", inv, v.reshape(n, dim)).reshape(-1)  # noqa: E731
    b = np.asarray(rhs, dtype=np.float64).reshape(-1)
    nd = b.size
    x, r = np.zeros(nd), b.copy()
    u, w, p, s = np.zeros(nd), np.zeros(nd), np.zeros(nd), np.zeros(nd)
    if P is not None:
        Ainv = np.linalg.inv((P.T @ H @ P).toarray())
        rc = P.T @ r                       # the one extra reduction of the solve
        rcs = np.zeros_like(rc)
        red_c = np.zeros_like(rc)
    red = np.zeros(3)
    gamma_old = alpha_old = 1.0
    bb = rr = 0.0
    stage = iters = 0
    for _ in range(maxiter + 1):
        # --- pcg2_scalars_kernel
        alpha = beta = 0.0
        if stage > 0:
            gamma, delta, rr = red
            if stage == 1:
                bb = rr
            if not bb > 0.0 or not rr > rtol * rtol * bb:
                break
            beta = 0.0 if stage == 1 else gamma / gamma_old
            den = delta if stage == 1 else delta - beta * gamma / alpha_old
            if not den > 0.0 or not gamma > 0.0:
                raise RuntimeError("breakdown")
            alpha = gamma / den
            gamma_old, alpha_old = gamma, alpha
            iters += 1
        stage += 1
        if P is not None:
            rcs = red_c + beta * rcs
            rc = rc - alpha * rcs
            zc = Ainv @ rc                 # --- pcg2_gemv_kernel
        # --- pcg2_update_kernel
        p = u + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        u = dinv(r)
        if P is not None:
            u = u + P @ zc
        # --- halo exchange of u, pcg2_spmv_kernel, pcg2_restrict_kernel, the all-reduce
        w = H @ u
        red = np.array([r @ u, w @ u, r @ r])
        if P is not None:
            red_c = P.T @ w
    return x, iters


# --------------------------------------------------------------------------- #
# reduced operators                                                           #
# --------------------------------------------------------------------------- #
def fst_precompute(A, B, l, dim=3):
    """``ARBs[p,q,c,i,j] = sum_{t in c} sum_k A[p, b t + d i + k] B[b t + d j + k, q]``
    (fast_sandwich_transform_clustered.py:66-93)."""
    A = sps.csr_matrix(A)
    B = sps.csr_matrix(B)
    b = dim * dim
    l = np.asarray(l).reshape(-1)
    nc = int(l.max()) + 1
    out = np.zeros((A.shape[0], B.shape[1], nc, dim, dim))
    Ad = A.toarray().reshape(A.shape[0], -1, dim, dim)     # [p, t, i, k]
    Bd = B.toarray().reshape(-1, dim, dim, B.shape[1])     # [t, j, k, q]
    for c in range(nc):
        sel = np.where(l == c)[0]
        out[:, :, c] = np.einsum("ptik,tjkq->pqij", Ad[:, sel], Bd[sel])
    return out


def fst_eval(ARBs, r, dim=3):
    """fast_sandwich_transform_clustered.py:150-158."""
    r = np.asarray(r).reshape(-1, dim, dim)
    return np.einsum("pqcij,cij->pq", ARBs, r)
