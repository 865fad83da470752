"""Freeze reference outputs into ``tests/golden/*.npz`` (run in the build container).

The reference (otmanon/simkit @ 4e19c36) is imported read-only from
``/root/reference``; inputs are seeded and stored next to the outputs so the
fixtures replay without the reference (it does not exist on the GPU box).

    python oracle/make_golden.py
"""

import os
import sys

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import simkit  # noqa: E402
import simkit.energies as ske  # noqa: E402
from simkit.integrators import backward_euler as ref_be  # noqa: E402
from simkit.solvers import newton_solver as ref_newton  # noqa: E402
from simkit.fast_sandwich_transform_clustered import fast_sandwich_transform_clustered as ref_fst  # noqa: E402
from simkit.rotation_gradient import rotation_gradient_F as ref_rotgrad  # noqa: E402

from simkit_b200 import synthetic as syn  # noqa: E402
from oracle.mfem_problem import mfem_problem  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MATS = {
    "stable_neo_hookean": True,
    "neo_hookean": True,
    "arap": False,
    "stvk": True,
    "linear_elasticity": True,
    "fcr": True,
    "macklin_mueller_neo_hookean": True,
}
DISPATCH = (("arap", "arap"), ("linear_elasticity", "linear-elasticity"), ("fcr", "fcr"),
            ("macklin_mueller_neo_hookean", "macklin-mueller-neo-hookean"))


def canon(Q):
    Q = sps.csr_matrix(Q)
    Q.sum_duplicates()
    Q.sort_indices()
    return Q


def golden_mesh(tag, cells, sigma, seed):
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    ext = tuple(1.0 for _ in cells)
    t = T.shape[0]
    rng = np.random.default_rng(seed)
    # shuffle vertex numbering so the caller's order is not lexicographic
    perm = rng.permutation(X.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0])
    X = X[perm]
    T = inv[T]
    U = syn.jittered_state(X, cells, ext, sigma=sigma, seed=seed)
    mu, lam = syn.heterogeneous_lame(t, seed=seed + 1)
    J = simkit.deformation_jacobian(X, T)
    vol = simkit.volume(X, T)
    F = np.asarray(J @ U.reshape(-1, 1)).reshape(-1, dim, dim)
    out = dict(X=X, T=T, U=U, mu=mu, lam=lam, vol=vol, F=F, dim=dim, sigma=sigma,
               J_data=canon(J).data, J_indices=canon(J).indices, J_indptr=canon(J).indptr)
    R, S = simkit.polar_svd(F)
    out["polar_R"], out["polar_S"] = R, S
    out["rotgrad"] = ref_rotgrad(F)
    xb = X + 0.02 * rng.standard_normal(X.shape)
    out["x_bar"] = xb
    Jxb = J @ xb.reshape(-1, 1)
    inverted = bool((np.linalg.det(F) <= 0).any())
    out["inverted"] = inverted
    for m, has_lam in MATS.items():
        if m == "neo_hookean" and inverted:
            continue
        a = (mu, lam) if has_lam else (mu,)
        g = lambda kind, tier: getattr(ske, f"{m}_{kind}_{tier}")  # noqa: E731
        out[f"{m}_psi"] = g("energy", "element_F")(F, *a)
        out[f"{m}_P"] = g("gradient", "element_F")(F, *a)
        out[f"{m}_He"] = g("hessian", "element_F")(F, *a)
        out[f"{m}_E"] = g("energy", "x")(U, J, *a, vol)
        out[f"{m}_g"] = g("gradient", "x")(U, J, *a, vol)
        for psd in (True, False):
            Q = canon(g("hessian", "x")(U, J, *a, vol, psd=psd))
            k = f"{m}_Q_psd{int(psd)}"
            out[k + "_data"], out[k + "_indices"], out[k + "_indptr"] = Q.data, Q.indices, Q.indptr
        out[f"{m}_E_u"] = g("energy", "u")(U - xb, J, Jxb, *a, vol)
        out[f"{m}_g_u"] = g("gradient", "u")(U - xb, J, Jxb, *a, vol)
    # elastic dispatcher (psd floor *before* vol): every routed material
    for m, name in DISPATCH:
        Q = canon(ske.elastic_hessian_x(U, J, mu, lam, vol, name, psd=True))
        k = f"{m}_Qdisp"
        out[k + "_data"], out[k + "_indices"], out[k + "_indptr"] = Q.data, Q.indices, Q.indptr
        out[f"{m}_Edisp"] = ske.elastic_energy_x(U, J, mu, lam, vol, name)
        out[f"{m}_gdisp"] = ske.elastic_gradient_x(U, J, mu, lam, vol, name)
        out[f"{m}_Hedisp"] = ske.elastic_hessian_element_F(F, mu, lam, name, psd=True)
    # stretch (S) tier: symmetric stretches of the polar decomposition, full and compact form
    from simkit.symmetric_stretch_map import symmetric_stretch_map
    _, Sei = symmetric_stretch_map(1, dim)
    Sc = S.reshape(t, dim * dim) @ np.asarray(Sei.todense()).T
    out["S_compact"] = Sc
    for name, tag_s in (("arap", "arap"), ("macklin-mueller-neo-hookean", "mm")):
        for form, Sin in (("full", S), ("compact", Sc)):
            out[f"{tag_s}_S_{form}_E"] = ske.elastic_energy_S(Sin, mu, lam, vol, name)
            out[f"{tag_s}_S_{form}_g"] = ske.elastic_gradient_S(Sin, mu, lam, vol, name)
            out[f"{tag_s}_S_{form}_H"] = ske.elastic_hessian_S(Sin, mu, lam, vol, name)
    # psd_project on arbitrary symmetric blocks
    A = rng.standard_normal((40, dim * dim, dim * dim))
    A = A + np.swapaxes(A, 1, 2)
    out["psd_in"] = A
    out["psd_proj"] = simkit.psd_project(A)
    out["psd_abs"] = simkit.psd_project(A, "abs")
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag, "t =", t, "inverted:", inverted)


def golden_step(tag, cells, seed):
    """One backward-Euler step (3 Newton iterations, line search) and plain Newton."""
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    ext = tuple(1.0 for _ in cells)
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    J = simkit.deformation_jacobian(X, T)
    vol = simkit.volume(X, T)
    Mv = sps.kron(simkit.massmatrix(X, T, rho), sps.identity(dim)).tocsc()
    fg = simkit.gravity_force(X, T, -9.8, rho).reshape(-1, 1)
    out = dict(X=X, T=T, mu=mu, lam=lam, rho=rho, h=h, dim=dim, mass_diag=Mv.diagonal(), fg=fg)
    for m, has_lam in MATS.items():
        a = (mu, lam) if has_lam else (mu,)
        e_x = getattr(ske, f"{m}_energy_x")
        g_x = getattr(ske, f"{m}_gradient_x")
        h_x = getattr(ske, f"{m}_hessian_x")

        def E(x):
            return e_x(x.reshape(-1, dim), J, *a, vol) - float((fg.T @ x).item())

        def G(x):
            return g_x(x.reshape(-1, dim), J, *a, vol) - fg

        def H(x):
            return h_x(x.reshape(-1, dim), J, *a, vol)

        x_curr = syn.jittered_state(X, cells, ext, sigma=0.05, seed=seed).reshape(-1, 1)
        x_prev = X.reshape(-1, 1)
        x, info = ref_be(x_curr, x_prev, E, G, H, Mv, h, max_iter=3, return_info=True)
        out[f"{m}_be_x_curr"], out[f"{m}_be_x_prev"] = x_curr, x_prev
        out[f"{m}_be_x"] = x
        out[f"{m}_be_alphas"] = np.array(info["alphas"])
        out[f"{m}_be_iters"] = info["iters"]
        out[f"{m}_be_dx0"] = info["dx"][0]
        out[f"{m}_be_g0"] = info["g"][0]
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag)


def golden_reduced(tag, cells, r, seed):
    """Reduced Hessian through the `_u` tier with a dense operator and FST."""
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    ext = tuple(1.0 for _ in cells)
    t = T.shape[0]
    mu, lam = syn.lame()
    J = simkit.deformation_jacobian(X, T)
    vol = simkit.volume(X, T)
    B = syn.smooth_modes(X, r, seed=seed)
    rng = np.random.default_rng(seed)
    z = 0.02 * rng.standard_normal((r, 1))
    JB = np.asarray(J @ B)
    Jx0 = np.asarray(J @ X.reshape(-1, 1))
    out = dict(X=X, T=T, B=B, z=z, mu=mu, lam=lam, vol=vol, dim=dim)
    zz = z.reshape(-1, 1)

    class _Z:  # the `_u` functions only read ``u.shape[1]`` and ``u.reshape(-1,1)``
        pass

    for m, has_lam in (("stable_neo_hookean", True), ("arap", False)):
        a = (mu, lam) if has_lam else (mu,)
        u = np.zeros((r // dim if r % dim == 0 else r, dim)) if False else None
        # call with u of shape (r/dim, dim) as example 011 does (z.reshape(-1, dim))
        assert r % dim == 0
        u = zz.reshape(-1, dim)
        out[f"{m}_Hr"] = getattr(ske, f"{m}_hessian_u")(u, JB, Jx0, *a, vol)
        out[f"{m}_gr"] = getattr(ske, f"{m}_gradient_u")(u, JB, Jx0, *a, vol)
        out[f"{m}_Er"] = getattr(ske, f"{m}_energy_u")(u, JB, Jx0, *a, vol)
    # FST
    m1, m2, nc = 6, 5, 4
    A = rng.standard_normal((m1, dim * dim * t))
    Bs = sps.random(dim * dim * t, m2, density=0.2, random_state=seed, format="csr")
    l = rng.integers(0, nc, size=t)
    l[:nc] = np.arange(nc)
    f = ref_fst(A, Bs, l, dim=dim)
    rr = rng.standard_normal((nc, dim, dim))
    out.update(fst_A=A, fst_B=Bs.toarray(), fst_l=l, fst_ARBs=f.ARBs, fst_r=rr, fst_out=f(rr))
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag)


def golden_contact(tag, cells, seed):
    """Plane contact springs and a backward-Euler step (3 Newton iterations) of stable neo-Hookean + contact."""
    from simkit.energies.contact_springs_plane import (contact_springs_plane_energy as ce, contact_springs_plane_gradient as cg,
                                                       contact_springs_plane_hessian as ch)
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    ext = tuple(1.0 for _ in cells)
    rng = np.random.default_rng(seed)
    mu, lam = syn.lame()
    rho, h, k = 1e3, 1e-2, 1e5
    nrm = np.zeros(dim)
    nrm[1] = 1.0
    nrm[0] = 0.3
    nrm = nrm / np.linalg.norm(nrm)
    pt = np.full(dim, 0.25)
    U = syn.jittered_state(X, cells, ext, sigma=0.2, seed=seed)
    Mv = simkit.massmatrix(X, T, rho)
    E, inds = ce(U, k, pt, nrm, Mv, return_contact_inds=True)
    out = dict(X=X, T=T, U=U, dim=dim, k=k, p=pt, n=nrm, mass=Mv.diagonal(), E=E, inds=inds.ravel(),
               g=cg(U, k, pt, nrm, Mv), H=ch(U, k, pt, nrm, Mv).toarray(), E_noM=ce(U, k, pt, nrm),
               g_noM=cg(U, k, pt, nrm), E_above=ce(U + 10.0 * nrm, k, pt, nrm, Mv))
    from simkit.energies.contact_springs_sphere import (contact_springs_sphere_energy as se, contact_springs_sphere_gradient as sg,
                                                        contact_springs_sphere_hessian as sh)
    sc = np.full(dim, 0.55)
    sr = 0.3
    out.update(s_p=sc, s_r=sr, s_E=se(U, k, sc, sr, Mv), s_g=sg(U, k, sc, sr, Mv), s_H=sh(U, k, sc, sr, Mv).toarray(),
               s_E_noM=se(U, k, sc, sr), s_E_far=se(U + 10.0, k, sc, sr, Mv))
    J = simkit.deformation_jacobian(X, T)
    vol = simkit.volume(X, T)
    Md = sps.kron(Mv, sps.identity(dim)).tocsc()
    fg = simkit.gravity_force(X, T, -9.8, rho).reshape(-1, 1)

    def En2(x):
        return En(x) + se(x.reshape(-1, dim), k, sc, sr, Mv)

    def Gr2(x):
        return Gr(x) + sg(x.reshape(-1, dim), k, sc, sr, Mv)

    def He2(x):
        return He(x) + sh(x.reshape(-1, dim), k, sc, sr, Mv)

    def En(x):
        return ske.stable_neo_hookean_energy_x(x.reshape(-1, dim), J, mu, lam, vol) - float((fg.T @ x).item()) + ce(x.reshape(-1, dim), k, pt, nrm, Mv)

    def Gr(x):
        return ske.stable_neo_hookean_gradient_x(x.reshape(-1, dim), J, mu, lam, vol) - fg + cg(x.reshape(-1, dim), k, pt, nrm, Mv)

    def He(x):
        return ske.stable_neo_hookean_hessian_x(x.reshape(-1, dim), J, mu, lam, vol) + ch(x.reshape(-1, dim), k, pt, nrm, Mv)

    x_curr = U.reshape(-1, 1)
    x, info = ref_be(x_curr, X.reshape(-1, 1), En, Gr, He, Md, h, max_iter=3, return_info=True)
    out.update(mu=mu, lam=lam, rho=rho, h=h, fg=fg, be_x=x, be_alphas=np.array(info["alphas"]))
    x2, info2 = ref_be(x_curr, X.reshape(-1, 1), En2, Gr2, He2, Md, h, max_iter=3, return_info=True)
    out.update(be2_x=x2, be2_alphas=np.array(info2["alphas"]))
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag, "contacting", inds.size, "inside sphere", int((np.linalg.norm(U - sc, axis=1) < sr).sum()))


def golden_quadratic(tag, cells, seed):
    """quadratic_* and dirichlet_penalty, and a backward-Euler step (3 Newton iterations) of stable neo-Hookean +
    gravity + a general sparse quadratic term: pinned vertices (dirichlet_penalty) plus anisotropic springs along the
    mesh edges (full dim x dim blocks on and off the block diagonal, all inside the mesh's CSR pattern)."""
    from simkit.energies.quadratic import quadratic_energy as qe, quadratic_gradient as qg, quadratic_hessian as qh
    from simkit.dirichlet_penalty import dirichlet_penalty as dp
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    ext = tuple(1.0 for _ in cells)
    rng = np.random.default_rng(seed)
    nv = X.shape[0]
    mu, lam = syn.lame()
    rho, h = 1e3, 1e-2
    U = syn.jittered_state(X, cells, ext, sigma=0.2, seed=seed)
    bI = np.where(X[:, 0] == 0.0)[0]
    y = X[bI] + 0.02 * rng.standard_normal((bI.size, dim))
    gamma = 1e6 * (1.0 + rng.random(bI.size))
    Qd, bd = dp(bI, y, nv, gamma)
    Qs, bs = dp(bI, y, nv, 1e6)
    edges = set()
    for el in T:
        for a in range(dim + 1):
            for c in range(a + 1, dim + 1):
                edges.add((min(el[a], el[c]), max(el[a], el[c])))
    edges = np.array(sorted(edges))[::3]                       # every third edge carries a spring
    rows, cols, vals = [], [], []
    for v, w in edges:
        A = rng.standard_normal((dim, dim))
        A = 2e3 * (A @ A.T + 0.1 * np.eye(dim))
        for (r0, c0, sg) in ((v, v, 1.0), (w, w, 1.0), (v, w, -1.0), (w, v, -1.0)):
            for i in range(dim):
                for j in range(dim):
                    rows.append(r0 * dim + i)
                    cols.append(c0 * dim + j)
                    vals.append(sg * A[i, j])
    Qe = sps.csr_matrix((vals, (rows, cols)), (nv * dim, nv * dim))
    Q = canon(Qe + Qd)
    b = bd + 5.0 * rng.standard_normal((nv * dim, 1))
    x = U.reshape(-1, 1)
    assert qh(Q) is Q
    mu_het = 1.0 + rng.random((T.shape[0], 1))
    lap = dict(mu_het=mu_het, L_het=simkit.dirichlet_laplacian(X, T, mu_het).toarray(),
               Lv_scalar=simkit.dirichlet_laplacian(X, T, 2.5, vector=True).toarray())
    out = dict(X=X, T=T, U=U, dim=dim, bI=bI, y=y, gamma=gamma, Qd=Qd.toarray(), bd=bd, Qs=Qs.toarray(), bs=bs,
               Q_data=Q.data, Q_indices=Q.indices, Q_indptr=Q.indptr, b=b, E=qe(x, Q, b), g=qg(x, Q, b))
    J = simkit.deformation_jacobian(X, T)
    vol = simkit.volume(X, T)
    Mv = simkit.massmatrix(X, T, rho)
    Md = sps.kron(Mv, sps.identity(dim)).tocsc()
    fg = simkit.gravity_force(X, T, -9.8, rho).reshape(-1, 1)

    def En(x):
        return ske.stable_neo_hookean_energy_x(x.reshape(-1, dim), J, mu, lam, vol) - float((fg.T @ x).item()) + qe(x, Q, b)

    def Gr(x):
        return ske.stable_neo_hookean_gradient_x(x.reshape(-1, dim), J, mu, lam, vol) - fg + qg(x, Q, b)

    def He(x):
        return ske.stable_neo_hookean_hessian_x(x.reshape(-1, dim), J, mu, lam, vol) + qh(Q)

    xn, info = ref_be(x, X.reshape(-1, 1), En, Gr, He, Md, h, max_iter=3, return_info=True)
    out.update(mu=mu, lam=lam, rho=rho, h=h, fg=fg, mass=Mv.diagonal(), be_x=xn, be_alphas=np.array(info["alphas"]))
    out.update(lap)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag, "pinned", bI.size, "springs", len(edges), "alphas", info["alphas"])


def golden_mfem(tag, cells, rho_aug, seed):
    """MFEM blocks (stretch, dS/dF, ds/dz, symmetric stretch map) and three SQP iterations of the mixed solver."""
    from simkit.solvers import sqp_mfem as ref_sqp
    dim = len(cells)
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(seed)
    F = np.eye(dim)[None] + 0.35 * rng.standard_normal((60, dim, dim))
    out = dict(X=X, T=T, dim=dim, rho_aug=rho_aug, F=F, stretch=simkit.stretch(F), dSdF=simkit.stretch_gradient_dF(F))
    prob = mfem_problem(simkit, X, T, rho_aug)
    p0 = prob["p0"]
    u0 = p0[:prob["nz"]]
    out["dsdz"] = simkit.stretch_gradient_dz(u0, prob["GJB"], Ci=prob["Ci"], dim=dim, GJq=prob["GJq"]).toarray()
    out["p0"] = p0
    out["energy0"] = prob["energy"](p0)
    fb = prob["grad_blocks"](p0)
    out["f_u0"], out["f_z0"], out["f_mu0"] = (np.asarray(v) for v in fb)
    out["p3"] = ref_sqp(p0, prob["energy"], prob["hess_blocks"], prob["grad_blocks"], tolerance=1e-12, max_iter=3)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag)


def golden_subspace(tag, cells, seed):
    """orthonormalize / project_into_subspace of the reference on a small basis with a dependent column."""
    from simkit.orthonormalize import orthonormalize
    from simkit.project_into_subspace import project_into_subspace
    import scipy.sparse as sps
    rng = np.random.default_rng(seed)
    X, T = syn.make_mesh(cells)
    dim = X.shape[1]
    nd = X.size
    M = sps.kron(simkit.massmatrix(X, T, 1e3), sps.identity(dim)).tocsc()
    B = rng.standard_normal((nd, 9))
    B[:, 8] = B[:, 2] - 0.5 * B[:, 4]      # dependent LAST column: its row of R vanishes and the reference drops it
    y = rng.standard_normal((nd, 1))
    out = dict(X=X, T=T, dim=dim, B=B, y=y, mass_diag=M.diagonal(),
               ortho_mass=np.asarray(orthonormalize(B, M, 1e-8)), ortho_id=np.asarray(orthonormalize(B[:, :6])),
               z_mass=project_into_subspace(y, B[:, :6], M), z_id=project_into_subspace(y, B[:, :6]))
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag)


def golden_cubature(tag, cells, p, k, seed):
    """spectral_cubature of the reference (scipy kmeans2 with k-means++ seeding, seed 0) on smooth modes."""
    from simkit.spectral_cubature import spectral_cubature
    from simkit.spectral_clustering import spectral_clustering
    rng = np.random.default_rng(seed)
    X, T = syn.make_mesh(cells)
    X = X + 0.2 * syn.cell_size(cells, tuple(1.0 for _ in cells)) * rng.standard_normal(X.shape)
    W = np.cos(X @ (2.0 * np.pi * rng.standard_normal((X.shape[1], p))) + rng.random((1, p)))   # smooth per-vertex modes (n, p)
    lI, mc, labels, cen = spectral_cubature(X, T, W, k, return_labels=True, return_centroids=True)
    Dw = 0.5 + rng.random((X.shape[0], 1))
    l2, c2 = spectral_clustering(W, k, D=Dw, seed=3)
    out = dict(X=X, T=T, W=W, k=k, lI=lI, mc=mc, labels=labels, centroids=cen, Dw=Dw, labels_w=l2, centroids_w=c2)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "cubature":
        golden_cubature("cubature_tet", (7, 6, 5), 8, 12, 80)
        golden_cubature("cubature_tri", (24, 19), 6, 9, 81)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "subspace":      # add the newer fixtures without rewriting the old ones
        golden_subspace("subspace_tet", (3, 2, 2), 70)
        golden_subspace("subspace_tri", (5, 4), 71)
        sys.exit(0)
    golden_mesh("tet_s01", (3, 2, 2), 0.1, 10)
    golden_mesh("tet_s04", (3, 2, 2), 0.4, 11)
    golden_mesh("tri_s01", (5, 4), 0.1, 12)
    golden_mesh("tri_s04", (5, 4), 0.4, 13)
    golden_step("step_tet", (3, 3, 2), 20)
    golden_step("step_tri", (6, 5), 21)
    golden_reduced("reduced_tet", (3, 2, 2), 12, 30)
    golden_reduced("reduced_tri", (5, 4), 8, 31)
    golden_contact("contact_tet", (3, 3, 2), 50)
    golden_contact("contact_tri", (6, 5), 51)
    golden_quadratic("quadratic_tet", (3, 3, 2), 60)
    golden_quadratic("quadratic_tri", (6, 5), 61)
    golden_mfem("mfem_tri", (4, 2), 10.0, 40)
    golden_mfem("mfem_tet", (2, 2, 1), 10.0, 41)
