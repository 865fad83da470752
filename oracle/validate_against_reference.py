"""Pin the oracle to the reference itself (run in the build container only).

Imports otmanon/simkit read-only from ``/root/reference`` and compares every
oracle function with the reference function it restates, on seeded inputs that
include inverted elements, heterogeneous materials, 2D and 3D.  Prints one line
per check and exits non-zero on any mismatch.  ``/root/reference`` does not
exist on the GPU box; nothing in tests/bench imports this module.
"""

import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, "/root/reference")
sys.path.insert(0, __file__.rsplit("/", 2)[0])

import simkit  # noqa: E402
import simkit.energies as ske  # noqa: E402
from simkit.solvers import newton_solver as ref_newton  # noqa: E402
from simkit.integrators import backward_euler as ref_be, bdf2 as ref_bdf2  # noqa: E402
from simkit.fast_sandwich_transform_clustered import fast_sandwich_transform_clustered as ref_fst  # noqa: E402
from simkit.rotation_gradient import rotation_gradient_F as ref_rotgrad  # noqa: E402
from simkit.svd_rv import svd_rv as ref_svd_rv  # noqa: E402

from oracle import elasticity as oe  # noqa: E402
from simkit_b200 import synthetic as syn  # noqa: E402

FAIL = []


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / den


def check(name, a, b, tol=1e-12):
    r = rel(a, b)
    ok = r <= tol
    print(f"{'ok ' if ok else 'BAD'} {name:60s} rel={r:.3e}")
    if not ok:
        FAIL.append(name)


REF = {
    "stable_neo_hookean": ("stable_neo_hookean", True),
    "neo_hookean": ("neo_hookean", True),
    "arap": ("arap", False),
    "stvk": ("stvk", True),
    "linear_elasticity": ("linear_elasticity", True),
    "fcr": ("fcr", True),
    "macklin_mueller_neo_hookean": ("macklin_mueller_neo_hookean", True),
}


def ref_call(material, kind, tier, *args, **kw):
    stem, _ = REF[material]
    fn = getattr(ske, f"{stem}_{kind}_{tier}" if tier else f"{stem}_{kind}")
    return fn(*args, **kw)


def main():
    rng = np.random.default_rng(0)
    for dim, cells in ((3, (4, 3, 5)), (2, (7, 6))):
        X, T = syn.make_mesh(cells)
        n, t = X.shape[0], T.shape[0]
        ext = tuple(1.0 for _ in cells)
        J_ref = simkit.deformation_jacobian(X, T)
        J = oe.deformation_jacobian(X, T)
        check(f"[{dim}D] deformation_jacobian", J.toarray(), J_ref.toarray(), 1e-13)
        print("    nnz ours/ref", J.nnz, J_ref.nnz)
        vol_ref = simkit.volume(X, T)
        vol = oe.volume(X, T)
        check(f"[{dim}D] volume", vol, vol_ref, 1e-14)
        M_ref = simkit.massmatrix(X, T, 1e3)
        check(f"[{dim}D] massmatrix", oe.massmatrix(X, T, 1e3).diagonal(), M_ref.diagonal(), 1e-14)
        check(f"[{dim}D] gravity", oe.gravity_force(X, T, -9.8, 1e3), simkit.gravity_force(X, T, -9.8, 1e3), 1e-14)
        mu_h, lam_h = syn.heterogeneous_lame(t)
        for sigma in (0.1, 0.4):
            U = syn.jittered_state(X, cells, ext, sigma=sigma)
            F = np.asarray(J_ref @ U.reshape(-1, 1)).reshape(-1, dim, dim)
            ninv = int((np.linalg.det(F) <= 0).sum())
            print(f"  sigma={sigma}: inverted elements {ninv}/{t}")
            Ur, Sr, Vr = ref_svd_rv(F)
            Uo, So, Vo = oe.svd_rv(F)
            check(f"[{dim}D s={sigma}] svd_rv U", Uo, Ur, 1e-13)
            check(f"[{dim}D s={sigma}] svd_rv S", So, Sr, 1e-13)
            check(f"[{dim}D s={sigma}] svd_rv V", Vo, Vr, 1e-13)
            Rr, SSr = simkit.polar_svd(F)
            Ro, SSo = oe.polar_svd(F)
            check(f"[{dim}D s={sigma}] polar R", Ro, Rr, 1e-13)
            check(f"[{dim}D s={sigma}] polar S", SSo, SSr, 1e-13)
            check(f"[{dim}D s={sigma}] rotation_gradient_F", oe.rotation_gradient_F(F), ref_rotgrad(F), 1e-12)
            for material in oe.MATERIALS:
                if material == "neo_hookean" and ninv > 0:
                    continue  # NaNs by construction (neo_hookean.py:21-25)
                has_lam = REF[material][1]
                margs = (mu_h, lam_h) if has_lam else (mu_h,)
                tag = f"[{dim}D s={sigma}] {material}"
                check(tag + " energy_element_F", oe.energy_element_F(material, F, mu_h, lam_h),
                      ref_call(material, "energy", "element_F", F, *margs))
                check(tag + " gradient_element_F", oe.gradient_element_F(material, F, mu_h, lam_h),
                      ref_call(material, "gradient", "element_F", F, *margs))
                check(tag + " hessian_element_F", oe.hessian_element_F(material, F, mu_h, lam_h),
                      ref_call(material, "hessian", "element_F", F, *margs), 1e-11)
                E_ref = ref_call(material, "energy", "x", U, J_ref, *margs, vol_ref)
                check(tag + " energy_x", oe.energy_x(material, U, J, mu_h, lam_h, vol), E_ref, 1e-13)
                g_ref = ref_call(material, "gradient", "x", U, J_ref, *margs, vol_ref)
                check(tag + " gradient_x", oe.gradient_x(material, U, J, mu_h, lam_h, vol), g_ref)
                for psd in (True, False):
                    Q_ref = ref_call(material, "hessian", "x", U, J_ref, *margs, vol_ref, psd=psd)
                    Q = oe.hessian_x(material, U, J, mu_h, lam_h, vol, psd=psd)
                    check(tag + f" hessian_x psd={psd}", Q.toarray(), Q_ref.toarray(), 1e-11)
                # _u tier
                xb = X + 0.01 * rng.standard_normal(X.shape)
                u = U - xb
                Jxb = J_ref @ xb.reshape(-1, 1)
                Q_ref = ref_call(material, "hessian", "u", u, J_ref, Jxb, *margs, vol_ref)
                Q = oe.hessian_x(material, u, J, mu_h, lam_h, vol, Jx_bar=Jxb)
                check(tag + " hessian_u", Q.toarray(), Q_ref.toarray(), 1e-11)
        # psd_project on arbitrary symmetric input
        A = rng.standard_normal((50, dim * dim, dim * dim))
        A = A + np.swapaxes(A, 1, 2)
        check(f"[{dim}D] psd_project proj", oe.psd_project(A), simkit.psd_project(A), 1e-12)
        check(f"[{dim}D] psd_project abs", oe.psd_project(A, "abs"), simkit.psd_project(A, "abs"), 1e-12)

        # elastic dispatcher: psd before vol (arap, linear-elasticity routed)
        U = syn.jittered_state(X, cells, ext, sigma=0.4)
        for material, name in (("arap", "arap"), ("linear_elasticity", "linear-elasticity"), ("fcr", "fcr"),
                               ("macklin_mueller_neo_hookean", "macklin-mueller-neo-hookean")):
            Q_ref = ske.elastic_hessian_x(U, J_ref, mu_h, lam_h, vol_ref, name, psd=True)
            Q = oe.hessian_x(material, U, J, mu_h, lam_h, vol, psd=True, psd_before_vol=True)
            check(f"[{dim}D] elastic_hessian_x {name}", Q.toarray(), Q_ref.toarray(), 1e-11)

        # MFEM blocks (stretch.py, stretch_gradient.py, symmetric_stretch_map.py, the _S tier of the dispatcher)
        from simkit.stretch_gradient import stretch_gradient_dF as ref_dSdF, stretch_gradient_dz as ref_dsdz
        from simkit.symmetric_stretch_map import symmetric_stretch_map as ref_ssm
        Fm = np.asarray(J_ref @ U.reshape(-1, 1)).reshape(-1, dim, dim)
        check(f"[{dim}D] stretch", oe.stretch(Fm), simkit.stretch(Fm), 1e-13)
        check(f"[{dim}D] stretch_gradient_dF", oe.stretch_gradient_dF(Fm), ref_dSdF(Fm), 1e-11)
        Se_r, Sei_r = ref_ssm(t, dim)
        Se_o, Sei_o = oe.symmetric_stretch_map(t, dim)
        check(f"[{dim}D] symmetric_stretch_map", np.r_[abs(Se_o - Se_r).max(), abs(Sei_o - Sei_r).max()] + 1.0, np.ones(2), 0.0)
        dz_r = ref_dsdz(U, sps.csc_matrix(J_ref), dim, Ci=Sei_r)
        dz_o = oe.stretch_gradient_dz(U, sps.csc_matrix(J), dim, Ci=Sei_o)
        check(f"[{dim}D] stretch_gradient_dz", dz_o.toarray(), dz_r.toarray(), 1e-11)
        _, Sm = simkit.polar_svd(Fm)
        Sc = Sm.reshape(t, dim * dim) @ np.asarray(ref_ssm(1, dim)[1].todense()).T
        mm = "macklin-mueller-neo-hookean"
        check(f"[{dim}D] elastic_energy_S (MM)", oe.elastic_S("energy", Sc, mu_h, lam_h, vol, mm), ske.elastic_energy_S(Sc, mu_h, lam_h, vol_ref, mm), 1e-13)
        check(f"[{dim}D] elastic_gradient_S (MM)", oe.elastic_S("gradient", Sc, mu_h, lam_h, vol, mm), ske.elastic_gradient_S(Sc, mu_h, lam_h, vol_ref, mm))
        check(f"[{dim}D] elastic_hessian_S (MM)", oe.elastic_S("hessian", Sc, mu_h, lam_h, vol, mm), ske.elastic_hessian_S(Sc, mu_h, lam_h, vol_ref, mm), 1e-11)

        # structural pattern is a superset of the canonicalised reference pattern
        mu0, lam0 = syn.lame()
        Q_ref = oe.canonical_csr(ske.stable_neo_hookean_hessian_x(U, J_ref, mu0, lam0, vol_ref))
        indptr, indices, bptr, bcol = oe.structural_pattern(T, n, dim)
        same = Q_ref.nnz == indices.shape[0] and np.array_equal(Q_ref.indptr, indptr) and np.array_equal(Q_ref.indices, indices)
        print(f"{'ok ' if same else 'BAD'} [{dim}D] structural pattern == canonical reference pattern (sNH)")
        if not same:
            FAIL.append("pattern")

        # Newton / integrators on the small mesh
        rho, h = 1e3, 1e-2
        Mv = sps.kron(simkit.massmatrix(X, T, rho), sps.identity(dim)).tocsc()
        fg = simkit.gravity_force(X, T, -9.8, rho).reshape(-1, 1)

        def mk(mod_e, mod_g, mod_h, Jm, volm):
            def E(x):
                return mod_e(x.reshape(-1, dim), Jm, mu0, lam0, volm) - float((fg.T @ x).item())

            def G(x):
                return mod_g(x.reshape(-1, dim), Jm, mu0, lam0, volm) - fg

            def H(x):
                return mod_h(x.reshape(-1, dim), Jm, mu0, lam0, volm)
            return E, G, H

        Er, Gr, Hr = mk(ske.stable_neo_hookean_energy_x, ske.stable_neo_hookean_gradient_x,
                        ske.stable_neo_hookean_hessian_x, J_ref, vol_ref)

        def oe_e(x, Jm, mu, lam, volm):
            return oe.energy_x("stable_neo_hookean", x, Jm, mu, lam, volm)

        def oe_g(x, Jm, mu, lam, volm):
            return oe.gradient_x("stable_neo_hookean", x, Jm, mu, lam, volm)

        def oe_h(x, Jm, mu, lam, volm):
            return oe.hessian_x("stable_neo_hookean", x, Jm, mu, lam, volm)

        Eo, Go, Ho = mk(oe_e, oe_g, oe_h, J, vol)
        x0 = U.reshape(-1, 1)
        x1 = X.reshape(-1, 1)
        xr, info_r = ref_be(x0, x1, Er, Gr, Hr, Mv, h, max_iter=3, return_info=True)
        xo, info_o = oe.backward_euler(x0, x1, Eo, Go, Ho, Mv, h, max_iter=3, return_info=True)
        check(f"[{dim}D] backward_euler iterate", xo, xr, 1e-10)
        assert info_r["alphas"] == info_o["alphas"] and info_r["iters"] == info_o["iters"], (info_r["alphas"], info_o["alphas"])
        x2 = x1 + 1e-3 * rng.standard_normal(x1.shape)
        x3 = x1 + 1e-3 * rng.standard_normal(x1.shape)
        xr = ref_bdf2(x0, x1, x2, x3, Er, Gr, Hr, Mv, h, max_iter=2)
        xo = oe.bdf2(x0, x1, x2, x3, Eo, Go, Ho, Mv, h, max_iter=2)
        check(f"[{dim}D] bdf2 iterate", xo, xr, 1e-10)
        # plain Newton on the (regularised) static problem
        Kreg = Mv / h ** 2
        xs = syn.jittered_state(X, cells, ext, sigma=0.1).reshape(-1, 1)

        def reg(E, G, H):
            return (lambda x: E(x) + 0.5 * float(((x - x1).T @ Kreg @ (x - x1)).item()),
                    lambda x: G(x) + Kreg @ (x - x1), lambda x: H(x) + Kreg)

        xr, ir = ref_newton(xs, *reg(Er, Gr, Hr), max_iter=4, return_info=True)
        xo, io = oe.newton_solver(xs, *reg(Eo, Go, Ho), max_iter=4, return_info=True)
        check(f"[{dim}D] newton iterate (alphas {ir['alphas']})", xo, xr, 1e-10)
        # block-Jacobi CG substitute agrees with the direct solve
        Hm = Hr(x0) + Mv / h ** 2
        rhs = -Gr(x0)
        import scipy.sparse.linalg as spla
        d_ref = spla.spsolve(Hm.tocsc(), rhs)
        d_cg, its = oe.block_jacobi_cg(Hm, rhs, dim)
        check(f"[{dim}D] block_jacobi_cg vs spsolve ({its} its)", d_cg, d_ref, 1e-9)

        # FST
        m1, m2, nc = 5, 4, 3
        A = sps.random(m1, dim * dim * t, density=0.3, random_state=1, format="csr").toarray()
        B = sps.random(dim * dim * t, m2, density=0.3, random_state=2, format="csr")
        l = rng.integers(0, nc, size=t)
        l[:nc] = np.arange(nc)
        fr = ref_fst(A, B, l, dim=dim)
        ARBs = oe.fst_precompute(A, B, l, dim=dim)
        check(f"[{dim}D] fst ARBs", ARBs, fr.ARBs, 1e-12)
        r = rng.standard_normal((nc, dim, dim))
        check(f"[{dim}D] fst eval", oe.fst_eval(ARBs, r, dim), fr(r), 1e-12)

        # quadratic term and Dirichlet penalty
        from simkit.energies.quadratic import quadratic_energy as rqe, quadratic_gradient as rqg
        from simkit.dirichlet_penalty import dirichlet_penalty as rdp
        nvq = X.shape[0]
        bI = np.sort(rng.choice(nvq, size=5, replace=False))
        yq = X[bI] + 0.05 * rng.standard_normal((5, dim))
        for gname, gam in (("scalar", 1e6), ("per vertex", 1e5 * (1.0 + rng.random(5)))):
            Qr, br = rdp(bI, yq, nvq, gam)
            Qo, bo = oe.dirichlet_penalty(bI, yq, nvq, gam)
            check(f"[{dim}D] dirichlet_penalty Q ({gname})", Qo.toarray(), Qr.toarray(), 1e-15)
            check(f"[{dim}D] dirichlet_penalty b ({gname})", bo, br, 1e-15)
        mul = 1.0 + rng.random((T.shape[0], 1))
        check(f"[{dim}D] dirichlet_laplacian", oe.dirichlet_laplacian(X, T, mul).toarray(), simkit.dirichlet_laplacian(X, T, mul).toarray(), 1e-13)
        check(f"[{dim}D] dirichlet_laplacian (vector, scalar mu)", oe.dirichlet_laplacian(X, T, 2.5, vector=True).toarray(),
              simkit.dirichlet_laplacian(X, T, 2.5, vector=True).toarray(), 1e-13)
        Ls = sps.random(nvq * dim, nvq * dim, density=0.02, random_state=3, format="csr")
        Qs = (Ls + Ls.T + Qr).tocsr()
        xq = rng.standard_normal((nvq * dim, 1))
        check(f"[{dim}D] quadratic_energy", oe.quadratic_energy(xq, Qs, br), rqe(xq, Qs, br), 1e-13)
        check(f"[{dim}D] quadratic_gradient", oe.quadratic_gradient(xq, Qs, br), rqg(xq, Qs, br), 1e-13)

        # subspace construction helpers
        from simkit.orthonormalize import orthonormalize as ref_on
        from simkit.project_into_subspace import project_into_subspace as ref_pis
        nd = X.shape[0] * dim
        Bs = rng.standard_normal((nd, 7))
        Bs[:, 5] = Bs[:, 1] * 2.0                     # a dependent column: dropped by the threshold test
        Md = sps.diags(1.0 + rng.random(nd))
        check(f"[{dim}D] orthonormalize (mass)", oe.orthonormalize(Bs, Md, 1e-10), ref_on(Bs, Md, 1e-10), 1e-12)
        check(f"[{dim}D] orthonormalize (identity)", oe.orthonormalize(Bs[:, :5]), ref_on(Bs[:, :5]), 1e-12)
        from simkit.spectral_cubature import spectral_cubature as ref_sc
        from simkit.average_onto_simplex import average_onto_simplex as ref_avg
        Wm = rng.standard_normal((X.shape[0], 6)) * np.array([3.0, 2.0, 1.5, 1.0, 0.7, 0.5])
        check(f"[{dim}D] average_onto_simplex", oe.average_onto_simplex(Wm, T), ref_avg(Wm, T), 1e-15)
        rl = ref_sc(X, T, Wm, 5, return_labels=True, return_centroids=True)
        ol = oe.spectral_cubature(X, T, Wm, 5)
        for nm, a, b in zip(("lI", "mc", "labels", "centroids"), ol, rl):
            check(f"[{dim}D] spectral_cubature {nm}", np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), 1e-13)
        from simkit.lbs_jacobian import lbs_jacobian as ref_lbs
        Ww = rng.standard_normal((X.shape[0], 4))
        check(f"[{dim}D] lbs_jacobian", oe.lbs_jacobian(X, Ww), ref_lbs(X, Ww), 1e-15)
        yv = rng.standard_normal((nd, 1))
        check(f"[{dim}D] project_into_subspace (mass)", oe.project_into_subspace(yv, Bs[:, :5], Md), ref_pis(yv, Bs[:, :5], Md), 1e-12)
        check(f"[{dim}D] project_into_subspace (identity)", oe.project_into_subspace(yv, Bs[:, :5]), ref_pis(yv, Bs[:, :5]), 1e-12)

    print("FAILED: " + ", ".join(FAIL) if FAIL else "ALL OK")
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
