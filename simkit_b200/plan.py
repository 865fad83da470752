"""MeshPlan: Python handle on the per-mesh device plan (``skb_plan`` in include/simkit_b200.h).

A plan is what the reference rebuilds implicitly on every call: the operator
``J`` (deformation_jacobian.py:9-87), the sparsity of ``J^T H J`` and the scatter
order.  Here it is built once on the device and cached on the ``J`` object the
drop-in ``deformation_jacobian`` returns (or recovered from a plain scipy ``J``).
"""

import ctypes
import weakref

import numpy as np
import scipy.sparse as sps

from . import _lib
from ._lib import MATERIAL_IDS, check, f64, material_arg, ptr


class MeshPlan:
    def __init__(self, X=None, T=None, dim=None, n=None, D=None, device=0, tile_elems=0, t_active=None, t_energy=None):
        lib = _lib.load()
        T = np.ascontiguousarray(T)
        if T.dtype not in (np.int64, np.int32):
            T = T.astype(np.int64)
        self.t = int(T.shape[0])
        handle = ctypes.c_void_p()
        if X is not None:
            X = f64(X)
            self.n, self.dim = int(X.shape[0]), int(X.shape[1])
            if T.shape[1] != self.dim + 1:
                raise ValueError("Only dim == 2 or 3 simplices (dim+1 corners) are supported")
            if (t_active is not None and t_active != self.t) or t_energy is not None:
                # one rank of a sharded mesh: trailing elements only reserve pattern slots (t_active < t), or are the
                # lower neighbour's interface elements evaluated here as well (t_energy < t_active = t)
                t_total = self.t
                self.t = t_total if t_active is None else int(t_active)
                check(lib.skb_plan_create_sharded(ptr(X), ptr(T), T.dtype.itemsize, self.n, self.t, t_total,
                                                  0 if t_energy is None else int(t_energy), self.dim,
                                                  device, tile_elems, ctypes.byref(handle)))
            else:
                check(lib.skb_plan_create(ptr(X), ptr(T), T.dtype.itemsize, self.n, self.t, self.dim, device,
                                          tile_elems, ctypes.byref(handle)))
        else:
            D = f64(D)
            self.n, self.dim = int(n), int(dim)
            check(lib.skb_plan_create_from_operator(ptr(T), ptr(D), T.dtype.itemsize, self.n, self.t, self.dim,
                                                    device, tile_elems, ctypes.byref(handle)))
        self._lib = lib
        self._h = handle
        self.device = int(device)
        self.n_agg = 0
        # rest positions kept for the automatic two-level preconditioner of solves on lazy Hessians (device_csr.py)
        self._X_rest = None if X is None or t_active is not None or t_energy is not None else X
        self.T = T
        info = np.zeros(8, dtype=np.int64)
        check(lib.skb_plan_info(self._h, ptr(info)))
        self.nnzb, self.nnz, self.n_tiles = int(info[3]), int(info[4]), int(info[5])
        self.n_block_partials, self.n_vertex_partials = int(info[6]), int(info[7])
        self.ndof = self.n * self.dim
        self._pattern = None
        self._finalizer = weakref.finalize(self, lib.skb_plan_destroy, handle)

    # ---------------------------------------------------------------- pattern
    def csr_pattern(self):
        """Canonical ``(indptr, indices)`` int32, sorted."""
        if self._pattern is None:
            indptr = np.empty(self.ndof + 1, dtype=np.int32)
            indices = np.empty(self.nnz, dtype=np.int32)
            check(self._lib.skb_plan_csr_pattern(self._h, ptr(indptr), ptr(indices)))
            self._pattern = (indptr, indices)
        return self._pattern

    def block_pattern(self):
        bptr = np.empty(self.n + 1, dtype=np.int32)
        bcol = np.empty(self.nnzb, dtype=np.int32)
        check(self._lib.skb_plan_block_pattern(self._h, ptr(bptr), ptr(bcol)))
        return bptr, bcol

    def slot_map(self):
        K = self.dim + 1
        slot = np.empty((self.t, K, self.dim, K, self.dim), dtype=np.int32)
        check(self._lib.skb_plan_slot_map(self._h, ptr(slot)))
        return slot

    def element_D(self):
        D = np.empty((self.t, self.dim, self.dim + 1))
        check(self._lib.skb_plan_element_D(self._h, ptr(D)))
        return D

    def element_order(self):
        """``order[i]`` = index (in ``T``) of the element the plan lists ``i``-th (its internal spatial order)."""
        order = np.empty(self.t, dtype=np.int32)
        check(self._lib.skb_plan_element_order(self._h, ptr(order)))
        return order

    def volume(self):
        vol = np.empty((self.t, 1))
        check(self._lib.skb_plan_volume(self._h, ptr(vol)))
        return vol

    def vertex_masses(self, rho=1.0):
        rho_a, rho_n = material_arg(rho, self.t, "rho")
        m = np.empty(self.n)
        check(self._lib.skb_plan_vertex_masses(self._h, ptr(rho_a), rho_n, ptr(m)))
        return m

    def csr_matrix(self, vals):
        indptr, indices = self.csr_pattern()
        return sps.csr_matrix((vals, indices, indptr), shape=(self.ndof, self.ndof))

    # ------------------------------------------------------------- evaluation
    def _mats(self, mu, lam, vol):
        # every call that passes materials overwrites the plan's device copies: an ElasticPotential's cached upload
        # (set_materials(owner=...)) is no longer what sits there
        self._mat_owner = None
        mu_a, mu_n = material_arg(mu, self.t, "mu")
        lam_a, lam_n = material_arg(lam, self.t, "lam")
        vol_a, vol_n = material_arg(vol, self.t, "vol")
        return (mu_a, lam_a, vol_a), (ptr(mu_a), mu_n, ptr(lam_a), lam_n, ptr(vol_a), vol_n)

    def _x(self, x):
        x = f64(x).reshape(-1)
        if x.size != self.ndof:
            raise ValueError("x has %d entries, the mesh has %d dofs" % (x.size, self.ndof))
        return x

    def _fbar(self, Fbar):
        if Fbar is None:
            return None
        Fbar = f64(Fbar).reshape(-1)
        if Fbar.size != self.t * self.dim * self.dim:
            raise ValueError("Jx_bar has the wrong size")
        return Fbar

    def energy(self, material, x, mu, lam, vol, Fbar=None):
        x = self._x(x)
        Fbar = self._fbar(Fbar)
        keep, margs = self._mats(mu, lam, vol)
        out = ctypes.c_double(0.0)
        check(self._lib.skb_energy(self._h, MATERIAL_IDS[material], ptr(x), ptr(Fbar), *margs, ctypes.byref(out)))
        return float(out.value)

    def gradient(self, material, x, mu, lam, vol, Fbar=None):
        x = self._x(x)
        Fbar = self._fbar(Fbar)
        keep, margs = self._mats(mu, lam, vol)
        g = np.empty((self.ndof, 1))
        check(self._lib.skb_gradient(self._h, MATERIAL_IDS[material], ptr(x), ptr(Fbar), *margs, ptr(g)))
        return g

    def hessian_values(self, material, x, mu, lam, vol, psd_mode, Fbar=None, out=None):
        x = self._x(x)
        Fbar = self._fbar(Fbar)
        keep, margs = self._mats(mu, lam, vol)
        vals = np.empty(self.nnz) if out is None else out
        check(self._lib.skb_hessian(self._h, MATERIAL_IDS[material], int(psd_mode), ptr(x), ptr(Fbar), *margs, ptr(vals)))
        return vals

    def hessian(self, material, x, mu, lam, vol, psd_mode, Fbar=None):
        """The assembled Hessian as a :class:`simkit_b200.device_csr.DeviceCSR`: a ``scipy.sparse.csr_matrix`` whose
        values stay in HBM until something on the host touches them (SKB_LAZY_HESSIAN=0: plain host matrix)."""
        import os
        if os.environ.get("SKB_LAZY_HESSIAN", "1") == "0":
            return self.csr_matrix(self.hessian_values(material, x, mu, lam, vol, psd_mode, Fbar))
        from .device_csr import DeviceCSR, _Buffer
        x = self._x(x)
        Fbar = self._fbar(Fbar)
        keep, margs = self._mats(mu, lam, vol)
        buf = _Buffer(self.device, self.nnz)
        check(self._lib.skb_gradient_hessian_resident(self._h, MATERIAL_IDS[material], int(psd_mode), ptr(x), ptr(Fbar),
                                                      *margs, None, buf.ptr))
        return DeviceCSR(_plan=self, _buf=buf)

    def value_positions_of(self, S):
        """``(pos, vals)`` of the stored entries of the scipy sparse ``S`` inside this mesh's CSR values, or
        ``(None, None)`` if ``S`` has a non-zero outside the pattern (``device_csr.value_positions_of``)."""
        from .device_csr import value_positions_of
        return value_positions_of(self, S)

    def gradient_hessian(self, material, x, mu, lam, vol, psd_mode, Fbar=None, g_out=None, vals_out=None):
        x = self._x(x)
        Fbar = self._fbar(Fbar)
        keep, margs = self._mats(mu, lam, vol)
        g = np.empty((self.ndof, 1)) if g_out is None else g_out
        vals = np.empty(self.nnz) if vals_out is None else vals_out
        check(self._lib.skb_gradient_hessian(self._h, MATERIAL_IDS[material], int(psd_mode), ptr(x), ptr(Fbar),
                                             *margs, ptr(g), ptr(vals)))
        return g, vals

    def set_materials(self, mu, lam, vol, owner=None):
        """Uploads the materials the device-resident step uses.  ``owner``: a token the caller can compare with
        ``plan._mat_owner`` later to learn that its upload is still the one on the device (any other call that passes
        materials resets it)."""
        keep, margs = self._mats(mu, lam, vol)
        check(self._lib.skb_set_materials(self._h, *margs))
        self._mat_owner = owner

    def last_launch_count(self):
        return int(self._lib.skb_last_launch_count(self._h))

    # ------------------------------------------------------------ linear solve
    def pcg(self, vals, rhs, diag_add=None, rtol=1e-10, max_iter=10000):
        vals = f64(vals).reshape(-1)
        rhs = f64(rhs).reshape(-1)
        dadd = None if diag_add is None else f64(diag_add).reshape(-1)
        x = np.empty(self.ndof)
        iters = ctypes.c_int(0)
        relres = ctypes.c_double(0.0)
        check(self._lib.skb_pcg(self._h, ptr(vals), ptr(dadd), ptr(rhs), float(rtol), int(max_iter), ptr(x),
                                ctypes.byref(iters), ctypes.byref(relres)))
        return x, int(iters.value), float(relres.value)

    def newton(self, material, x0, psd_mode=1, x_tilde=None, mass=None, kin_scale=0.0, f_ext=None, pin_k=None,
               pin_target=None, max_iter=1, do_line_search=True, tolerance=1e-6, ls_alpha=0.01, ls_beta=0.5,
               ls_max_iter=100, ls_threshold=1e-12, pcg_rtol=1e-10, pcg_max_iter=20000):
        opts = _lib.NewtonOpts(MATERIAL_IDS[material], int(psd_mode), int(max_iter), int(bool(do_line_search)),
                               float(tolerance), float(ls_alpha), float(ls_beta), int(ls_max_iter),
                               float(ls_threshold), float(pcg_rtol), int(pcg_max_iter))
        info = _lib.NewtonInfo()
        x0 = self._x(x0)
        opt = [None if a is None else self._x(a) for a in (x_tilde, mass, f_ext, pin_k, pin_target)]
        out = np.empty((self.ndof, 1))
        check(self._lib.skb_newton(self._h, ctypes.byref(opts), ptr(x0), ptr(opt[0]), ptr(opt[1]), float(kin_scale),
                                   ptr(opt[2]), ptr(opt[3]), ptr(opt[4]), ptr(out), ctypes.byref(info)))
        if not np.isfinite(info.last_pcg_relres) or not np.isfinite(info.last_step_norm):
            raise _lib.SimkitB200Error(
                "Newton step failed: the linear solve produced a non-finite residual (relres %r). The Newton system must be "
                "symmetric positive definite: project the Hessian (psd=True) or add inertia / penalty terms; NaN in the "
                "state (inverted elements under neo_hookean) propagates as in the reference." % (info.last_pcg_relres,))
        n_it = info.iters + 1
        # The reference's ``info`` (solvers/newton.py:43-72) holds per-iteration ``g`` / ``dx`` copies (2 x 66 MB per
        # iteration at 16 M tets); the device-resident step keeps them in HBM, so the keys exist but are empty lists.
        # ``alphas`` holds the first 64 step lengths (NewtonInfo's fixed array); ``pcg_*`` / ``step_norm`` are extras.
        return out, dict(iters=info.iters, alphas=[info.alphas[i] for i in range(min(n_it, 64))], g=[], dx=[],
                         pcg_iters=info.pcg_iters_total, pcg_relres=info.last_pcg_relres,
                         step_norm=info.last_step_norm)

    def set_contact_plane(self, k=0.0, p=None, n=None, weights=None):
        """Plane contact springs inside the device-resident Newton step (``energies/contact_springs_plane.py``):
        ``k/2 sum_{v under the plane} m_v (n.(x_v - p))^2`` is added to the energy of the line search, the gradient and
        (as ``k m_v n n^T`` in the diagonal blocks) the Hessian of every following ``newton``.  ``k = 0`` removes it."""
        if not k or p is None or n is None:
            check(self._lib.skb_newton_set_contact_plane(self._h, 0.0, None, None, None))
            return
        p = f64(np.asarray(p, dtype=np.float64).reshape(-1))
        n = f64(np.asarray(n, dtype=np.float64).reshape(-1))
        if p.size != self.dim or n.size != self.dim:
            raise ValueError("p and n must have dim entries")
        w = None if weights is None else f64(np.asarray(weights, dtype=np.float64).reshape(-1))
        if w is not None and w.size != self.n:
            raise ValueError("weights must have one entry per vertex")
        check(self._lib.skb_newton_set_contact_plane(self._h, float(k), ptr(p), ptr(n), ptr(w)))

    def set_contact_sphere(self, k=0.0, p=None, r=0.0, weights=None):
        """Sphere contact springs inside the device-resident Newton step (``energies/contact_springs_sphere.py``);
        ``k = 0`` removes them.  May be combined with ``set_contact_plane``."""
        if not k or p is None:
            check(self._lib.skb_newton_set_contact_sphere(self._h, 0.0, None, 0.0, None))
            return
        p = f64(np.asarray(p, dtype=np.float64).reshape(-1))
        if p.size != self.dim:
            raise ValueError("p must have dim entries")
        w = None if weights is None else f64(np.asarray(weights, dtype=np.float64).reshape(-1))
        if w is not None and w.size != self.n:
            raise ValueError("weights must have one entry per vertex")
        check(self._lib.skb_newton_set_contact_sphere(self._h, float(k), ptr(p), float(r), ptr(w)))

    def set_quadratic(self, Q=None, b=None):
        """General sparse quadratic term ``1/2 x^T Q x + b^T x`` (``energies/quadratic.py``) inside the device-resident
        Newton step: energy in the line search, gradient ``Q x + b``, Hessian ``Q`` added into the CSR values on the
        device.  ``Q``: symmetric ``(n*dim, n*dim)`` scipy sparse matrix or ndarray whose entries lie inside the mesh's
        CSR pattern (``ValueError`` otherwise) -- a ``dirichlet_penalty`` matrix, a mass or Laplacian regulariser;
        ``Q = None`` removes the term."""
        self._quad_owner = None   # ElasticPotential's upload cache: any direct call invalidates it
        self._quad_active = False
        if Q is None:
            check(self._lib.skb_newton_set_quadratic(self._h, None, None, None, None))
            return
        from .energies.quadratic import _csr
        indptr, indices, vals = _csr(Q, self.ndof)
        bb = None if b is None else f64(np.asarray(b, dtype=np.float64).reshape(-1))
        if bb is not None and bb.size != self.ndof:
            raise ValueError("b must have n*dim entries")
        check(self._lib.skb_newton_set_quadratic(self._h, ptr(indptr), ptr(indices), ptr(vals), ptr(bb)))
        self._quad_active = True

    COARSE_MIN_ITERS = 300   # block-Jacobi iteration count above which the coarse correction pays for itself

    def auto_aggregates(self):
        """Default size of the coarse space: about one aggregate per 1,000 vertices (150 in 2D, where an aggregate
        has 3 coarse unknowns instead of 6), 8 ... 729.  The dense inverse of the coarse matrix costs
        O((6 n_agg)^3) per Newton iteration (20 ms at 729 aggregates in 3D) and every CG iteration three more
        kernels, so the correction only pays when block-Jacobi needs more than ``COARSE_MIN_ITERS`` iterations
        (measured: 709 -> 303 at 16 M tets, 1635 -> 267 on a 200 k-triangle sheet; not on a 1 M-tet beam with 179)."""
        return int(min(729, max(8, self.n // (1000 if self.dim == 3 else 150))))

    def set_coarse_space(self, X, n_agg_target=729):
        """Two-level PCG preconditioner (``csrc/coarse.cuh``): vertices are binned by position into about
        ``n_agg_target`` box-shaped aggregates (at most 2048) whose rigid-body modes form the coarse space.
        ``X`` are the rest positions ``(n, dim)``.  ``n_agg_target = 0`` (or ``X=None``) removes it.  The reference
        solves the Newton system directly (solvers/newton.py:52); this only changes the CG iteration count."""
        if X is None or not n_agg_target:
            check(self._lib.skb_pcg_set_coarse(self._h, 0, None, None))
            self.n_agg = 0
            return 0
        X = f64(X).reshape(self.n, self.dim)
        lo, hi = X.min(axis=0), X.max(axis=0)
        ext = np.maximum(hi - lo, 1e-300)
        # boxes of equal edge length: bins per axis proportional to the extent
        edge = (np.prod(ext) / float(min(int(n_agg_target), 2048))) ** (1.0 / self.dim)
        nb = np.maximum(1, np.floor(ext / edge + 0.5).astype(np.int64))
        while int(np.prod(nb)) > 2048:
            nb[np.argmax(nb)] -= 1
        ib = np.minimum((np.floor((X - lo) / ext * nb)).astype(np.int64), nb - 1)
        flat = ib[:, 0]
        for a in range(1, self.dim):
            flat = flat * nb[a] + ib[:, a]
        used, agg = np.unique(flat, return_inverse=True)          # drop empty boxes
        n_agg = int(used.size)
        cnt = np.bincount(agg, minlength=n_agg).astype(np.float64)
        cen = np.stack([np.bincount(agg, weights=X[:, a], minlength=n_agg) / cnt for a in range(self.dim)], axis=1)
        xrel = np.ascontiguousarray(X - cen[agg])
        agg32 = np.ascontiguousarray(agg.astype(np.int32))
        check(self._lib.skb_pcg_set_coarse(self._h, n_agg, ptr(agg32), ptr(xrel)))
        self.n_agg = n_agg
        return n_agg

    def set_basis(self, B):
        """Keeps the subspace basis ``B (n*dim, r)`` resident on the device (``None`` releases it);
        ``reduced(..., B=None)`` then skips the upload (3.4 GB at BASELINE config 4)."""
        if B is None:
            check(self._lib.skb_plan_set_basis(self._h, 0, None))
            self._basis_r = 0
            return
        B = f64(B)
        if B.ndim != 2 or B.shape[0] != self.ndof:
            raise ValueError("B must be (n*dim, r)")
        check(self._lib.skb_plan_set_basis(self._h, B.shape[1], ptr(B)))
        self._basis_r = B.shape[1]

    def reduced(self, material, B, z, x0=None, psd_mode=1, want=("E", "g", "H")):
        z = f64(z).reshape(-1)
        if B is None:
            r = getattr(self, "_basis_r", 0)
            if r == 0:
                raise ValueError("no resident basis: call set_basis(B) first or pass B")
        else:
            B = f64(B)
            r = B.shape[1]
        if z.size != r:
            raise ValueError("reduced coordinates do not match the basis")
        x0 = None if x0 is None else self._x(x0)
        E = ctypes.c_double(0.0)
        g = np.empty((r, 1)) if "g" in want else None
        H = np.empty((r, r)) if "H" in want else None
        check(self._lib.skb_reduced_hessian_from_basis(self._h, MATERIAL_IDS[material], int(psd_mode), r,
                                                       None if B is None else ptr(B),
                                                       ptr(x0), ptr(z), ctypes.byref(E), ptr(g), ptr(H)))
        return float(E.value), g, H


# ------------------------------------------------------------------ plan lookup
_PLAN_CACHE = {}


def plan_from_operator(J, dim):
    """Recover ``(T, D)`` from a scipy ``J`` built by any ``deformation_jacobian`` and
    build (or fetch) its plan.  SURVEY §7: rows ``e*b + j`` (``i = 0``) hold
    ``D[j, a]`` in columns ``T[e, a]*dim``; pruned zeros do not hide vertices."""
    plan = getattr(J, "_skb_plan", None)
    if plan is not None:
        return plan
    key = id(J)
    hit = _PLAN_CACHE.get(key)
    if hit is not None and hit[0]() is J:
        return hit[1]
    if not sps.issparse(J):
        raise TypeError("plan_from_operator needs a scipy sparse J")
    b = dim * dim
    t = J.shape[0] // b
    n = J.shape[1] // dim
    K = dim + 1
    Jc = J.tocoo()
    sel = ((Jc.row % b) < dim) & ((Jc.col % dim) == 0)      # rows (i=0, j), columns (v, 0)
    e = Jc.row[sel] // b
    j = Jc.row[sel] % b
    v = Jc.col[sel] // dim
    val = Jc.data[sel]
    keys = e.astype(np.int64) * n + v
    uniq, inv = np.unique(keys, return_inverse=True)
    if uniq.size != t * K:
        raise ValueError("could not recover %d corners per element from J" % K)
    T = (uniq % n).reshape(t, K)
    D = np.zeros((t, dim, K))
    np.add.at(D, (e, j, inv % K), val)
    plan = MeshPlan(T=T, dim=dim, n=n, D=D)
    try:
        _PLAN_CACHE[key] = (weakref.ref(J, lambda _r, k=key: _PLAN_CACHE.pop(k, None)), plan)
    except TypeError:
        pass
    return plan
