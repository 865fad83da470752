"""ElasticPotential: the closures a simkit sim class hands to the solver, as one object.

In the reference a user sim builds ``energy(x) / gradient(x) / hessian(x)`` closures around the
``*_x`` functions plus external forces and penalty terms and passes them to ``backward_euler`` /
``newton_solver`` (examples/interactive_demos/010_interactive_contact_plane_3D.py:81-116).  This class
is that bundle for this library's energies.  Its bound methods work as plain callables with any solver;
when ``simkit_b200``'s ``newton_solver`` / ``backward_euler`` / ``bdf2`` receive all three methods of
one instance they run the whole Newton step device-resident (``skb_newton``) instead of crossing the
host boundary per evaluation.
"""

import numpy as np
import scipy.sparse as sps

from ._lib import PSD_AFTER_VOL, PSD_NONE
from .plan import MeshPlan, plan_from_operator


class ElasticPotential:
    _skb_potential = True

    def __init__(self, material, mu, lam, vol=None, plan=None, J=None, X=None, T=None, dim=None, f_ext=None,
                 pin_k=None, pin_target=None, psd=True, coarse="auto", contact_plane=None, contact_sphere=None,
                 quadratic=None):
        """``coarse``: vertex aggregates of the two-level PCG preconditioner of the device-resident step
        (``MeshPlan.set_coarse_space``; needs the rest positions ``X``).  ``"auto"``: start with block-Jacobi and switch
        the coarse correction on (``MeshPlan.auto_aggregates`` aggregates) after the first Newton iteration whose
        PCG needed more than ``MeshPlan.COARSE_MIN_ITERS`` iterations; an integer: that many aggregates from the
        start; ``0`` / ``None``: block-Jacobi only.
        ``contact_plane``: ``dict(k=, p=, n=[, M=])`` -- penalty springs against a ground plane
        (``contact_springs_plane_*``), added to the three callables and to the device-resident step.
        ``contact_sphere``: ``dict(k=, p=, r=[, M=])`` -- the same against a sphere (``contact_springs_sphere_*``).
        ``quadratic``: ``(Q, b)`` -- a general sparse quadratic term ``1/2 x^T Q x + b^T x`` (``quadratic_*``; e.g. the
        output of ``dirichlet_penalty``), ``Q`` symmetric and inside the mesh's CSR pattern."""
        if plan is None:
            if J is not None:
                plan = plan_from_operator(J, dim if dim is not None else (X.shape[1] if X is not None else 3))
            else:
                plan = MeshPlan(X=X, T=T)
        self.plan = plan
        self.material = material
        self.mu, self.lam = mu, lam
        self.vol = plan.volume() if vol is None else vol
        self.psd_mode = PSD_AFTER_VOL if (psd and material != "linear_elasticity") else PSD_NONE
        nd = plan.ndof
        self.f_ext = None if f_ext is None else np.asarray(f_ext, dtype=np.float64).reshape(nd, 1)
        self.pin_k = None if pin_k is None else np.asarray(pin_k, dtype=np.float64).reshape(nd, 1)
        self.pin_target = None if pin_target is None else np.asarray(pin_target, dtype=np.float64).reshape(nd, 1)
        self._mat_version = 0
        self.contact_plane = None
        if contact_plane is not None:
            c = dict(contact_plane)
            M = c.get("M")
            w = None if M is None else np.asarray(sps.csr_matrix(M).diagonal() if sps.issparse(M) else np.diag(M), dtype=np.float64)
            self.contact_plane = dict(k=float(c["k"]), p=np.asarray(c["p"], dtype=np.float64).reshape(-1),
                                      n=np.asarray(c["n"], dtype=np.float64).reshape(-1), w=w)
        self.contact_sphere = None
        if contact_sphere is not None:
            c = dict(contact_sphere)
            M = c.get("M")
            w = None if M is None else np.asarray(sps.csr_matrix(M).diagonal() if sps.issparse(M) else np.diag(M), dtype=np.float64)
            self.contact_sphere = dict(k=float(c["k"]), p=np.asarray(c["p"], dtype=np.float64).reshape(-1), r=float(c["r"]), w=w)
        self.quadratic = None
        if quadratic is not None:
            Q, b = quadratic
            Q = sps.csr_matrix(Q)
            if Q.shape != (nd, nd):
                raise ValueError("quadratic: Q must be (%d, %d)" % (nd, nd))
            b = np.zeros((nd, 1)) if b is None else np.asarray(b, dtype=np.float64).reshape(nd, 1)
            self.quadratic = (Q, b)
        self._X_rest = None if X is None else np.asarray(X, dtype=np.float64).reshape(plan.n, plan.dim)
        self._coarse_auto = (coarse == "auto") and self._X_rest is not None
        if coarse and coarse != "auto" and self._X_rest is not None:
            plan.set_coarse_space(self._X_rest, int(coarse))

    # -- callables (host boundary per call) -----------------------------------------------------
    def energy(self, x):
        xx = np.asarray(x, dtype=np.float64).reshape(-1, 1)
        e = self.plan.energy(self.material, xx, self.mu, self.lam, self.vol)
        if self.f_ext is not None:
            e -= float((self.f_ext.T @ xx).item())
        if self.pin_k is not None:
            d = xx - self.pin_target
            e += 0.5 * float((self.pin_k * d * d).sum())
        if self.contact_plane is not None or self.contact_sphere is not None:
            e += self._contact("energy", xx)
        if self.quadratic is not None:
            from .energies.quadratic import quadratic_energy
            e += quadratic_energy(xx, *self.quadratic)
        return e

    def _contact(self, kind, xx):
        from .energies import contact_springs_plane as cs, contact_springs_sphere as css
        X = xx.reshape(self.plan.n, self.plan.dim)
        out = None
        c = self.contact_plane
        if c is not None:
            M = None if c["w"] is None else sps.diags(c["w"])
            out = getattr(cs, "contact_springs_plane_" + kind)(X, c["k"], c["p"], c["n"], M)
        c = self.contact_sphere
        if c is not None:
            M = None if c["w"] is None else sps.diags(c["w"])
            o2 = getattr(css, "contact_springs_sphere_" + kind)(X, c["k"], c["p"], c["r"], M)
            out = o2 if out is None else out + o2
        return out

    def gradient(self, x):
        xx = np.asarray(x, dtype=np.float64).reshape(-1, 1)
        g = self.plan.gradient(self.material, xx, self.mu, self.lam, self.vol)
        if self.f_ext is not None:
            g = g - self.f_ext
        if self.pin_k is not None:
            g = g + self.pin_k * (xx - self.pin_target)
        if self.contact_plane is not None or self.contact_sphere is not None:
            g = g + self._contact("gradient", xx)
        if self.quadratic is not None:
            from .energies.quadratic import quadratic_gradient
            g = g + quadratic_gradient(xx, *self.quadratic)
        return g

    def hessian(self, x):
        H = self.plan.hessian(self.material, x, self.mu, self.lam, self.vol, self.psd_mode)
        if self.pin_k is not None:
            H = H + sps.diags(self.pin_k.ravel())
        if self.contact_plane is not None or self.contact_sphere is not None:
            H = H + self._contact("hessian", np.asarray(x, dtype=np.float64).reshape(-1, 1))
        if self.quadratic is not None:
            H = H + self.quadratic[0]
        return H

    # -- device-resident steps -------------------------------------------------------------------
    def update_materials(self, mu=None, lam=None, vol=None):
        """Replaces (or, with no arguments, re-reads after an in-place change) the materials of this potential.  The
        device-resident step uploads ``mu`` / ``lam`` / ``vol`` once and reuses the device copies while nothing else
        has written the plan's materials -- per-element arrays at 16 M tets are 390 MB per upload."""
        if mu is not None:
            self.mu = mu
        if lam is not None:
            self.lam = lam
        if vol is not None:
            self.vol = vol
        self._mat_version += 1

    def _run(self, x0, x_tilde, mass, kin_scale, tolerance, max_iter, do_line_search, return_info, **kw):
        tok = (id(self), self._mat_version)
        if getattr(self.plan, "_mat_owner", None) != tok:
            self.plan.set_materials(self.mu, self.lam, self.vol, owner=tok)
        c = self.contact_plane
        if c is not None:
            self.plan.set_contact_plane(c["k"], c["p"], c["n"], c["w"])
        else:
            self.plan.set_contact_plane(0.0)
        c = self.contact_sphere
        if c is not None:
            self.plan.set_contact_sphere(c["k"], c["p"], c["r"], c["w"])
        else:
            self.plan.set_contact_sphere(0.0)
        if self.quadratic is None:
            if getattr(self.plan, "_quad_active", False):
                self.plan.set_quadratic(None)
        elif getattr(self.plan, "_quad_owner", None) is not self.quadratic:   # uploaded once, not per step
            self.plan.set_quadratic(*self.quadratic)
            self.plan._quad_owner = self.quadratic
        x, info = self.plan.newton(self.material, x0, psd_mode=self.psd_mode, x_tilde=x_tilde, mass=mass,
                                   kin_scale=kin_scale, f_ext=self.f_ext, pin_k=self.pin_k,
                                   pin_target=self.pin_target, max_iter=max_iter, do_line_search=do_line_search,
                                   tolerance=tolerance, **kw)
        if self._coarse_auto and info["pcg_iters"] > self.plan.COARSE_MIN_ITERS * max(1, info["iters"] + 1):
            # the next steps get the coarse correction (same iterates, fewer CG iterations)
            self.plan.set_coarse_space(self._X_rest, self.plan.auto_aggregates())
            self._coarse_auto = False
        return (x, info) if return_info else x

    def newton(self, x0, tolerance=1e-6, max_iter=1, do_line_search=True, return_info=False, **kw):
        return self._run(x0, None, None, 0.0, tolerance, max_iter, do_line_search, return_info, **kw)

    def implicit_step(self, x_tilde, M, kin_scale, tolerance=1e-6, max_iter=1, do_line_search=True,
                      return_info=False, **kw):
        """One implicit step with inertial target ``x_tilde`` and kinetic term ``kin_scale/2 |x - x_tilde|_M^2``."""
        Md = sps.csr_matrix(M)
        diag = Md.diagonal()
        if (Md - sps.diags(diag)).nnz != 0 and abs(Md - sps.diags(diag)).sum() > 0:
            raise ValueError("the device-resident step needs a lumped (diagonal) mass matrix")
        nd = self.plan.ndof
        if diag.size == self.plan.n:            # per-vertex masses: expand to dofs
            diag = np.repeat(diag, self.plan.dim)
        if diag.size != nd:
            raise ValueError("mass matrix size does not match the mesh")
        return self._run(x_tilde, x_tilde, diag, kin_scale, tolerance, max_iter, do_line_search, return_info, **kw)
