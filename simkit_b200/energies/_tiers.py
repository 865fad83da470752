"""Generic three-tier implementation shared by the material modules.

Mirrors the reference layout (energies/arap.py:1-27): element tier ``*_element_F``,
global tiers ``*_x`` / ``*_u`` (take a prebuilt ``J`` and ``vol``), self-contained tier
(builds the plan from ``(X, T)``).  ``J`` may be the scipy operator of any
``deformation_jacobian`` (sparse: full-space path on the mesh plan) or a dense
ndarray (reduced path).
"""

import ctypes

import numpy as np
import scipy.sparse as sps

from .. import _lib
from .._lib import MATERIAL_IDS, PSD_AFTER_VOL, PSD_BEFORE_VOL, PSD_NONE, check, f64, material_arg, ptr
from ..plan import MeshPlan, plan_from_operator


def _elem_args(F, mu, lam):
    F = f64(F)
    dim = F.shape[-1]
    if dim not in (2, 3):
        raise ValueError("supports dim=2 or dim=3")
    F = np.ascontiguousarray(F.reshape(-1, dim, dim))
    t = F.shape[0]
    mu_a, mu_n = material_arg(mu, t, "mu")
    lam_a, lam_n = material_arg(lam, t, "lam")
    return F, dim, t, (mu_a, lam_a), (ptr(mu_a), mu_n, ptr(lam_a), lam_n)


def energy_element_F(material, F, mu, lam=None):
    F, dim, t, keep, margs = _elem_args(F, mu, lam)
    psi = np.empty((t, 1))
    check(_lib.load().skb_element_energy(MATERIAL_IDS[material], dim, t, ptr(F), *margs, ptr(psi)))
    return psi


def gradient_element_F(material, F, mu, lam=None):
    F, dim, t, keep, margs = _elem_args(F, mu, lam)
    P = np.empty((t, dim, dim))
    check(_lib.load().skb_element_gradient(MATERIAL_IDS[material], dim, t, ptr(F), *margs, ptr(P)))
    return P


def hessian_element_F(material, F, mu, lam=None):
    F, dim, t, keep, margs = _elem_args(F, mu, lam)
    H = np.empty((t, dim * dim, dim * dim))
    check(_lib.load().skb_element_hessian(MATERIAL_IDS[material], dim, t, ptr(F), *margs, ptr(H)))
    return H


# ----------------------------------------------------------------- global tiers
def _psd_mode(material, psd, before_vol=False):
    if not psd:
        return PSD_NONE
    if before_vol:
        return PSD_BEFORE_VOL
    # linear elasticity's own module ignores ``psd`` (linear_elasticity.py:199-230)
    return PSD_NONE if material == "linear_elasticity" else PSD_AFTER_VOL


def _dense_reduced(material, u, dim, J, Jx_bar, mu, lam, vol, psd_mode, want):
    """Reduced path: ``J`` is a dense (t*b, r) operator, ``u`` holds the reduced coordinates."""
    J = f64(J)
    b = dim * dim
    t = J.shape[0] // b
    r = J.shape[1]
    z = f64(u).reshape(-1)
    if z.size != r:
        raise ValueError("reduced coordinates do not match the dense operator")
    Jx0 = None if Jx_bar is None else f64(Jx_bar).reshape(-1)
    mu_a, mu_n = material_arg(mu, t, "mu")
    lam_a, lam_n = material_arg(lam, t, "lam")
    vol_a, vol_n = material_arg(vol, t, "vol")
    E = ctypes.c_double(0.0)
    g = np.empty((r, 1)) if "g" in want else None
    H = np.empty((r, r)) if "H" in want else None
    check(_lib.load().skb_reduced_gradient_hessian(
        MATERIAL_IDS[material], int(psd_mode), dim, t, r, ptr(J), ptr(Jx0), ptr(z), ptr(mu_a), mu_n, ptr(lam_a),
        lam_n, ptr(vol_a), vol_n, ctypes.byref(E), ptr(g), ptr(H)))
    return float(E.value), g, H


def energy_x(material, X, J, mu, lam, vol, Jx_bar=None):
    X = np.asarray(X)
    if not sps.issparse(J):
        return _dense_reduced(material, X, X.shape[1], J, Jx_bar, mu, lam, vol, PSD_NONE, ("E",))[0]
    return plan_from_operator(J, X.shape[1]).energy(material, X, mu, lam, vol, Fbar=Jx_bar)


def gradient_x(material, X, J, mu, lam, vol, Jx_bar=None):
    X = np.asarray(X)
    if not sps.issparse(J):
        return _dense_reduced(material, X, X.shape[1], J, Jx_bar, mu, lam, vol, PSD_NONE, ("g",))[1]
    return plan_from_operator(J, X.shape[1]).gradient(material, X, mu, lam, vol, Fbar=Jx_bar)


def hessian_x(material, X, J, mu, lam, vol, psd=True, Jx_bar=None, before_vol=False):
    X = np.asarray(X)
    mode = _psd_mode(material, psd, before_vol)
    if not sps.issparse(J):
        return _dense_reduced(material, X, X.shape[1], J, Jx_bar, mu, lam, vol, mode, ("H",))[2]
    return plan_from_operator(J, X.shape[1]).hessian(material, X, mu, lam, vol, mode, Fbar=Jx_bar)


# ----------------------------------------------------------- self-contained tier
def _self_plan(X, T):
    return MeshPlan(X=np.asarray(X, dtype=np.float64), T=np.asarray(T))


def energy(material, X, T, mu, lam, U=None):
    plan = _self_plan(X, T)
    return plan.energy(material, X if U is None else U, mu, lam, None)


def gradient(material, X, T, mu, lam, U=None):
    plan = _self_plan(X, T)
    return plan.gradient(material, X if U is None else U, mu, lam, None)


def hessian(material, X, T, mu, lam, U=None, psd=True, before_vol=False):
    plan = _self_plan(X, T)
    return plan.hessian(material, X if U is None else U, mu, lam, None, _psd_mode(material, psd, before_vol))
