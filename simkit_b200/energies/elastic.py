"""Material dispatcher and reduced (``_z``) tier (reference: energies/elastic.py).

The reference dispatcher routes ``('linear-elasticity','arap','fcr','macklin-mueller-neo-hookean')``
(elastic.py:75); all four are implemented, and the names of this library's other three materials are
accepted as well.  Every tier of the reference is mirrored: element (``_element_F``), ``_x``, ``_u``,
``_S`` (stretch), ``_z`` / ``_filtered_z`` (reduced) and the self-contained one.  Note the dispatcher's PSD
semantics differ from the per-material modules: eigenvalues are floored *before* the ``vol`` weighting
(elastic.py:663-664) and linear elasticity *is* projected here.
"""

from typing import Optional

import numpy as np

from . import _tiers

_NAMES = {
    "linear-elasticity": "linear_elasticity",
    "arap": "arap",
    "fcr": "fcr",
    "macklin-mueller-neo-hookean": "macklin_mueller_neo_hookean",
    "stable-neo-hookean": "stable_neo_hookean",
    "neo-hookean": "neo_hookean",
    "stvk": "stvk",
}
_MATERIALS = ("linear-elasticity", "arap", "fcr", "macklin-mueller-neo-hookean")   # elastic.py:75


def _mat(material):
    try:
        return _NAMES[material]
    except KeyError:
        raise ValueError("Unknown material type: " + str(material))


def elastic_energy_element_F(F, mu, lam, material):
    return _tiers.energy_element_F(_mat(material), F, mu, lam)


def elastic_gradient_element_F(F, mu, lam, material):
    return _tiers.gradient_element_F(_mat(material), F, mu, lam)


def elastic_hessian_element_F(F, mu, lam, material, psd=True):
    H = _tiers.hessian_element_F(_mat(material), F, mu, lam)
    if psd:
        from ..smallmat import psd_project
        H = psd_project(H)
    return H


def elastic_energy_x(X, J, mu, lam, vol, material):
    return _tiers.energy_x(_mat(material), X, J, mu, lam, vol)


def elastic_gradient_x(X, J, mu, lam, vol, material):
    return _tiers.gradient_x(_mat(material), X, J, mu, lam, vol)


def elastic_hessian_x(X, J, mu, lam, vol, material, psd=True):
    """elastic.py:632-665: ``psd_project`` on the unweighted blocks, then ``* vol``."""
    return _tiers.hessian_x(_mat(material), X, J, mu, lam, vol, psd=psd, before_vol=True)


def elastic_energy_u(u, J, Jx_bar, mu, lam, vol, material):
    return _tiers.energy_x(_mat(material), u, J, mu, lam, vol, Jx_bar=Jx_bar)


def elastic_gradient_u(u, J, Jx_bar, mu, lam, vol, material):
    return _tiers.gradient_x(_mat(material), u, J, mu, lam, vol, Jx_bar=Jx_bar)


def elastic_hessian_u(u, J, Jx_bar, mu, lam, vol, material, psd=True):
    """Dispatches to the per-material ``*_hessian_u`` (floor after ``vol``), elastic.py:705-745."""
    return _tiers.hessian_x(_mat(material), u, J, mu, lam, vol, psd=psd, Jx_bar=Jx_bar)


class ElasticEnergyZPrecomp:
    """Reduced operator precompute ``JB = G J B``, ``Jx0 = G J x0`` (elastic.py:191-222).

    Host-side sparse/dense products done once per simulation, exactly as in the reference; the
    per-step work (``elastic_*_z``) runs on the device with ``JB`` as a dense operator.
    """

    def __init__(self, B, x0, G, J, dim):
        JB = G @ J @ B
        self.JB = np.ascontiguousarray(JB.toarray() if hasattr(JB, "toarray") else np.asarray(JB))
        self.dim = dim
        if x0 is None:
            x0 = np.zeros((B.shape[0], 1))
        self.Jx0 = np.asarray(G @ J @ x0).reshape(-1, 1)


def _z_eval(z, mu, lam, vol, material, precomp, psd_mode, want):
    z = np.ascontiguousarray(np.asarray(z, dtype=np.float64).reshape(-1))
    return _tiers._dense_reduced(_mat(material), z, precomp.dim, precomp.JB, precomp.Jx0, mu, lam, vol, psd_mode, want)


def elastic_energy_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None) -> float:
    """elastic.py:265-294."""
    if F is not None:
        raise NotImplementedError("precomputed F is not supported on the device path; pass F=None")
    return _z_eval(z, mu, lam, vol, material, precomp, 0, ("E",))[0]


def elastic_gradient_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None):
    """elastic.py:500-530."""
    if F is not None:
        raise NotImplementedError("precomputed F is not supported on the device path; pass F=None")
    return _z_eval(z, mu, lam, vol, material, precomp, 0, ("g",))[1]


def elastic_hessian_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None, psd: bool = True):
    """elastic.py:749-782: floor before ``vol`` (dispatcher element tier), dense ``(r, r)`` result."""
    if F is not None:
        raise NotImplementedError("precomputed F is not supported on the device path; pass F=None")
    return _z_eval(z, mu, lam, vol, material, precomp, 2 if psd else 0, ("H",))[2]


class ElasticEnergyZFilteredPrecomp(ElasticEnergyZPrecomp):
    """Reduced precompute plus the quadratic filter term ``B^T J^T (mu vol) J B`` (elastic.py:225-262)."""

    def __init__(self, B, x0, G, J, dim, mu, vol):
        super().__init__(B, x0, G, J, dim)
        import scipy.sparse as sps
        if x0 is None:
            x0 = np.zeros((B.shape[0], 1))
        AMu = sps.diags(np.asarray(mu).flatten() * np.asarray(vol).flatten())
        AMue = sps.kron(AMu, sps.eye(dim * dim))
        BJAMuJ = B.T @ (J.T @ AMue @ J)
        self.BJAMuJB = BJAMuJ @ B
        self.BJAMuJx0 = BJAMuJ @ x0


def elastic_energy_filtered_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None) -> float:
    """elastic.py:297-324."""
    e = elastic_energy_z(z, mu, lam, vol, material, precomp, F=F)
    z = np.asarray(z).reshape(-1, 1)
    return e + float((0.5 * z.T @ (precomp.BJAMuJB @ z) + z.T @ precomp.BJAMuJx0).item())


def elastic_gradient_filtered_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None):
    """elastic.py:533-560."""
    g = elastic_gradient_z(z, mu, lam, vol, material, precomp, F=F)
    return g + (precomp.BJAMuJx0 + precomp.BJAMuJB @ np.asarray(z).reshape(-1, 1))


def elastic_hessian_filtered_z(z, mu, lam, vol, material, precomp, F: Optional[np.ndarray] = None, psd: bool = True):
    """elastic.py:785-813."""
    return elastic_hessian_z(z, mu, lam, vol, material, precomp, F=F, psd=psd) + precomp.BJAMuJB


# ---- stretch (S) tier: ARAP and Macklin-Mueller only (elastic.py:327-357, 563-593, 816-846) --------
def _S_funcs(material):
    if material == "arap":
        from . import arap as m
        return (lambda S, mu, lam: m.arap_energy_element_S(S, mu), lambda S, mu, lam: m.arap_gradient_element_S(S, mu),
                lambda S, mu, lam: m.arap_hessian_element_S(S, mu))
    if material == "macklin-mueller-neo-hookean":
        from . import macklin_mueller_neo_hookean as m
        return (m.macklin_mueller_neo_hookean_energy_element_S, m.macklin_mueller_neo_hookean_gradient_element_S,
                m.macklin_mueller_neo_hookean_hessian_element_S)
    raise ValueError("Unknown or unsupported material type for S: " + str(material))


def elastic_energy_S(S, mu, lam, vol, material) -> float:
    psi = _S_funcs(material)[0](S, mu, lam)
    return float((np.asarray(vol).reshape(-1, 1) * psi).sum())


def elastic_gradient_S(S, mu, lam, vol, material):
    g = _S_funcs(material)[1](S, mu, lam)
    return g * np.asarray(vol).reshape((-1,) + (1,) * (g.ndim - 1))


def elastic_hessian_S(S, mu, lam, vol, material):
    H = _S_funcs(material)[2](S, mu, lam)
    return H * np.asarray(vol).reshape((-1,) + (1,) * (H.ndim - 1))


# ---- self-contained tier (elastic.py:360-393, 596-629, 849-877) ----------------------------------
def elastic_energy(X, T, mu, lam, material, U: Optional[np.ndarray] = None) -> float:
    return _tiers.energy(_mat(material), X, T, mu, lam, U)


def elastic_gradient(X, T, mu, lam, material, U: Optional[np.ndarray] = None):
    return _tiers.gradient(_mat(material), X, T, mu, lam, U)


def elastic_hessian(X, T, mu, lam, material, U: Optional[np.ndarray] = None, psd: bool = True):
    """Routes through ``elastic_hessian_x`` (floor before ``vol``), elastic.py:872-877."""
    return _tiers.hessian(_mat(material), X, T, mu, lam, U, psd=psd, before_vol=True)
