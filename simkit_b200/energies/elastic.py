"""Material dispatcher and reduced (``_z``) tier (reference: energies/elastic.py).

The reference dispatcher routes ``('linear-elasticity','arap','fcr','macklin-mueller-neo-hookean')``
(elastic.py:75); of those this library implements the two on its path (``'linear-elasticity'``,
``'arap'``) and additionally accepts the names of its other three materials.  Note the dispatcher's PSD
semantics differ from the per-material modules: eigenvalues are floored *before* the ``vol`` weighting
(elastic.py:663-664) and linear elasticity *is* projected here.
"""

from typing import Optional

import numpy as np

from . import _tiers

_NAMES = {
    "linear-elasticity": "linear_elasticity",
    "arap": "arap",
    "stable-neo-hookean": "stable_neo_hookean",
    "neo-hookean": "neo_hookean",
    "stvk": "stvk",
}
_MATERIALS = tuple(_NAMES)


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
