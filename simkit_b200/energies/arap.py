"""As-rigid-as-possible energy: drop-in for the reference module (no ``lam`` anywhere).

Reference: energies/arap.py:71-145 (element F), 151-236 (element S), 242-325 (_x), 331-424 (_u),
430-499 (_S), 505-582 (self-contained); polar factor polar_svd.py:59-89, dR/dF rotation_gradient.py:12-75.
The F-representation tiers run in the CUDA library.  The stretch (``_S``) tier is a closed-form
expression in ``S`` itself (no operator, no decomposition; it is not on the SURVEY §8a path) and is
evaluated by the host exactly as the reference does.
"""

from typing import Optional

import numpy as np
import scipy as sp

from . import _tiers

_M = "arap"


def arap_energy_element_F(F: np.ndarray, mu: np.ndarray) -> np.ndarray:
    return _tiers.energy_element_F(_M, F, mu, None)


def arap_gradient_element_F(F: np.ndarray, mu: np.ndarray) -> np.ndarray:
    return _tiers.gradient_element_F(_M, F, mu, None)


def arap_hessian_element_F(F: np.ndarray, mu: np.ndarray) -> np.ndarray:
    """``mu (I - dR/dF)`` with the reference's denominator clamps."""
    return _tiers.hessian_element_F(_M, F, mu, None)


# ---- stretch (S) representation: host glue, see module docstring -------------
def _voigt_arap(k):
    # weights of the off-diagonal entries and the identity in compact form (arap.py:30-66)
    if k == 3:
        return np.array([1.0, 1.0, 2.0])[None, :], np.array([1.0, 1.0, 0.0])[None, :]
    if k == 6:
        return np.array([1.0, 1.0, 1.0, 2.0, 2.0, 2.0])[None, :], np.array([1.0, 1.0, 1.0, 0.0, 0.0, 0.0])[None, :]
    raise ValueError("Compact S must have k=3 (2D) or k=6 (3D)")


def arap_energy_element_S(S: np.ndarray, mu: np.ndarray) -> np.ndarray:
    assert S.ndim == 2 or S.ndim == 3
    mu = np.asarray(mu).reshape(-1, 1)
    if S.ndim == 3:
        d = S - np.eye(S.shape[-1])[None]
        return 0.5 * mu * np.sum(d ** 2, axis=(1, 2))[:, None]
    w, i = _voigt_arap(S.shape[-1])
    return 0.5 * mu * np.sum((S - i) ** 2 * w, axis=1)[:, None]


def arap_gradient_element_S(S: np.ndarray, mu: np.ndarray) -> np.ndarray:
    assert S.ndim == 2 or S.ndim == 3
    if S.ndim == 3:
        return np.asarray(mu).reshape(-1, 1, 1) * (S - np.eye(S.shape[-1])[None])
    w, i = _voigt_arap(S.shape[-1])
    return np.asarray(mu).reshape(-1, 1) * ((S - i) * w)


def arap_hessian_element_S(S: np.ndarray, mu: np.ndarray) -> np.ndarray:
    assert S.ndim == 2 or S.ndim == 3
    mu = np.asarray(mu).reshape(-1, 1, 1)
    if S.ndim == 3:
        b = S.shape[-1] ** 2
        return mu * np.tile(np.identity(b), (S.shape[0], 1, 1))
    w, _ = _voigt_arap(S.shape[-1])
    return mu * (w[0] * np.eye(S.shape[-1]))[None]


def arap_energy_S(S: np.ndarray, mu: np.ndarray, vol: np.ndarray) -> float:
    return float((np.asarray(vol).reshape(-1, 1) * arap_energy_element_S(S, mu)).sum())


def arap_gradient_S(S: np.ndarray, mu: np.ndarray, vol: np.ndarray) -> np.ndarray:
    P = arap_gradient_element_S(S, mu)
    w = np.asarray(vol).reshape(-1, 1, 1) if S.ndim == 3 else np.asarray(vol).reshape(-1, 1)
    return (P * w).reshape(-1, 1)


def arap_hessian_S(S: np.ndarray, mu: np.ndarray, vol: np.ndarray):
    He = arap_hessian_element_S(S, mu) * np.asarray(vol).reshape(-1, 1, 1)
    return sp.sparse.block_diag(He)


# ---- global tiers ------------------------------------------------------------
def arap_energy_x(X: np.ndarray, J, mu: np.ndarray, vol: np.ndarray) -> float:
    return _tiers.energy_x(_M, X, J, mu, None, vol)


def arap_gradient_x(X: np.ndarray, J, mu: np.ndarray, vol: np.ndarray) -> np.ndarray:
    return _tiers.gradient_x(_M, X, J, mu, None, vol)


def arap_hessian_x(X: np.ndarray, J, mu: np.ndarray, vol: np.ndarray, psd: bool = True):
    return _tiers.hessian_x(_M, X, J, mu, None, vol, psd=psd)


def arap_energy_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, vol: np.ndarray) -> float:
    return _tiers.energy_x(_M, u, J, mu, None, vol, Jx_bar=Jx_bar)


def arap_gradient_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, vol: np.ndarray) -> np.ndarray:
    return _tiers.gradient_x(_M, u, J, mu, None, vol, Jx_bar=Jx_bar)


def arap_hessian_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, vol: np.ndarray, psd: bool = True):
    return _tiers.hessian_x(_M, u, J, mu, None, vol, psd=psd, Jx_bar=Jx_bar)


def arap_energy(X: np.ndarray, T: np.ndarray, mu: np.ndarray, U: Optional[np.ndarray] = None) -> float:
    return _tiers.energy(_M, X, T, mu, None, U)


def arap_gradient(X: np.ndarray, T: np.ndarray, mu: np.ndarray, U: Optional[np.ndarray] = None) -> np.ndarray:
    return _tiers.gradient(_M, X, T, mu, None, U)


def arap_hessian(X: np.ndarray, T: np.ndarray, mu: np.ndarray, U: Optional[np.ndarray] = None, psd: bool = True):
    return _tiers.hessian(_M, X, T, mu, None, U, psd=psd)
