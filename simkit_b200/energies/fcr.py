"""Fixed corotational (FCR): drop-in for the reference module (same names, argument order and return types).

Reference: energies/fcr.py:41-62, 65-125, 128-302 (element), 308-397 (_x), 403-502 (_u), 508-590 (self-contained);
``psi = 2 psi_arap + lam/2 (det F - 1)^2`` (fcr.py:7-8), so the shear part carries ARAP's dR/dF clamps (rotation_gradient.py:38,65-67).
All arithmetic runs in the CUDA library (include/simkit_b200.h); see energies/_tiers.py.
"""

from typing import Optional

import numpy as np
import scipy as sp

from . import _tiers

_M = "fcr"


def fcr_energy_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element energy density ``psi`` (t, 1); no quadrature weighting."""
    return _tiers.energy_element_F(_M, F, mu, lam)


def fcr_gradient_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element first Piola-Kirchhoff stress (t, dim, dim)."""
    return _tiers.gradient_element_F(_M, F, mu, lam)


def fcr_hessian_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element ``d2psi/dF2`` (t, dim*dim, dim*dim), row-major F layout, unweighted, unprojected."""
    return _tiers.hessian_element_F(_M, F, mu, lam)


def fcr_energy_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Assembled energy ``float(sum(vol * psi))`` at positions ``X``."""
    return _tiers.energy_x(_M, X, J, mu, lam, vol)


def fcr_gradient_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    """Assembled gradient ``J^T vec(vol * P)`` -> (n*dim, 1)."""
    return _tiers.gradient_x(_M, X, J, mu, lam, vol)


def fcr_hessian_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    """Assembled Hessian ``J^T blockdiag(psd(vol * He)) J`` -> scipy csr (n*dim, n*dim), canonical sorted pattern."""
    return _tiers.hessian_x(_M, X, J, mu, lam, vol, psd=psd)


def fcr_energy_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Energy at displacement ``u`` from a reference with ``Jx_bar = J @ x_bar``."""
    return _tiers.energy_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def fcr_gradient_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    return _tiers.gradient_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def fcr_hessian_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    return _tiers.hessian_x(_M, u, J, mu, lam, vol, psd=psd, Jx_bar=Jx_bar)


def fcr_energy(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> float:
    """Self-contained tier: builds the operator and weights from rest geometry ``(X, T)``."""
    return _tiers.energy(_M, X, T, mu, lam, U)


def fcr_gradient(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> np.ndarray:
    return _tiers.gradient(_M, X, T, mu, lam, U)


def fcr_hessian(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None, psd: bool = True):
    return _tiers.hessian(_M, X, T, mu, lam, U, psd=psd)
