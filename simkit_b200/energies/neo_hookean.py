"""Classical neo-Hookean: drop-in for the reference module (same names, argument order and return types).

Reference: energies/neo_hookean.py:64-178 (element), 184-232 (_x), 238-267 (_u), 273-297 (self-contained).
All arithmetic runs in the CUDA library (include/simkit_b200.h); see energies/_tiers.py.
"""

from typing import Optional

import numpy as np
import scipy as sp

from . import _tiers

_M = "neo_hookean"


def neo_hookean_energy_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element energy density ``psi`` (t, 1); no quadrature weighting."""
    return _tiers.energy_element_F(_M, F, mu, lam)


def neo_hookean_gradient_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element first Piola-Kirchhoff stress (t, dim, dim)."""
    return _tiers.gradient_element_F(_M, F, mu, lam)


def neo_hookean_hessian_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element ``d2psi/dF2`` (t, dim*dim, dim*dim), row-major F layout, unweighted, unprojected."""
    return _tiers.hessian_element_F(_M, F, mu, lam)


def neo_hookean_energy_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Assembled energy ``float(sum(vol * psi))`` at positions ``X``."""
    return _tiers.energy_x(_M, X, J, mu, lam, vol)


def neo_hookean_gradient_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    """Assembled gradient ``J^T vec(vol * P)`` -> (n*dim, 1)."""
    return _tiers.gradient_x(_M, X, J, mu, lam, vol)


def neo_hookean_hessian_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    """Assembled Hessian ``J^T blockdiag(psd(vol * He)) J`` -> scipy csr (n*dim, n*dim), canonical sorted pattern."""
    return _tiers.hessian_x(_M, X, J, mu, lam, vol, psd=psd)


def neo_hookean_energy_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Energy at displacement ``u`` from a reference with ``Jx_bar = J @ x_bar``."""
    return _tiers.energy_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def neo_hookean_gradient_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    return _tiers.gradient_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def neo_hookean_hessian_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    return _tiers.hessian_x(_M, u, J, mu, lam, vol, psd=psd, Jx_bar=Jx_bar)


def neo_hookean_energy(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> float:
    """Self-contained tier: builds the operator and weights from rest geometry ``(X, T)``."""
    return _tiers.energy(_M, X, T, mu, lam, U)


def neo_hookean_gradient(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> np.ndarray:
    return _tiers.gradient(_M, X, T, mu, lam, U)


def neo_hookean_hessian(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None, psd: bool = True):
    return _tiers.hessian(_M, X, T, mu, lam, U, psd=psd)
