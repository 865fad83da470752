"""Macklin-Mueller stable neo-Hookean: drop-in for the reference module (same names, argument order and return types).

Reference: energies/macklin_mueller_neo_hookean.py:66-122, 125-185, 188-370 (element F), 377-478 (element S),
484-573 (_x), 579-683 (_u), 689-772 (self-contained).  ``psi = mu (1 - J) + lam/2 (1 - J)^2 + mu/2 (I_C - dim)`` (:15-17).
The stretch (``_S``) tier is the F tier evaluated at the symmetric stretch plus the compact <-> full change of
variables ``C0`` of symmetric_stretch_map.py:46-70 (index glue on the host; densities, derivatives and the PSD
projection run on the device).
All arithmetic runs in the CUDA library (include/simkit_b200.h); see energies/_tiers.py.
"""

from typing import Optional

import numpy as np
import scipy as sp

from . import _tiers

_M = "macklin_mueller_neo_hookean"


def macklin_mueller_neo_hookean_energy_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element energy density ``psi`` (t, 1); no quadrature weighting."""
    return _tiers.energy_element_F(_M, F, mu, lam)


def macklin_mueller_neo_hookean_gradient_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element first Piola-Kirchhoff stress (t, dim, dim)."""
    return _tiers.gradient_element_F(_M, F, mu, lam)


def macklin_mueller_neo_hookean_hessian_element_F(F: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Per-element ``d2psi/dF2`` (t, dim*dim, dim*dim), row-major F layout, unweighted, unprojected."""
    return _tiers.hessian_element_F(_M, F, mu, lam)

# ---- stretch (S) representation (macklin_mueller_neo_hookean.py:377-478) --------
def _stretch_embedding_map(dim: int) -> np.ndarray:
    """Dense ``(dim*dim, dim*(dim+1)//2)`` compact-to-full embedding: diagonal entries first, then the upper
    triangle row by row, each off-diagonal duplicated (symmetric_stretch_map.py:46-70)."""
    k = dim * (dim + 1) // 2
    C0 = np.zeros((dim * dim, k))
    col = {}
    c = 0
    for i in range(dim):
        col[(i, i)] = c
        c += 1
    for i in range(dim):
        for j in range(i + 1, dim):
            col[(i, j)] = col[(j, i)] = c
            c += 1
    for i in range(dim):
        for j in range(dim):
            C0[i * dim + j, col[(i, j)]] = 1.0
    return C0


def _stretch_compact_to_full(S: np.ndarray) -> np.ndarray:
    if S.ndim == 3:
        return S
    t, k = S.shape
    dim = 2 if k == 3 else 3 if k == 6 else None
    if dim is None:
        raise ValueError("Compact stretch must have 3 (2D) or 6 (3D) components, got " + str(k))
    return (S @ _stretch_embedding_map(dim).T).reshape(t, dim, dim)


def macklin_mueller_neo_hookean_energy_element_S(S: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    return macklin_mueller_neo_hookean_energy_element_F(_stretch_compact_to_full(np.asarray(S)), mu, lam)


def macklin_mueller_neo_hookean_gradient_element_S(S: np.ndarray, mu: np.ndarray, lam: np.ndarray) -> np.ndarray:
    """Full-matrix input: the F-tier stress; compact input: mapped to the independent components by ``C0^T``."""
    S = np.asarray(S)
    if S.ndim == 3:
        return macklin_mueller_neo_hookean_gradient_element_F(S, mu, lam)
    t, k = S.shape
    dim = 2 if k == 3 else 3
    Pf = macklin_mueller_neo_hookean_gradient_element_F(_stretch_compact_to_full(S), mu, lam).reshape(t, dim * dim)
    return Pf @ _stretch_embedding_map(dim)


def macklin_mueller_neo_hookean_hessian_element_S(S: np.ndarray, mu: np.ndarray, lam: np.ndarray, psd: bool = True) -> np.ndarray:
    """PSD-projected by default (consumed directly by the mixed solver, :446-478)."""
    from ..smallmat import psd_project
    S = np.asarray(S)
    if S.ndim == 3:
        Hf = macklin_mueller_neo_hookean_hessian_element_F(S, mu, lam)
        return psd_project(Hf) if psd else Hf
    t, k = S.shape
    dim = 2 if k == 3 else 3
    Hf = macklin_mueller_neo_hookean_hessian_element_F(_stretch_compact_to_full(S), mu, lam)
    C0 = _stretch_embedding_map(dim)
    H = np.einsum("ji,tjk,kl->til", C0, Hf, C0)
    return psd_project(H) if psd else H


def macklin_mueller_neo_hookean_energy_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Assembled energy ``float(sum(vol * psi))`` at positions ``X``."""
    return _tiers.energy_x(_M, X, J, mu, lam, vol)


def macklin_mueller_neo_hookean_gradient_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    """Assembled gradient ``J^T vec(vol * P)`` -> (n*dim, 1)."""
    return _tiers.gradient_x(_M, X, J, mu, lam, vol)


def macklin_mueller_neo_hookean_hessian_x(X: np.ndarray, J, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    """Assembled Hessian ``J^T blockdiag(psd(vol * He)) J`` -> scipy csr (n*dim, n*dim), canonical sorted pattern."""
    return _tiers.hessian_x(_M, X, J, mu, lam, vol, psd=psd)


def macklin_mueller_neo_hookean_energy_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> float:
    """Energy at displacement ``u`` from a reference with ``Jx_bar = J @ x_bar``."""
    return _tiers.energy_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def macklin_mueller_neo_hookean_gradient_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray) -> np.ndarray:
    return _tiers.gradient_x(_M, u, J, mu, lam, vol, Jx_bar=Jx_bar)


def macklin_mueller_neo_hookean_hessian_u(u: np.ndarray, J, Jx_bar: np.ndarray, mu: np.ndarray, lam: np.ndarray, vol: np.ndarray, psd: bool = True):
    return _tiers.hessian_x(_M, u, J, mu, lam, vol, psd=psd, Jx_bar=Jx_bar)


def macklin_mueller_neo_hookean_energy(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> float:
    """Self-contained tier: builds the operator and weights from rest geometry ``(X, T)``."""
    return _tiers.energy(_M, X, T, mu, lam, U)


def macklin_mueller_neo_hookean_gradient(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None) -> np.ndarray:
    return _tiers.gradient(_M, X, T, mu, lam, U)


def macklin_mueller_neo_hookean_hessian(X: np.ndarray, T: np.ndarray, mu: np.ndarray, lam: np.ndarray, U: Optional[np.ndarray] = None, psd: bool = True):
    return _tiers.hessian(_M, X, T, mu, lam, U, psd=psd)
