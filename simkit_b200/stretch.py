"""Polar stretch and its derivatives -- the per-element blocks of the mixed (MFEM) solver.

Drop-ins for simkit/stretch.py:9-27, simkit/stretch_gradient.py:12-131 and
simkit/symmetric_stretch_map.py:11-67 (same names, argument order and return types).  The per-element
arithmetic (polar factor, ``dR/dF``, ``dS/dF``) runs in the CUDA library; the operator products
``J^T blockdiag(dS/dF) Ci^T`` stay scipy sparse products on the host exactly as in the reference (they are
index plumbing around the element blocks, not on the SURVEY §8a path).
"""

from typing import Optional, Tuple

import numpy as np
import scipy as sp

from . import _lib
from ._lib import check, ptr
from .smallmat import _batch, polar_svd


def stretch(F: np.ndarray) -> np.ndarray:
    """Symmetric factor ``S`` of ``F = R S`` for every element, stacked as a column ``(t*d*d, 1)``."""
    R, S = polar_svd(F)
    return S.reshape(-1, 1)


def stretch_gradient_dF(F: np.ndarray) -> np.ndarray:
    """``dS/dF`` per element, ``(t, d, d, d, d)`` indexed ``[m, n, i, j] = dS_ij / dF_mn``
    (``dR/dF . F + R (x) I``, stretch_gradient.py:45-54)."""
    F, dim = _batch(F)
    out = np.empty((F.shape[0], dim * dim, dim * dim))
    check(_lib.load().skb_stretch_gradient(dim, F.shape[0], ptr(F), ptr(out)))
    return out.reshape(-1, dim, dim, dim, dim)


def stretch_gradient(F: np.ndarray) -> np.ndarray:
    """Alias of :func:`stretch_gradient_dF` (stretch_gradient.py:12-25)."""
    return stretch_gradient_dF(F)


def stretch_gradient_dx(X: np.ndarray, J, Ci=None, dim: Optional[int] = None, Jq: Optional[np.ndarray] = None):
    """``ds/dx = J^T blockdiag(dS/dF) [Ci^T]`` (stretch_gradient.py:57-99)."""
    if dim is None:
        dim = X.shape[1]
    x = np.asarray(X, dtype=np.float64).reshape(-1, 1)
    f = J @ x if Jq is None else J @ x + Jq
    F = np.asarray(f).reshape(-1, dim, dim)
    dSdF = stretch_gradient(F).reshape(-1, dim * dim, dim * dim)
    dsdx = J.T @ sp.sparse.block_diag(dSdF)
    if Ci is not None:
        dsdx = dsdx @ Ci.T
    return dsdx


def stretch_gradient_dz(z: np.ndarray, GJB, dim: int, Ci=None, GJq: Optional[np.ndarray] = None):
    """Same in reduced coordinates (stretch_gradient.py:102-131)."""
    return stretch_gradient_dx(z, GJB, Ci, dim, Jq=GJq)


def symmetric_stretch_map(t: int, dim: int) -> Tuple[sp.sparse.csc_matrix, sp.sparse.csc_matrix]:
    """Embedding ``Se (t d^2, t d(d+1)/2)`` of the independent stretch components (diagonal first, then the
    upper triangle row by row, off-diagonals duplicated) and the averaging extraction ``Sei``
    (symmetric_stretch_map.py:46-67)."""
    k = dim * (dim + 1) // 2
    col = -np.ones((dim, dim), dtype=int)
    c = 0
    for i in range(dim):
        col[i, i] = c
        c += 1
    for i in range(dim):
        for j in range(i + 1, dim):
            col[i, j] = col[j, i] = c
            c += 1
    rows = np.arange(dim * dim)
    S = sp.sparse.csc_matrix((np.ones(dim * dim), (rows, col.ravel())), shape=(dim * dim, k))
    wi = np.where(np.eye(dim, dtype=bool), 1.0, 0.5).ravel()
    Si = sp.sparse.csc_matrix((wi, (col.ravel(), rows)), shape=(k, dim * dim))
    Se = sp.sparse.kron(sp.sparse.identity(t), S)
    Sei = sp.sparse.kron(sp.sparse.identity(t), Si)
    return Se, Sei
