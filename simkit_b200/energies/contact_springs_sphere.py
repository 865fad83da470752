"""Penalty springs against a sphere: drop-in for simkit/energies/contact_springs_sphere.py:242-360.

Vertices with ``|x_v - p| < r`` are pushed out along ``n_v = (x_v - p)/|x_v - p|`` (held fixed in the derivatives, as in
the reference): ``E = k/2 sum m_v (|x_v - p| - r)^2``, gradient ``k m_v (|x_v - p| - r) n_v``, Hessian blocks
``k m_v n_v n_v^T``.  Same device kernel as the plane springs (``skb_contact_springs_sphere``).
"""

import numpy as np
import scipy as sp

from .contact_springs_plane import _eval


def contact_springs_sphere_energy(X: np.ndarray, k: float, p: np.ndarray, r: float, M=None) -> float:
    return _eval(X, k, p, None, M, False, False, r=r)[0]


def contact_springs_sphere_gradient(X: np.ndarray, k: float, p: np.ndarray, r: float, M=None) -> np.ndarray:
    return _eval(X, k, p, None, M, True, False, r=r)[1]


def contact_springs_sphere_hessian(X: np.ndarray, k: float, p: np.ndarray, r: float, M=None):
    _, _, blocks, inds = _eval(X, k, p, None, M, False, True, r=r)
    nv, dim = np.asarray(X).shape
    if inds is None:
        return sp.sparse.csc_matrix((nv * dim, nv * dim))
    return sp.sparse.block_diag(list(blocks), format="csc")
