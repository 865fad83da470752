"""Penalty springs against a ground plane: drop-in for simkit/energies/contact_springs_plane.py:245-388
(same names, argument order, return types and the ``return_contact_inds`` option).

``E = k/2 sum_{v: n.(x_v - p) < 0} m_v (n.(x_v - p))^2`` with ``m = diag(M)`` (identity by default).  The per-vertex
arithmetic runs in the CUDA library (``skb_contact_springs_plane``); the Hessian is block diagonal (``k m_v n n^T`` on
contacting vertices) and is assembled into a scipy matrix from the blocks.  Inside the device-resident Newton step
(``ElasticPotential(contact_plane=...)``) the same kernel adds the term without leaving the GPU.
"""

import ctypes
from typing import Optional

import numpy as np
import scipy as sp

from .. import _lib
from .._lib import check, f64, ptr


def _eval(X, k, p, n, M, want_g, want_h, r=None):
    """Shared by the plane (``n`` given) and the sphere (``r`` given) springs."""
    X = f64(X)
    nv, dim = X.shape
    p = f64(np.asarray(p, dtype=np.float64).reshape(-1))
    nrm = None if n is None else f64(np.asarray(n, dtype=np.float64).reshape(-1))
    if p.size != dim or (nrm is not None and nrm.size != dim):
        raise ValueError("p and n must have dim entries")
    w = None if M is None else f64(sp.sparse.csr_matrix(M).diagonal() if sp.sparse.issparse(M) else np.diag(np.asarray(M)))
    E = ctypes.c_double(0.0)
    g = np.zeros((nv * dim, 1)) if want_g else None
    blocks = np.empty((nv, dim, dim)) if want_h else None
    under = np.empty(nv, dtype=np.int32)
    if r is None:
        check(_lib.load().skb_contact_springs_plane(dim, nv, ptr(X), float(k), ptr(p), ptr(nrm), ptr(w), ctypes.byref(E),
                                                   ptr(g), ptr(blocks), ptr(under)))
    else:
        check(_lib.load().skb_contact_springs_sphere(dim, nv, ptr(X), float(k), ptr(p), float(r), ptr(w), ctypes.byref(E),
                                                    ptr(g), ptr(blocks), ptr(under)))
    inds = np.where(under != 0)[0][:, None] if under.any() else None
    return float(E.value), g, blocks, inds


def contact_springs_plane_energy(X: np.ndarray, k: float, p: np.ndarray, n: np.ndarray, M=None,
                                 return_contact_inds: bool = False):
    E, _, _, inds = _eval(X, k, p, n, M, False, False)
    return (E, inds) if return_contact_inds else E


def contact_springs_plane_gradient(X: np.ndarray, k: float, p: np.ndarray, n: np.ndarray, M=None,
                                   return_contact_inds: bool = False):
    _, g, _, inds = _eval(X, k, p, n, M, True, False)
    return (g, inds) if return_contact_inds else g


def contact_springs_plane_hessian(X: np.ndarray, k: float, p: np.ndarray, n: np.ndarray, M=None,
                                  return_contact_inds: bool = False):
    _, _, blocks, inds = _eval(X, k, p, n, M, False, True)
    nv, dim = np.asarray(X).shape
    if inds is None:
        H = sp.sparse.csc_matrix((nv * dim, nv * dim))
    else:
        H = sp.sparse.block_diag(list(blocks), format="csc")
    return (H, inds) if return_contact_inds else H
