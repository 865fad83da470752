"""Batched small-matrix kernels: psd_project, svd_rv, polar_svd, rotation_gradient_F
(reference: psd_project.py:12-47, svd_rv.py:8-53, polar_svd.py:59-89, rotation_gradient.py:12-75)."""

import numpy as np

from . import _lib
from ._lib import check, f64, ptr


def psd_project(H: np.ndarray, method: str = "proj") -> np.ndarray:
    """Floor eigenvalues at 1e-6 ('proj') or take |.| ('abs'); a 2-D input comes back 3-D, as in the reference."""
    H = f64(H)
    if H.ndim == 2:
        H = H[None, :, :]
    if H.ndim != 3 or H.shape[1] != H.shape[2]:
        raise ValueError("psd_project expects (n, d, d) or (d, d)")
    m = {"proj": 0, "abs": 1}.get(method, None)
    out = np.empty_like(H)
    if m is None:
        # the reference silently skips the eigenvalue edit for unknown methods; it still rebuilds
        raise ValueError("method must be 'proj' or 'abs'")
    check(_lib.load().skb_psd_project(H.shape[0], H.shape[1], ptr(H), m, ptr(out)))
    return out


def _batch(F):
    F = f64(F)
    if F.ndim == 2:
        F = F[None, :, :]
    dim = F.shape[-1]
    F = np.ascontiguousarray(F.reshape(-1, dim, dim))
    if dim not in (2, 3):
        raise ValueError("Only dim == 2 or 3 are supported")
    return F, dim


def svd_rv(F: np.ndarray):
    """``F = U S V^T`` with ``U V^T`` a proper rotation; S diagonal matrices (signed last value)."""
    F, dim = _batch(F)
    U, S, V = np.empty_like(F), np.empty_like(F), np.empty_like(F)
    check(_lib.load().skb_svd_rv(dim, F.shape[0], ptr(F), ptr(U), ptr(S), ptr(V)))
    return U, S, V


def polar_svd(F: np.ndarray, flip: bool = True):
    """``F = R S``; only ``flip=True`` is live in the reference (polar_svd.py:80-84 raises NameError)."""
    if not flip:
        raise NameError("name 'd' is not defined")  # the reference's behaviour for flip=False
    F, dim = _batch(F)
    R, S = np.empty_like(F), np.empty_like(F)
    check(_lib.load().skb_polar(dim, F.shape[0], ptr(F), ptr(R), ptr(S)))
    return R, S


def rotation_gradient_F(F: np.ndarray) -> np.ndarray:
    F, dim = _batch(F)
    K = np.empty((F.shape[0], dim * dim, dim * dim))
    check(_lib.load().skb_rotation_gradient(dim, F.shape[0], ptr(F), ptr(K)))
    return K
