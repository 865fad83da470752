"""Drop-in for ``simkit.orthonormalize`` (orthonormalize.py:9-49): mass-weighted orthonormalisation of a column basis.

``M^{1/2} B`` is factored by a thin Householder QR on the GPU (``skb_qr_thin``: cuSOLVER geqrf + orgqr, the LAPACK
routines ``numpy.linalg.qr`` runs, so ``Q`` and ``R`` carry the same signs), rank-deficient directions -- rows of ``R``
whose absolute sum is not above ``threshold`` -- are dropped and the result is mapped back with ``M^{-1/2}``."""

import numpy as np
import scipy as sp

from . import _lib
from ._lib import check, f64, ptr


def orthonormalize(B, M=None, threshold: float = 1e-16):
    if M is None:
        M = sp.sparse.identity(B.shape[0])
    msqrt = np.sqrt(M.diagonal())
    Bm = sp.sparse.diags(msqrt, 0) @ B
    Bm = f64(Bm.toarray() if sp.sparse.issparse(Bm) else np.asarray(Bm))
    n, r = Bm.shape
    if n < r:
        raise ValueError("orthonormalize needs at least as many rows as columns")
    Q = np.empty((n, r))
    R = np.empty((r, r))
    check(_lib.load().skb_qr_thin(n, r, ptr(Bm), ptr(Q), ptr(R)))
    nonsing = np.abs(R).sum(axis=1) > threshold      # rows of R, as the reference tests them (orthonormalize.py:44-45)
    return sp.sparse.diags(1.0 / msqrt, 0) @ Q[:, nonsing]
