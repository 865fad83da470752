"""Drop-in for ``simkit.project_into_subspace`` (project_into_subspace.py:9-59): ``argmin_z 1/2 |B z - y|_M^2``.

The normal equations ``(B^T M B) z = B^T M y`` are formed on the GPU for a dense basis (``skb_weighted_gram``; a diagonal
``M`` is applied as row weights, any other ``M`` first multiplies ``B`` and ``y`` on the host) and solved by the GPU dense
LU (``skb_dense_solve``); a sparse ``B^T M B`` takes the sparse GPU solve.  Precomputed ``BMB`` / ``BMy`` are honoured."""

import numpy as np
import scipy as sp

from . import _lib
from ._lib import check, f64, ptr
from .device_csr import diagonal_of
from .linear_solve import solve_dense, solve_sparse


def _gram(A, w, Bm):
    A, Bm = f64(A), f64(Bm)
    n, r = A.shape
    s = Bm.shape[1]
    G = np.empty((r, s))
    check(_lib.load().skb_weighted_gram(n, r, s, ptr(A), None if w is None else ptr(f64(w)), ptr(Bm), ptr(G)))
    return G


def project_into_subspace(y, B, M=None, BMB=None, BMy=None):
    y = np.asarray(y, dtype=np.float64)
    dense_B = not sp.sparse.issparse(B)
    w = None
    MB, My = B, y.reshape(y.shape[0], -1)
    if M is not None:
        w = diagonal_of(M) if sp.sparse.issparse(M) else None
        if w is None:                                    # general mass matrix: one sparse product on the host
            MB, My = M @ B, M @ My
    if BMy is None:
        BMy = _gram(B, w, My) if dense_B else B.T @ (My if w is None else w.reshape(-1, 1) * My)
    if BMB is None:
        if dense_B:
            BMB = _gram(B, w, np.asarray(MB))
        else:
            BMB = B.T @ (MB if w is None else sp.sparse.diags(w) @ B)
    if sp.sparse.issparse(BMB):
        z = solve_sparse(BMB, np.asarray(BMy), block=1)
    else:
        z = solve_dense(np.asarray(BMB), np.asarray(BMy))
    return np.asarray(z).reshape(-1, 1)
