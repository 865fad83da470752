"""Dirichlet (deformation-gradient) Laplacian ``J^T diag(vol * mu) J``: drop-in for
simkit/dirichlet_laplacian.py:16-76 (same name, argument order, ``vector`` option, csc return type).

The Dirichlet energy ``mu/2 |F|^2`` has the element Hessian ``mu I``, so its assembled Hessian is the scalar vertex
Laplacian ``L`` repeated on every coordinate.  The fused assembly kernel produces it as the constant linear-elasticity
block with ``lam = -mu`` and no projection: ``K_ab[i][k] = vol (mu d_ik da.db + mu da[k] db[i] - mu da[i] db[k])``,
whose coordinate blocks ``i = k`` are exactly ``vol mu da.db`` (linear_elasticity.py:103-136 with ``lam = -mu``).  The
per-coordinate blocks are averaged as the reference does (:68-75).
"""

import numpy as np
import scipy as sp

from ._lib import PSD_NONE
from .plan import MeshPlan


def dirichlet_laplacian(X: np.ndarray, T: np.ndarray, mu=1, vector: bool = False) -> "sp.sparse.csc_matrix":
    X = np.asarray(X, dtype=np.float64)
    T = np.asarray(T)
    if mu is not None:
        if isinstance(mu, (int, float)):
            mu = np.ones((T.shape[0], 1)) * mu
        else:
            mu = np.asarray(mu, dtype=np.float64).reshape(-1, 1)
        assert mu.shape[0] == T.shape[0]
    n, dim = X.shape
    plan = MeshPlan(X=X, T=T)
    H = plan.hessian("linear_elasticity", X, mu, -mu, None, PSD_NONE).tocsc()
    L = sp.sparse.csc_matrix((n, n))
    for i in range(dim):
        Ii = np.arange(n) * dim + i
        L = L + H[Ii, :][:, Ii]
    L = sp.sparse.csc_matrix(L / dim)
    if vector:
        return sp.sparse.kron(L, sp.sparse.identity(dim), format="csc")
    return L
