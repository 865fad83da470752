"""Quadratic pinning penalty ``1/2 |S x - y|_Gamma^2``: drop-in for simkit/dirichlet_penalty.py:65-142 (same name,
argument order, return tuples and the ``only_b`` / ``SGamma`` / ``return_SGamma`` options).

This is set-up code (a selection matrix and a diagonal scaling, built once per simulation); the resulting ``(Q, b)``
is what ``quadratic_*`` and ``ElasticPotential(quadratic=(Q, b))`` evaluate on the GPU every Newton iteration.
"""

import numpy as np
import scipy as sp


def dirichlet_penalty(bI, y, nv, gamma, only_b=False, SGamma=None, return_SGamma=False):
    y = np.asarray(y)
    assert y.ndim == 2
    d = y.shape[1]
    bI = np.asarray(bI).reshape(-1)
    nc = bI.shape[0]
    rows = (bI[:, None] * d + np.arange(d)[None, :]).ravel()
    cols = np.arange(nc * d)
    S = sp.sparse.csc_matrix((np.ones(nc * d), (rows, cols)), (nv * d, nc * d))
    if SGamma is None:
        gam = np.ones(nc) * gamma if np.isscalar(gamma) else np.asarray(gamma, dtype=np.float64).reshape(-1)
        SGamma = sp.sparse.csc_matrix((np.repeat(gam, d), (rows, cols)), (nv * d, nc * d))
    b = -SGamma @ y.reshape(-1, 1)
    out = (b,) if only_b else (sp.sparse.csc_matrix(SGamma @ S.T), b)
    if return_SGamma:
        out = out + (SGamma,)
    return out
