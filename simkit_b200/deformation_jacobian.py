"""Drop-in for ``simkit.deformation_jacobian`` (reference: deformation_jacobian.py:9-87).

The per-element operators ``D`` are computed by the device plan; the scipy
``csc_matrix`` the reference API promises is assembled directly from ``(T, D)``
(no SpGEMM) and carries the plan, so the ``*_x`` / ``*_u`` tiers that receive it
do not have to rebuild anything.
"""

import numpy as np
import scipy.sparse as sps

from .plan import MeshPlan


class DeformationJacobian(sps.csc_matrix):
    """``csc_matrix`` with the device plan of its mesh attached (``_skb_plan``)."""


def deformation_jacobian(X: np.ndarray, T: np.ndarray) -> "sps.csc_matrix":
    """Sparse ``J`` with ``(J @ x.reshape(-1,1)).reshape(-1,dim,dim)`` the row-major ``F`` blocks.

    Parameters / returns as the reference: ``X (n,dim)``, ``T (t,dim+1)`` ->
    ``csc_matrix (dim*dim*t, dim*n)`` (exact zeros pruned, as the reference's SpGEMM does).
    """
    X = np.asarray(X, dtype=np.float64)
    T = np.asarray(T)
    dt = T.shape[-1]
    T = T.reshape(-1, dt)
    dim = X.shape[1]
    if dim not in (2, 3) or dt != dim + 1:
        raise ValueError("Only dim == 2 or 3 are supported")
    plan = MeshPlan(X=X, T=T)
    D = plan.element_D()                              # (t, dim, dim+1)
    t, n = T.shape[0], X.shape[0]
    e = np.arange(t)[:, None, None, None]
    i = np.arange(dim)[None, :, None, None]
    j = np.arange(dim)[None, None, :, None]
    rows = np.broadcast_to(e * dim * dim + i * dim + j, (t, dim, dim, dt)).ravel()
    cols = np.broadcast_to(T[:, None, None, :] * dim + i, (t, dim, dim, dt)).ravel()
    vals = np.broadcast_to(D[:, None, :, :], (t, dim, dim, dt)).ravel()
    J = DeformationJacobian((vals, (rows, cols)), shape=(t * dim * dim, n * dim))
    J.sum_duplicates()
    J.eliminate_zeros()
    J._skb_plan = plan
    return J
