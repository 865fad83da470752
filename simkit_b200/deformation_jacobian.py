"""Drop-in for ``simkit.deformation_jacobian`` (reference: deformation_jacobian.py:9-87).

The per-element operators ``D`` are computed by the device plan; the scipy ``csc_matrix`` the reference API promises
is assembled directly from ``(T, D)`` (no SpGEMM) and carries the plan, so the ``*_x`` / ``*_u`` tiers that receive it
do not have to rebuild anything.  The matrix itself is built on first touch: the energy functions of this package only
read the attached plan, and at 16 M tets the host arrays of ``J`` (580 M stored values, 7 GB) would take longer to
build than a thousand assemblies.  Any scipy operation on ``J`` (``J @ x``, ``J.T``, ``J.data`` ...) builds it
transparently and gives exactly what the reference returns.
"""

import numpy as np
import scipy.sparse as sps

from .plan import MeshPlan


class DeformationJacobian(sps.csc_matrix):
    """``csc_matrix`` with the device plan of its mesh attached (``_skb_plan``); host arrays are built on first touch."""

    def __init__(self, arg1=None, shape=None, dtype=None, copy=False, *, maxprint=None, _plan=None, _T=None):
        self._skb_plan = None
        self._T = None
        self._d = self._i = self._p = None
        if _plan is not None:
            self._skb_plan, self._T = _plan, _T
            dim = _plan.dim
            self._shape = (int(_T.shape[0]) * dim * dim, int(_plan.n) * dim)
            self.maxprint = 50 if maxprint is None else maxprint
        else:
            super().__init__(arg1, shape=shape, dtype=dtype, copy=copy, maxprint=maxprint)

    def _materialize(self):
        if self._T is None:
            return
        T, plan = self._T, self._skb_plan
        self._T = None
        dim = plan.dim
        dt = dim + 1
        D = plan.element_D()                              # (t, dim, dim+1)
        t = T.shape[0]
        e = np.arange(t)[:, None, None, None]
        i = np.arange(dim)[None, :, None, None]
        j = np.arange(dim)[None, None, :, None]
        rows = np.broadcast_to(e * dim * dim + i * dim + j, (t, dim, dim, dt)).ravel()
        cols = np.broadcast_to(T[:, None, None, :] * dim + i, (t, dim, dim, dt)).ravel()
        vals = np.broadcast_to(D[:, None, :, :], (t, dim, dim, dt)).ravel()
        J = sps.csc_matrix((vals, (rows, cols)), shape=self._shape)
        J.sum_duplicates()
        J.eliminate_zeros()                               # exact zeros pruned, as the reference's SpGEMM does
        self._d, self._i, self._p = J.data, J.indices, J.indptr

    def _get(self, name):
        self._materialize()
        return getattr(self, name)

    data = property(lambda s: s._get("_d"), lambda s, v: setattr(s, "_d", v))
    indices = property(lambda s: s._get("_i"), lambda s, v: setattr(s, "_i", v))
    indptr = property(lambda s: s._get("_p"), lambda s, v: setattr(s, "_p", v))

    @property
    def dtype(self):
        if self._T is not None:
            return np.dtype(np.float64)
        return self._d.dtype

    def __repr__(self):
        if self._T is not None:
            return "<DeformationJacobian %dx%d of a mesh plan on the device (host arrays not built yet)>" % self._shape
        return super().__repr__()


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
    return DeformationJacobian(_plan=plan, _T=T)
