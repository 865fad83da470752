"""Drop-in for ``simkit.skinning_eigenmodes`` (skinning_eigenmodes.py:19-84): the ``k`` lowest generalised eigenmodes
``L w = lambda M w`` of the Dirichlet Laplacian as skinning weights, and the LBS Jacobian built from them.

The reference runs ARPACK in shift-invert mode around zero (``eigs.py:96-103``: ``eigsh(A, M=M, k=k, sigma=0,
which='LM', OPinv=..., v0=ones)``) with a UMFPACK LU of ``L`` as the inverse operator (``eigs.py:18-62``; needs cvxopt).
Here the Lanczos recurrence is the same ARPACK call -- same Krylov space, same start vector -- and the inverse operator
is this library's GPU PCG on the scalar Laplacian (``skb_csr_pcg``, Jacobi, to 1e-13): the linear solves are the whole
cost of the method (one per Lanczos step) and the only O(mesh) work.  The Laplacian itself comes out of the fused
assembly kernel (``dirichlet_laplacian``), the lumped mass out of the plan.

Without pinned vertices ``L`` is singular (constant mode); the reference's LU then divides by a rounding-size pivot.  A
PCG cannot do that, so the operator inverted is ``L + eps M`` with ``eps = 1e-10 * trace(L) / trace(M)``: the eigenvectors
are unchanged and the eigenvalues move by ``eps`` (relative 1e-10), which is subtracted again.

``Aeq`` (equality-constrained modes) leads to an indefinite saddle-point operator that the PCG cannot invert; that
variant is not provided (``ValueError``).
"""

import numpy as np
import scipy as sp
from scipy.sparse.linalg import LinearOperator, eigsh

from .dirichlet_laplacian import dirichlet_laplacian
from .lbs_jacobian import lbs_jacobian
from .linear_solve import solve_sparse
from .operators import massmatrix


class _GpuInverse(LinearOperator):
    """``v -> A^{-1} v`` by block-1 (Jacobi) PCG on the GPU; ``A`` symmetric positive definite."""

    def __init__(self, A, rtol=1e-13, max_iter=200000):
        self.A = sp.sparse.csr_matrix(A)
        self.rtol, self.max_iter = rtol, max_iter
        self.solves = 0
        self.iterations = 0
        super().__init__(np.dtype(np.float64), A.shape)

    def _matvec(self, v):
        x, it, _ = solve_sparse(self.A, np.asarray(v, dtype=np.float64).reshape(-1), rtol=self.rtol, max_iter=self.max_iter,
                                block=1, return_info=True)
        self.solves += 1
        self.iterations += it
        return x


def _lowest_modes(L, M, k):
    """``k`` eigenpairs of ``L x = lambda M x`` closest to zero, as ``eigs.py:65-111`` returns them."""
    n = L.shape[0]
    if k == 0:
        return np.zeros((0)), np.zeros((n, 0))
    if k >= n - 1:
        D, B = sp.linalg.eigh(L.toarray(), b=M.toarray())      # the reference's dense fall-back (eigs.py:104-107)
        return D[:k], B[:, :k]
    eps = 1e-10 * L.diagonal().sum() / max(M.diagonal().sum(), 1e-300)
    op = _GpuInverse(L + eps * M)
    D, B = eigsh(L, M=M, k=k, sigma=0, which="LM", OPinv=op, v0=np.ones(n))
    return D - eps, B


def skinning_eigenmodes(X, T, k, mu=1, bI=None, Aeq=None):
    X = np.asarray(X, dtype=np.float64)
    M = sp.sparse.csc_matrix(massmatrix(X, T))
    L = dirichlet_laplacian(X, T, mu=mu)
    if bI is not None:
        assert isinstance(bI, np.ndarray)
        Ii = np.setdiff1d(np.arange(X.shape[0]), bI)
        Lf = L[Ii, :][:, Ii]
        Mf = sp.sparse.diags(M.diagonal()[Ii,], 0).tocsc()
        E, Wi = _lowest_modes(Lf, Mf, k)
        W = np.zeros((X.shape[0], k))
        W[Ii, :] = Wi.real
        E = E.real
    elif Aeq is not None:
        raise ValueError("skinning_eigenmodes: equality-constrained modes (Aeq) need an indefinite saddle-point solve, "
                         "which the GPU PCG of this library does not provide")
    else:
        E, W = _lowest_modes(L, M, k)
        E, W = E.real, W.real
    return W, E, lbs_jacobian(X, W)
