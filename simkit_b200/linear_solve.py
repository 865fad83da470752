"""Linear solves of the Newton system on the GPU (replaces spsolve / scipy.linalg.solve at
solvers/newton.py:52,54)."""

import ctypes

import numpy as np
import scipy.sparse as sps

from . import _lib
from ._lib import check, f64, ptr


def _block_size(n):
    return 3 if n % 3 == 0 else (2 if n % 2 == 0 else 1)


def solve_sparse(H, rhs, rtol=1e-12, max_iter=20000, block=None, return_info=False):
    """Block-Jacobi PCG on any SPD scipy sparse matrix (converted to sorted CSR)."""
    H = sps.csr_matrix(H)
    H.sum_duplicates()
    n = H.shape[0]
    rhs = f64(rhs).reshape(-1)
    if rhs.size != n:
        raise ValueError("rhs size does not match the matrix")
    indptr = np.ascontiguousarray(H.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(H.indices, dtype=np.int32)
    vals = f64(H.data)
    x = np.empty(n)
    iters = ctypes.c_int(0)
    relres = ctypes.c_double(0.0)
    check(_lib.load().skb_csr_pcg(n, ptr(indptr), ptr(indices), ptr(vals), int(block or _block_size(n)), ptr(rhs),
                                  float(rtol), int(max_iter), ptr(x), ctypes.byref(iters), ctypes.byref(relres)))
    if return_info:
        return x, int(iters.value), float(relres.value)
    return x


def solve_dense(A, rhs):
    """Dense LU solve with partial pivoting (reduced-space Newton system)."""
    A = f64(A)
    n = A.shape[0]
    rhs = f64(rhs).reshape(-1)
    x = np.empty(n)
    check(_lib.load().skb_dense_solve(n, ptr(A), ptr(rhs), ptr(x)))
    return x
