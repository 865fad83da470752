"""Linear solves of the Newton system on the GPU (replaces spsolve / scipy.linalg.solve at
solvers/newton.py:52,54)."""

import ctypes

import numpy as np
import scipy.sparse as sps

from . import _lib
from ._lib import check, f64, ptr


def _block_size(n):
    return 3 if n % 3 == 0 else (2 if n % 2 == 0 else 1)


def _check_solve(iters, relres, rtol, max_iter):
    """The reference's direct solves raise or warn on a singular matrix (scipy spsolve: MatrixRankWarning / NaN;
    scipy.linalg.solve: LinAlgError).  An iterative solve must not hand back garbage silently: a non-finite
    residual raises, a residual that missed the tolerance warns."""
    import warnings
    from ._lib import SimkitB200Error
    if not np.isfinite(relres):
        raise SimkitB200Error("linear solve failed: the PCG residual is not finite after %d iterations. The Newton system "
                              "must be symmetric positive definite (project the Hessian, psd=True, or add the inertia / "
                              "penalty terms before solving); a vertex with a singular diagonal block also breaks the "
                              "block-Jacobi preconditioner." % iters)
    if relres > 10.0 * max(rtol, 1e-15) and iters >= max_iter:
        warnings.warn("simkit_b200: PCG stopped at max_iter=%d with relative residual %.2e (tolerance %.1e); the returned "
                      "direction is approximate" % (max_iter, relres, rtol), RuntimeWarning, stacklevel=3)


def solve_sparse(H, rhs, rtol=1e-12, max_iter=20000, block=None, return_info=False):
    """Block-Jacobi PCG on any SPD scipy sparse matrix (converted to sorted CSR).  A lazy Hessian
    (``device_csr.DeviceCSR`` whose values are still on the device, also after ``H + M/h**2`` style sums) is solved
    in place: only the right-hand side and the solution cross PCIe."""
    from .device_csr import DeviceCSR
    if isinstance(H, DeviceCSR) and H.on_device:
        x, it, rr = H.solve(rhs, rtol=rtol, max_iter=max_iter, return_info=True)
        _check_solve(it, rr, rtol, max_iter)
        return (x, it, rr) if return_info else x
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
    _check_solve(int(iters.value), float(relres.value), rtol, max_iter)
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
