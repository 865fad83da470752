"""Generic quadratic energy ``0.5 x^T Q x + b^T x``: drop-in for simkit/energies/quadratic.py:15-70 (same names,
argument order and return types).

The sparse product and the reduction run in the CUDA library (``skb_quadratic``: one thread per row of ``Q``, fixed
summation order).  Inside the device-resident Newton step (``ElasticPotential(quadratic=(Q, b))``) the same kernel adds
the term without leaving the GPU and ``Q`` is added into the CSR values through a precomputed position map.
"""

import ctypes

import numpy as np
import scipy as sp

from .. import _lib
from .._lib import check, f64, ptr


def _csr(Q, n):
    """(indptr, indices, vals) of ``Q`` as canonical int32 CSR (duplicates summed, sorted columns)."""
    Qc = sp.sparse.csr_matrix(Q, copy=True)   # canonicalised below: the caller's matrix is left as it is
    if Qc.shape != (n, n):
        raise ValueError("Q must be (%d, %d), got %s" % (n, n, Qc.shape))
    Qc = Qc.astype(np.float64)
    Qc.sum_duplicates()
    Qc.sort_indices()
    if Qc.nnz >= 2**31:
        raise ValueError("Q has too many non-zeros for int32 indices")
    return (np.ascontiguousarray(Qc.indptr, dtype=np.int32), np.ascontiguousarray(Qc.indices, dtype=np.int32),
            np.ascontiguousarray(Qc.data, dtype=np.float64))


def _eval(x, Q, b, want_g):
    xx = f64(np.asarray(x, dtype=np.float64).reshape(-1))
    n = xx.size
    indptr, indices, vals = _csr(Q, n)
    bb = f64(np.asarray(b, dtype=np.float64).reshape(-1))
    if bb.size != n:
        raise ValueError("b must have %d entries" % n)
    E = ctypes.c_double(0.0)
    g = np.zeros((n, 1)) if want_g else None
    check(_lib.load().skb_quadratic(n, ptr(indptr), ptr(indices), ptr(vals), ptr(bb), ptr(xx), ctypes.byref(E), ptr(g)))
    return float(E.value), g


def quadratic_energy(x: np.ndarray, Q, b: np.ndarray) -> float:
    """``0.5 x^T Q x + b^T x`` as a Python float (quadratic.py:15-34)."""
    return _eval(x, Q, b, False)[0]


def quadratic_gradient(x: np.ndarray, Q, b: np.ndarray) -> np.ndarray:
    """``Q x + b`` as an ``(n, 1)`` array (quadratic.py:37-54)."""
    return _eval(x, Q, b, True)[1]


def quadratic_hessian(Q):
    """The Hessian is ``Q`` itself, returned as given (quadratic.py:57-70)."""
    return Q
