"""Drop-in for ``simkit.average_onto_simplex`` (average_onto_simplex.py:8-37): per-simplex mean of per-vertex values,
one thread per (simplex, column) on the GPU, corners added in the reference's order."""

import numpy as np

from . import _lib
from ._lib import check, f64, ptr


def average_onto_simplex(A, T):
    A = f64(np.asarray(A))
    if A.ndim == 1:
        A = A.reshape(-1, 1)
    T32 = np.ascontiguousarray(np.asarray(T), dtype=np.int32)
    if T32.ndim != 2 or T32.size == 0 or A.size == 0:
        raise ValueError("average_onto_simplex needs a (t, s) connectivity and (n, d) values")
    if T32.min() < 0 or T32.max() >= A.shape[0]:
        raise IndexError("simplex references a vertex outside the value array")
    At = np.empty((T32.shape[0], A.shape[1]))
    check(_lib.load().skb_average_onto_simplex(A.shape[0], A.shape[1], T32.shape[0], T32.shape[1], ptr(A), ptr(T32), ptr(At)))
    return At
