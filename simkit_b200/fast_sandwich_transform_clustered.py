"""Drop-in for simkit/fast_sandwich_transform_clustered.py:15-158."""

import os
from typing import Optional

import numpy as np
import scipy as sp

from . import _lib
from ._lib import check, f64, ptr


class fast_sandwich_transform_clustered:
    """Precomputes ``ARBs[p,q,c,i,j]`` on the GPU; ``eval(r)`` / ``__call__`` contract them with ``r``."""

    def __init__(s, A, B, l: np.ndarray, read_cache: bool = False, cache_dir: Optional[str] = None, dim: int = 3) -> None:
        s.dim = dim
        l = np.asarray(l).reshape(-1)
        s.num_clusters = int(l.max()) + 1
        if cache_dir is not None and read_cache and os.path.exists(cache_dir + "/ARBs.npy"):
            s.ARBs = np.load(cache_dir + "/ARBs.npy")
            return
        Ad = f64(A.toarray() if sp.sparse.issparse(A) else A)
        Bd = f64(B.toarray() if sp.sparse.issparse(B) else B)
        t = l.shape[0]
        m1, m2 = Ad.shape[0], Bd.shape[1]
        if Ad.shape[1] != dim * dim * t or Bd.shape[0] != dim * dim * t:
            raise ValueError("A / B do not match dim*dim*len(l)")
        s.ARBs = np.empty((m1, m2, s.num_clusters, dim, dim))
        l32 = np.ascontiguousarray(l, dtype=np.int32)
        check(_lib.load().skb_fst_precompute(dim, t, m1, m2, s.num_clusters, ptr(Ad), ptr(Bd), ptr(l32), ptr(s.ARBs)))
        if cache_dir is not None:
            os.makedirs(cache_dir, exist_ok=True)
            np.save(cache_dir + "/ARBs.npy", s.ARBs)

    def __call__(s, r: np.ndarray) -> np.ndarray:
        return s.eval(r)

    def eval(s, r: np.ndarray) -> np.ndarray:
        r = f64(r).reshape((-1, s.dim, s.dim))
        assert r.shape[0] == s.num_clusters
        m1, m2 = s.ARBs.shape[0], s.ARBs.shape[1]
        out = np.empty((m1, m2))
        A = f64(s.ARBs)
        check(_lib.load().skb_fst_eval(s.dim, m1, m2, s.num_clusters, ptr(A), ptr(r), ptr(out)))
        return out
