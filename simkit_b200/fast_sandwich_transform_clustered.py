"""Drop-in for simkit/fast_sandwich_transform_clustered.py:15-158."""

import os
from typing import Optional

import numpy as np
import scipy as sp

from . import _lib
from ._lib import check, f64, ptr


class fast_sandwich_transform_clustered:
    """Precomputes ``ARBs[p,q,c,i,j]`` on the GPU; ``eval(r)`` / ``__call__`` contract them with ``r``."""

    CHUNK_BYTES = 1 << 31     # dense bytes of A and B per chunk when the operators arrive sparse

    def __init__(s, A, B, l: np.ndarray, read_cache: bool = False, cache_dir: Optional[str] = None, dim: int = 3) -> None:
        s.dim = dim
        l = np.asarray(l).reshape(-1)
        s.num_clusters = int(l.max()) + 1
        if cache_dir is not None and read_cache and os.path.exists(cache_dir + "/ARBs.npy"):
            s.ARBs = np.load(cache_dir + "/ARBs.npy")
            return
        t = l.shape[0]
        b = dim * dim
        m1, m2 = A.shape[0], B.shape[1]
        if A.shape[1] != b * t or B.shape[0] != b * t:
            raise ValueError("A / B do not match dim*dim*len(l)")
        l32 = np.ascontiguousarray(l, dtype=np.int32)
        lib = _lib.load()
        s.ARBs = np.empty((m1, m2, s.num_clusters, dim, dim))
        # The kernel contracts dense slabs.  Sparse operators (which the reference accepts, :66-93) are densified in
        # chunks of elements so that host and device never hold more than ~CHUNK_BYTES of A and B at once (ADVICE r1:
        # A.toarray() / B.toarray() are O(m * 9t)); ARBs is a sum over elements, so the chunks' results add.
        per_elem = 8 * b * (m1 + m2)
        tc = t if not (sp.sparse.issparse(A) or sp.sparse.issparse(B)) else max(1, min(t, s.CHUNK_BYTES // max(per_elem, 1)))
        if tc >= t:
            Ad = f64(A.toarray() if sp.sparse.issparse(A) else A)
            Bd = f64(B.toarray() if sp.sparse.issparse(B) else B)
            check(lib.skb_fst_precompute(dim, t, m1, m2, s.num_clusters, ptr(Ad), ptr(Bd), ptr(l32), ptr(s.ARBs)))
        else:
            Ac = A.tocsc() if sp.sparse.issparse(A) else None
            Br = B.tocsr() if sp.sparse.issparse(B) else None
            s.ARBs[...] = 0.0
            part = np.empty_like(s.ARBs)
            for e0 in range(0, t, tc):
                e1 = min(t, e0 + tc)
                sl = slice(e0 * b, e1 * b)
                Ad = f64(Ac[:, sl].toarray() if Ac is not None else np.asarray(A)[:, sl])
                Bd = f64(Br[sl, :].toarray() if Br is not None else np.asarray(B)[sl, :])
                lc = np.ascontiguousarray(l32[e0:e1])
                check(lib.skb_fst_precompute(dim, e1 - e0, m1, m2, s.num_clusters, ptr(Ad), ptr(Bd), ptr(lc), ptr(part)))
                s.ARBs += part
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
