"""Drop-in for ``simkit.spectral_clustering`` (spectral_clustering.py:9-48): k-means of the weighted rows of a spectral
basis, i.e. ``scipy.cluster.vq.kmeans2(W * D, k, seed=seed, minit="++")``.

scipy's algorithm is restated on the GPU (``csrc/capi_cluster.cu``: k-means++ seeding, 10 rounds of assignment + cluster
means).  The random numbers of the seeding are drawn HERE from the generator scipy builds for ``seed`` and in the order
scipy draws them (one integer for the first centre, then one uniform number per further centre), so the labels are the
reference's -- scipy and the reference are the specification of the seeding, not a detail to improve on."""

import warnings

import numpy as np

from . import _lib
from ._lib import check, f64, ptr

KMEANS2_ITER = 10          # scipy's default ``iter``, which the reference does not override


def _scipy_draws(seed, n, k):
    """First-centre index and the ``k - 1`` uniform numbers of scipy's ``_kpp`` for ``kmeans2(..., seed=seed)``."""
    from scipy._lib._util import check_random_state, rng_integers
    rng = check_random_state(seed)
    first = int(rng_integers(rng, n))
    return first, np.array([rng.uniform() for _ in range(k - 1)], dtype=np.float64)


def spectral_clustering(W, k, D=None, seed=0):
    W = np.asarray(W, dtype=np.float64)
    if D is None:
        D = np.ones((W.shape[0], 1))
    B = f64(W * D)
    if B.ndim == 1:
        B = B.reshape(-1, 1)
    n, p = B.shape
    nc = int(k)
    if nc < 1:
        raise ValueError("Cannot ask kmeans2 for %d clusters (k was %s)" % (nc, k))
    if n < 1:
        raise ValueError("Empty input is not supported.")
    first, uni = _scipy_draws(seed, n, nc)
    c = np.empty((nc, p))
    l = np.empty(n, dtype=np.int32)
    n_empty = np.zeros(1, dtype=np.int32)
    check(_lib.load().skb_kmeans2_pp(n, p, nc, KMEANS2_ITER, first, ptr(uni), ptr(B), ptr(c), ptr(l), ptr(n_empty)))
    if n_empty[0]:
        warnings.warn("One of the clusters is empty. Re-run kmeans with a different initialization.", stacklevel=2)
    return l, c
