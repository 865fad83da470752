"""Drop-in for ``simkit.spectral_cubature`` (spectral_cubature.py:18-74): cubature points from a k-means clustering of
the simplex-averaged spectral basis.  Averaging, clustering, the nearest element of every centroid and the cluster
volumes run on the GPU (``csrc/capi_cluster.cu``); the reference's ``k x t`` distance matrix is never formed."""

import numpy as np

from . import _lib
from ._lib import check, f64, ptr
from .average_onto_simplex import average_onto_simplex
from .operators import volume
from .spectral_clustering import spectral_clustering


def spectral_cubature(X, T, W, k, return_labels=False, return_centroids=False):
    Wt = average_onto_simplex(W, T)
    labels, centroids = spectral_clustering(Wt, k)
    m = f64(np.asarray(volume(X, T)).reshape(-1))
    nc = centroids.shape[0]
    lI = np.empty(nc, dtype=np.int64)
    mc = np.empty(nc)
    l32 = np.ascontiguousarray(labels, dtype=np.int32)
    check(_lib.load().skb_cubature_pick(Wt.shape[0], Wt.shape[1], nc, ptr(f64(Wt)), ptr(f64(centroids)), ptr(l32), ptr(m),
                                        ptr(lI), ptr(mc)))
    mc = mc[: int(np.max(labels)) + 1]          # np.bincount's length (spectral_cubature.py:66)
    ret = (lI, mc)
    if return_labels:
        ret = ret + (labels,)
    if return_centroids:
        ret = ret + (centroids,)
    return ret
