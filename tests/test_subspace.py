"""Subspace construction helpers: ``orthonormalize`` and ``project_into_subspace`` against outputs
frozen from the reference (tests/golden/subspace_*.npz, oracle/make_golden.py subspace) -- the oracle on the CPU,
``simkit_b200`` (GPU QR / Gram products / dense solve) on the GPU."""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import elasticity as oe

TAGS = ["subspace_tet", "subspace_tri"]


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _check(mod, g, tol):
    B, y = g["B"], g["y"]
    M = sps.diags(g["mass_diag"]).tocsc()
    Om = np.asarray(mod.orthonormalize(B, M, 1e-8))
    # the reference drops a direction when its whole ROW of R vanishes (orthonormalize.py:44-45): the dependent last column
    assert Om.shape == g["ortho_mass"].shape == (B.shape[0], 8)
    assert rel(Om, g["ortho_mass"]) < tol
    assert rel(Om.T @ (M @ Om), np.eye(8)) < 1e-9                        # mass-orthonormal
    Oi = np.asarray(mod.orthonormalize(B[:, :6]))
    assert rel(Oi, g["ortho_id"]) < tol
    zm = mod.project_into_subspace(y, B[:, :6], M)
    assert zm.shape == (6, 1) and rel(zm, g["z_mass"]) < tol
    assert rel(mod.project_into_subspace(y, B[:, :6]), g["z_id"]) < tol
    # precomputed normal equations are honoured; a vector in the span is reproduced
    BMB = B[:, :6].T @ (M @ B[:, :6])
    assert rel(mod.project_into_subspace(y, B[:, :6], M, BMB=BMB), g["z_mass"]) < tol
    c = np.arange(1.0, 7.0).reshape(-1, 1)
    assert rel(mod.project_into_subspace(B[:, :6] @ c, B[:, :6], M), c) < 1e-9


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_subspace_helpers(golden_dir, tag):
    _check(oe, np.load(os.path.join(golden_dir, tag + ".npz")), 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_subspace_helpers(golden_dir, tag):
    import simkit_b200 as sk
    _check(sk, np.load(os.path.join(golden_dir, tag + ".npz")), 1e-10)
    # a general (non-diagonal) mass matrix and a sparse basis take the host-product branches
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    n = g["B"].shape[0]
    Mg = sps.diags(g["mass_diag"]) + 1e-3 * sps.random(n, n, 0.02, random_state=1)
    Mg = (Mg + Mg.T).tocsc()
    zr = oe.project_into_subspace(g["y"], g["B"][:, :6], Mg)
    assert rel(sk.project_into_subspace(g["y"], g["B"][:, :6], Mg), zr) < 1e-10
    Bs = sps.csc_matrix(g["B"][:, :6])
    assert rel(sk.project_into_subspace(g["y"], Bs, Mg), zr) < 1e-9


def _jittered(cells, seed=2):
    from simkit_b200 import synthetic as syn
    X, T = syn.make_mesh(cells)
    rng = np.random.default_rng(seed)
    return X + 0.25 * syn.cell_size(cells, tuple(1.0 for _ in cells)) * rng.standard_normal(X.shape), T


@pytest.mark.parametrize("cells", [(6, 5, 4), (12, 9)])
def test_oracle_skinning_eigenmodes(cells):
    """The restated ARPACK shift-invert call: M-orthonormal modes that satisfy L w = lambda M w, a (near) zero constant
    mode without pins, strictly positive spectrum with pins, and the LBS Jacobian layout."""
    X, T = _jittered(cells)
    n, dim = X.shape
    L, M = oe.dirichlet_laplacian(X, T, 1), oe.massmatrix(X, T)
    W, E, B = oe.skinning_eigenmodes(X, T, 5)
    assert W.shape == (n, 5) and E.shape == (5,) and B.shape == (n * dim, n * 0 + 5 * (dim + 1) * dim)
    assert abs(E[0]) < 1e-8 * E[1] and np.all(np.diff(E) > 0)
    assert np.abs(L @ W - (M @ W) * E).max() < 1e-9 * np.abs(L @ W).max()
    assert rel(W.T @ (M @ W), np.eye(5)) < 1e-9
    bI = np.where(X[:, 0] < 0.08)[0]
    Wp, Ep, _ = oe.skinning_eigenmodes(X, T, 4, bI=bI)
    assert np.all(Ep > 0) and np.abs(Wp[bI]).max() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("cells", [(6, 5, 4), (12, 9)])
def test_gpu_skinning_eigenmodes(cells):
    """Same ARPACK recurrence with the inverse applied by the GPU PCG: eigenvalues to 1e-8, modes equal up to sign
    (simple spectrum on a jittered mesh), LBS Jacobian from them; with and without pinned vertices."""
    import simkit_b200 as sk
    X, T = _jittered(cells)
    for bI in (None, np.where(X[:, 0] < 0.08)[0]):
        k = 6
        Wo, Eo, Bo = oe.skinning_eigenmodes(X, T, k, bI=bI)
        W, E, B = sk.skinning_eigenmodes(X, T, k, bI=bI)
        assert W.shape == Wo.shape and B.shape == Bo.shape
        assert np.abs(E - Eo).max() <= 1e-8 * np.abs(Eo).max()
        sgn = np.sign(np.sum(W * Wo, axis=0))
        assert rel(W * sgn, Wo) < 1e-6
        assert rel(B, oe.lbs_jacobian(X, W)) < 1e-14
    assert rel(sk.lbs_jacobian(X, Wo), oe.lbs_jacobian(X, Wo)) == 0.0
    with pytest.raises(ValueError):
        sk.skinning_eigenmodes(X, T, 3, Aeq=sps.identity(X.shape[0]).tocsr()[:2])


# ------------------------------------------------------------------------------------------ spectral clustering / cubature
def _kmeans2_pp_numpy(B, k, seed, iters=10):
    """The algorithm csrc/capi_cluster.cu runs, step for step, in numpy: k-means++ seeding from the host-drawn numbers
    (simkit_b200.spectral_clustering._scipy_draws) with a running minimum and `target = u * total` in place of scipy's
    normalised cumulative sum, then the Lloyd rounds.  Checks the restatement (and the order of the random draws)
    against scipy on the CPU."""
    from simkit_b200.spectral_clustering import _scipy_draws
    n = B.shape[0]
    first, uni = _scipy_draws(seed, n, k)
    cen = np.empty((k, B.shape[1]))
    cen[0] = B[first]
    d2 = None
    for i in range(1, k):
        d = ((B - cen[i - 1]) ** 2).sum(axis=1)
        d2 = d if d2 is None else np.minimum(d2, d)
        cum = np.cumsum(d2)
        pick = int(np.searchsorted(cum, uni[i - 1] * cum[-1]))
        cen[i] = B[min(pick, n - 1)]
    for _ in range(iters):
        lab = np.argmin(((B[:, None, :] - cen[None, :, :]) ** 2).sum(axis=2), axis=1)
        for j in range(k):
            if np.any(lab == j):
                cen[j] = B[lab == j].mean(axis=0)
    return lab, cen


@pytest.mark.parametrize("tag", ["cubature_tet", "cubature_tri"])
def test_oracle_and_restated_kmeans_match_reference_cubature(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    X, T, W, k = g["X"], g["T"], g["W"], int(g["k"])
    lI, mc, labels, cen = oe.spectral_cubature(X, T, W, k)
    assert np.array_equal(lI, g["lI"]) and np.array_equal(labels, g["labels"])
    assert rel(mc, g["mc"]) < 1e-13 and rel(cen, g["centroids"]) < 1e-13
    l2, c2 = oe.spectral_clustering(W, k, D=g["Dw"], seed=3)
    assert np.array_equal(l2, g["labels_w"]) and rel(c2, g["centroids_w"]) < 1e-13
    # the restated algorithm with the host-drawn random numbers
    lab, c = _kmeans2_pp_numpy(oe.average_onto_simplex(W, T), k, 0)
    assert np.array_equal(lab, g["labels"]) and rel(c, g["centroids"]) < 1e-12
    lab, c = _kmeans2_pp_numpy(W * g["Dw"], k, 3)
    assert np.array_equal(lab, g["labels_w"]) and rel(c, g["centroids_w"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["cubature_tet", "cubature_tri"])
def test_gpu_spectral_cubature(golden_dir, tag):
    """Drop-ins against reference-frozen outputs: labels and cubature vertices exact, centroids / cluster volumes to
    rounding (the cluster means are tree sums); then a larger mesh against the oracle (scipy's kmeans2 itself)."""
    import simkit_b200 as sk
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    X, T, W, k = g["X"], g["T"], g["W"], int(g["k"])
    assert rel(sk.average_onto_simplex(W, T), oe.average_onto_simplex(W, T)) == 0.0
    lI, mc, labels, cen = sk.spectral_cubature(X, T, W, k, return_labels=True, return_centroids=True)
    assert np.array_equal(lI, g["lI"]) and np.array_equal(labels, g["labels"])
    assert rel(mc, g["mc"]) < 1e-12 and rel(cen, g["centroids"]) < 1e-12
    assert len(sk.spectral_cubature(X, T, W, k)) == 2 and len(sk.spectral_cubature(X, T, W, k, return_labels=True)) == 3
    l2, c2 = sk.spectral_clustering(W, k, D=g["Dw"], seed=3)
    assert np.array_equal(l2, g["labels_w"]) and rel(c2, g["centroids_w"]) < 1e-12
    # larger: 20^3 cells (48,000 tets) / 90 x 70 triangles, 10 modes, 40 clusters
    cells = (20, 20, 20) if X.shape[1] == 3 else (90, 70)
    Xb, Tb = _jittered(cells, seed=4)
    rng = np.random.default_rng(9)
    Wb = np.cos(Xb @ (2.0 * np.pi * rng.standard_normal((Xb.shape[1], 10))) + rng.random((1, 10)))
    lo, mo, labo, ceno = oe.spectral_cubature(Xb, Tb, Wb, 40)
    lg, mg, labg, ceng = sk.spectral_cubature(Xb, Tb, Wb, 40, return_labels=True, return_centroids=True)
    assert np.array_equal(labg, labo) and np.array_equal(lg, lo)
    assert rel(mg, mo) < 1e-12 and rel(ceng, ceno) < 1e-12
    with pytest.raises(ValueError):
        sk.spectral_clustering(Wb, 0)
