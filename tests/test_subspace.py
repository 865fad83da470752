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
