"""Mesh -> operator helpers of the path (reference: volume.py, massmatrix.py, gravity_force.py,
ympr_to_lame.py).  One-off setup quantities; the device plan computes the volumes."""

import numpy as np
import scipy.sparse as sps

from .plan import MeshPlan


def _plan(X, T):
    return MeshPlan(X=np.asarray(X, dtype=np.float64), T=np.asarray(T))


def volume(V: np.ndarray, F: np.ndarray, plan=None) -> np.ndarray:
    """(m,1) signed tet volume / unsigned triangle area (volume.py:14-41)."""
    if F.shape[1] not in (3, 4) or V.shape[1] != F.shape[1] - 1:
        raise ValueError("Only triangles in 2D and tetrahedra in 3D are supported")
    return (plan or _plan(V, F)).volume()


def massmatrix(X: np.ndarray, T: np.ndarray, rho=1, plan=None) -> "sps.dia_matrix":
    """Lumped diagonal mass matrix (massmatrix.py:16-50)."""
    return sps.diags((plan or _plan(X, T)).vertex_masses(rho))


def gravity_force(X: np.ndarray, T: np.ndarray, a: float = -9.8, rho=1, plan=None) -> np.ndarray:
    """``M @ [0, a, 0]`` per vertex -- always axis 1 (gravity_force.py:36-41)."""
    g = np.zeros(X.shape)
    g[:, 1] = a
    return massmatrix(X, T, rho=rho, plan=plan) @ g


def ympr_to_lame(ym, pr):
    """(mu, lam) from Young's modulus and Poisson ratio (ympr_to_lame.py:32-33)."""
    mu = ym / (2 * (1 + pr))
    lam = ym * pr / ((1 + pr) * (1 - 2 * pr))
    return mu, lam
