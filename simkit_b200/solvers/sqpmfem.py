"""SQP solver of the mixed (MFEM) system: drop-in for simkit/solvers/sqpmfem.py:7-89.

Per iteration the KKT system of the mixed problem

    [H_u   0    G_u ] [du]     [f_u ]
    [ 0   H_z   G_z ] [dz] = - [f_z ]          (G_z block diagonal, its inverse G_zi supplied by the caller)
    [G_u' G_z'   0  ] [mu]     [f_mu]

is condensed onto the primal block.  With ``Y = G_zi' H_z G_zi`` (block diagonal, formed once per iteration) and
``S = G_u Y`` the condensed system, the multiplier and the stretch update are

    (H_u + S G_u') du = -f_u + G_u (G_zi' f_z) - S f_mu ,     dz = -G_zi (f_mu + G_u' du) ,     mu = -G_zi' (f_z + H_z dz)

-- the same algebra as sqpmfem.py:58-77, with the two block-diagonal products shared between the matrix and the
right-hand side instead of being recomputed.  The condensation is sparse-times-block-diagonal algebra on the caller's
scipy blocks and runs on the host, as in the reference; the condensed solve -- the only part whose cost grows faster than
the mesh -- runs on the GPU (block-Jacobi PCG on a sparse ``Q``, pivoted LU on a dense one, replacing SuperLU / LAPACK at
sqpmfem.py:67-70).  Line search (:79-82) and stopping rule on ``g_u . du`` (:86-88) as in the reference.
"""

import numpy as np
import scipy.sparse as sps

from ..backtracking_line_search import backtracking_line_search
from ..linear_solve import solve_dense, solve_sparse


def _col(v):
    return np.asarray(v, dtype=np.float64).reshape(-1, 1)


class _Condensed:
    """One iteration's KKT blocks, condensed onto the primal unknowns."""

    def __init__(self, hess_blocks, grad_blocks):
        self.Hu, self.Hz, self.Gu, self.Gz, self.Gzi = hess_blocks
        self.fu, self.fz, self.fmu = (_col(f) for f in grad_blocks)
        GziT = self.Gzi.T
        self.Y = GziT @ self.Hz @ self.Gzi          # block diagonal: one small product per element
        self.S = self.Gu @ self.Y
        self.Q = self.Hu + self.S @ self.Gu.T
        self.rhs = _col(-self.fu + self.Gu @ (GziT @ self.fz) - self.S @ self.fmu)

    def solve_primal(self, rtol, max_iter):
        if sps.issparse(self.Q):
            du = solve_sparse(self.Q, self.rhs, rtol=rtol, max_iter=max_iter, block=1)
        else:
            du = solve_dense(np.asarray(self.Q), self.rhs)
        return _col(du)

    def recover(self, du):
        """Stretch update, multiplier and the gradient of the Lagrangian in (u, z) at that multiplier."""
        dz = _col(self.Gzi @ -(self.fmu + self.Gu.T @ du))
        mu = _col(-(self.Gzi.T @ (self.fz + self.Hz @ dz)))
        grad = np.vstack([self.fu + self.Gu @ mu, self.fz + self.Gz @ mu])
        return dz, mu, grad


def sqp_mfem(p0, energy_func, hess_blocks_func, grad_blocks_func, tolerance=1e-4, max_iter=100, do_line_search=True,
             verbose=False, pcg_rtol=1e-12, pcg_max_iter=20000):
    state = p0.copy()
    for _ in range(max_iter):
        kkt = _Condensed(hess_blocks_func(state), grad_blocks_func(state))
        du = kkt.solve_primal(pcg_rtol, pcg_max_iter)
        dz, mu, grad = kkt.recover(du)
        n_mu = mu.shape[0]
        step = np.vstack([du, dz])
        alpha = 1.0
        if do_line_search:
            alpha, _, _ = backtracking_line_search(lambda q: energy_func(np.vstack([q, mu])), state[:-n_mu], grad, step)
        state[:-n_mu] += alpha * step
        state[-n_mu:] = mu
        if float((kkt.rhs.T @ du).item()) < tolerance:
            break
    return state
