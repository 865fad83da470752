"""SQP solver of the mixed (MFEM) system: drop-in for simkit/solvers/sqpmfem.py:7-89.

Same block elimination (``(Hu + Gu Gz^-1 Hz Gz^-1 Gu^T) du = ...``, :58-77), line search (:79-82) and stopping
rule (:86-88, on ``g_u . du``); the condensed system is solved by this library's GPU solvers (block-Jacobi PCG
for sparse ``Q``, dense LU otherwise) instead of SuperLU / LAPACK.
"""

import numpy as np
import scipy as sp

from ..backtracking_line_search import backtracking_line_search
from ..linear_solve import solve_dense, solve_sparse


def sqp_mfem(p0, energy_func, hess_blocks_func, grad_blocks_func, tolerance=1e-4, max_iter=100, do_line_search=True,
             verbose=False, pcg_rtol=1e-12, pcg_max_iter=20000):
    p = p0.copy()
    for i in range(max_iter):
        H_u, H_z, G_u, G_z, G_zi = hess_blocks_func(p)
        f_u, f_z, f_mu = grad_blocks_func(p)
        K = G_u @ G_zi @ H_z @ G_zi @ G_u.T
        Q = H_u + K
        g_u = -f_u + G_u @ G_zi @ (f_z - H_z @ G_zi @ f_mu)
        if sp.sparse.issparse(Q):
            du = solve_sparse(Q, np.asarray(g_u), rtol=pcg_rtol, max_iter=pcg_max_iter, block=1)
        else:
            du = solve_dense(np.asarray(Q), np.asarray(g_u))
        du = np.asarray(du).reshape(-1, 1)
        g_z = -(f_mu + G_u.T @ du)
        dz = G_zi @ g_z
        mu = -G_zi @ (f_z + H_z @ dz)
        g = np.vstack([f_u + G_u @ mu, f_z + G_z @ mu])
        dp = np.vstack([du, dz])
        if do_line_search:
            energy_lambda = lambda q: energy_func(np.vstack([q, mu]))  # noqa: E731
            alpha, lx, ex = backtracking_line_search(energy_lambda, p[:-mu.shape[0]], g, dp)
        else:
            alpha = 1.0
        p[:-mu.shape[0]] += alpha * dp
        p[-mu.shape[0]:] = mu
        nd = float((np.asarray(g_u).T @ du).item())
        if nd < tolerance:
            break
    return p
