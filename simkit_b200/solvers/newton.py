"""Newton's method: drop-in for simkit/solvers/newton.py:7-75.

Same loop, stopping rule (``|alpha dx| < tol`` after the update, :61,:69) and ``return_info`` contract
(``iters`` is the last index, :67).  The linear solve ``H dx = -g`` replaces SuperLU ``spsolve`` (:52)
with this library's block-Jacobi PCG on the GPU for sparse ``H`` and a GPU dense solve for dense ``H``
(:54).  When the three callables come from one :class:`simkit_b200.ElasticPotential`, the whole loop
runs device-resident (``MeshPlan.newton``).
"""

import numpy as np
import scipy as sp

from ..backtracking_line_search import backtracking_line_search
from ..linear_solve import solve_dense, solve_sparse


def newton_solver(x0, energy_func, gradient_func, hessian_func, tolerance=1e-6, max_iter=1,
                  do_line_search=True, return_info=False, pcg_rtol=1e-12, pcg_max_iter=20000):
    pot = getattr(hessian_func, "__self__", None)
    if (pot is not None and getattr(pot, "_skb_potential", False)
            and getattr(energy_func, "__self__", None) is pot and getattr(gradient_func, "__self__", None) is pot):
        return pot.newton(x0, tolerance=tolerance, max_iter=max_iter, do_line_search=do_line_search,
                          return_info=return_info, pcg_rtol=pcg_rtol, pcg_max_iter=pcg_max_iter)
    x = x0.copy()
    if return_info:
        info = {"g": [], "dx": [], "alphas": [], "iters": -1}
    for i in range(max_iter):
        g = gradient_func(x)
        H = hessian_func(x)
        if sp.sparse.issparse(H):
            dx = solve_sparse(H, -g, rtol=pcg_rtol, max_iter=pcg_max_iter).reshape(-1, 1)
        else:
            dx = solve_dense(H, -g).reshape(-1, 1)
        if do_line_search:
            alpha, lx, ex = backtracking_line_search(energy_func, x, g, dx)
        else:
            alpha = 1.0
        x += alpha * dx
        if return_info:
            info["g"].append(g)
            info["dx"].append(dx)
            info["alphas"].append(alpha)
            info["iters"] = i
        if np.linalg.norm(alpha * dx) < tolerance:
            break
    if return_info:
        return x, info
    return x
