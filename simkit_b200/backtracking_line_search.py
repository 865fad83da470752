"""Armijo backtracking line search: drop-in for simkit/backtracking_line_search.py:8-66.

Host control flow over a user-supplied objective ``f`` (any callable), exactly the reference's
contract; with this library's energies ``f`` is a CUDA energy evaluation.  The fully
device-resident variant lives inside ``MeshPlan.newton`` (csrc/capi_solver.cu).
"""

from typing import Callable, Tuple

import numpy as np


def backtracking_line_search(f: Callable, x0: np.ndarray, g: np.ndarray, dx: np.ndarray, alpha: float = 0.01,
                             beta: float = 0.5, max_iter: int = 100, threshold: float = 1e-12) -> Tuple[float, np.ndarray, float]:
    assert alpha > 0 and alpha <= 0.5
    assert beta > 0 and beta < 1
    assert np.ndim(x0) == np.ndim(dx)
    t = 1.0
    fx0 = f(x0)
    slope = g.T @ dx
    for _ in range(max_iter):
        x = x0 + t * dx
        fx = f(x)
        if fx <= fx0 + alpha * t * slope + threshold:
            return t, x, fx
        t = beta * t
    return 0.0, x0, fx0
