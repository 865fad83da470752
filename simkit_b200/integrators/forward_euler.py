"""Forward (explicit) Euler: drop-in for simkit/integrators/forward_euler.py:23-61.

One gradient evaluation (on the GPU when ``gradient_func`` is one of this package's ``*_gradient_x`` closures or an
``ElasticPotential.gradient``), a lumped-mass divide and two vector updates; the state is ``(position, velocity)``."""

import numpy as np
import scipy as sp


def forward_euler(x_curr, v_curr, gradient_func, M, h):
    x_curr = x_curr.reshape(-1, 1)
    v_curr = v_curr.reshape(-1, 1)
    f = -gradient_func(x_curr).reshape(-1, 1)
    # lumped (row-summed) mass: the acceleration is an element-wise divide, not a solve (forward_euler.py:51-56)
    if sp.sparse.issparse(M):
        m = np.asarray(M.sum(axis=1)).reshape(-1, 1)
    else:
        m = np.asarray(M).sum(axis=1).reshape(-1, 1)
    a = f / m
    return x_curr + h * v_curr, v_curr + h * a
