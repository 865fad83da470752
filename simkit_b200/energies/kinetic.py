"""Kinetic (inertial) terms of the implicit integrators (reference: energies/kinetic.py:87-119,
125-194 backward Euler, 200-279 BDF2).  Small host-side vector algebra on the caller's arrays, as in
the reference; the device-resident Newton step (MeshPlan.newton) fuses these terms on the GPU."""

import numpy as np
import scipy as sp

_BE_COEFF = 1.0
_BDF2_COEFF = 9.0 / 4.0


def velocity_be(x_curr, x_prev, h):
    return (x_curr - x_prev) / h


def velocity_bdf2(x_curr, x_prev, x_prev2, h):
    return (3.0 * x_curr - 4.0 * x_prev + x_prev2) / (2.0 * h)


def be_target(x_curr, x_prev, h):
    """Backward-Euler inertial target ``x_curr + h v_curr`` (kinetic.py:87-90)."""
    return x_curr + h * velocity_be(x_curr, x_prev, h)


def bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h):
    """Constant-step BDF2 inertial target (kinetic.py:92-101)."""
    v_curr = velocity_bdf2(x_curr, x_prev, x_prev2, h)
    v_prev = velocity_bdf2(x_prev, x_prev2, x_prev3, h)
    return (4.0 / 3.0) * x_curr - (1.0 / 3.0) * x_prev + (8.0 * h / 9.0) * v_curr - (2.0 * h / 9.0) * v_prev


def kinetic_energy(d, M, h, c):
    return float((0.5 * c * (d.T @ M @ d) * (1 / (h ** 2))).item())


def kinetic_gradient(d, M, h, c):
    return c * (M @ d) * (1 / (h ** 2))


def kinetic_hessian(M, h, c):
    return M * (c / (h ** 2))


def kinetic_energy_be(x, x_curr, x_prev, M, h):
    return kinetic_energy(x - be_target(x_curr, x_prev, h), M, h, _BE_COEFF)


def kinetic_gradient_be(x, x_curr, x_prev, M, h):
    return kinetic_gradient(x - be_target(x_curr, x_prev, h), M, h, _BE_COEFF)


def kinetic_hessian_be(M, h):
    return kinetic_hessian(M, h, _BE_COEFF)


def kinetic_energy_bdf2(x, x_curr, x_prev, x_prev2, x_prev3, M, h):
    return kinetic_energy(x - bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h), M, h, _BDF2_COEFF)


def kinetic_gradient_bdf2(x, x_curr, x_prev, x_prev2, x_prev3, M, h):
    return kinetic_gradient(x - bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h), M, h, _BDF2_COEFF)


def kinetic_hessian_bdf2(M, h):
    return kinetic_hessian(M, h, _BDF2_COEFF)


# displacement (`_u`) tier: x = x_bar + u  (kinetic.py:291-407)
def kinetic_energy_be_u(u, x_curr, x_prev, M, h, x_bar):
    return kinetic_energy_be(x_bar + u, x_curr, x_prev, M, h)


def kinetic_gradient_be_u(u, x_curr, x_prev, M, h, x_bar):
    return kinetic_gradient_be(x_bar + u, x_curr, x_prev, M, h)


def kinetic_energy_bdf2_u(u, x_curr, x_prev, x_prev2, x_prev3, M, h, x_bar):
    return kinetic_energy_bdf2(x_bar + u, x_curr, x_prev, x_prev2, x_prev3, M, h)


def kinetic_gradient_bdf2_u(u, x_curr, x_prev, x_prev2, x_prev3, M, h, x_bar):
    return kinetic_gradient_bdf2(x_bar + u, x_curr, x_prev, x_prev2, x_prev3, M, h)


def kinetic_closures(x_tilde, M, h, c):
    """``(energy(x), gradient(x), hessian())`` of ``0.5 c / h^2 |x - x_tilde|_M^2`` for a FIXED inertial target, as the
    integrators use them inside one step (integrators/backward_euler.py:73-85 recompute the target and go through a
    sparse mat-vec on every evaluation; at 8 M dofs that host arithmetic costs more than the GPU side of the step).
    A diagonal ``M`` (the lumped mass every example uses) takes an element-wise path: same value to rounding."""
    from ..device_csr import diagonal_of
    s = c / (h ** 2)
    xt = np.asarray(x_tilde, dtype=np.float64).reshape(-1, 1)
    diag = diagonal_of(M) if sp.sparse.issparse(M) else None
    if diag is not None:
        md = diag.reshape(-1, 1) * s

        def energy(x):
            d = x.reshape(-1, 1) - xt
            return 0.5 * float(np.vdot(d, md * d))

        def gradient(x):
            return md * (x.reshape(-1, 1) - xt)
    else:
        def energy(x):
            return kinetic_energy(x.reshape(-1, 1) - xt, M, h, c)

        def gradient(x):
            return kinetic_gradient(x.reshape(-1, 1) - xt, M, h, c)

    Hk = kinetic_hessian(M, h, c)
    return energy, gradient, (lambda: Hk)
