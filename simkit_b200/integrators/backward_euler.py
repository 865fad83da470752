"""Backward Euler: drop-in for simkit/integrators/backward_euler.py:27-91."""

from ..energies.kinetic import be_target, kinetic_closures
from ..solvers.newton import newton_solver


def backward_euler(x_curr, x_prev, energy_func, gradient_func, hessian_func, M, h, tolerance: float = 1e-6,
                   max_iter: int = 1, do_line_search: bool = True, return_info: bool = False, **solver_kw):
    pot = getattr(hessian_func, "__self__", None)
    if (pot is not None and getattr(pot, "_skb_potential", False)
            and getattr(energy_func, "__self__", None) is pot and getattr(gradient_func, "__self__", None) is pot):
        # device-resident step: kinetic term fused on the GPU (c = 1)
        return pot.implicit_step(be_target(x_curr, x_prev, h), M, 1.0 / h ** 2, tolerance=tolerance,
                                 max_iter=max_iter, do_line_search=do_line_search, return_info=return_info,
                                 **solver_kw)

    x0 = be_target(x_curr, x_prev, h)
    k_e, k_g, k_h = kinetic_closures(x0, M, h, 1.0)      # kinetic.py:125-194 with the target evaluated once

    def energy(x):
        return energy_func(x) + k_e(x)

    def gradient(x):
        return gradient_func(x) + k_g(x)

    def hessian(x):
        return hessian_func(x) + k_h()

    return newton_solver(x0, energy, gradient, hessian, tolerance=tolerance, max_iter=max_iter,
                         do_line_search=do_line_search, return_info=return_info, **solver_kw)
