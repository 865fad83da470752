"""BDF2: drop-in for simkit/integrators/bdf2.py:31-101 (kinetic coefficient 9/4)."""

from ..energies.kinetic import bdf2_target, kinetic_closures
from ..solvers.newton import newton_solver


def bdf2(x_curr, x_prev, x_prev2, x_prev3, energy_func, gradient_func, hessian_func, M, h,
         tolerance: float = 1e-6, max_iter: int = 1, do_line_search: bool = True, return_info: bool = False,
         **solver_kw):
    pot = getattr(hessian_func, "__self__", None)
    if (pot is not None and getattr(pot, "_skb_potential", False)
            and getattr(energy_func, "__self__", None) is pot and getattr(gradient_func, "__self__", None) is pot):
        return pot.implicit_step(bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h), M, 9.0 / 4.0 / h ** 2,
                                 tolerance=tolerance, max_iter=max_iter, do_line_search=do_line_search,
                                 return_info=return_info, **solver_kw)

    x0 = bdf2_target(x_curr, x_prev, x_prev2, x_prev3, h)
    k_e, k_g, k_h = kinetic_closures(x0, M, h, 9.0 / 4.0)   # kinetic.py:200-279 with the target evaluated once

    def energy(x):
        return energy_func(x) + k_e(x)

    def gradient(x):
        return gradient_func(x) + k_g(x)

    def hessian(x):
        return hessian_func(x) + k_h()

    return newton_solver(x0, energy, gradient, hessian, tolerance=tolerance, max_iter=max_iter,
                         do_line_search=do_line_search, return_info=return_info, **solver_kw)
