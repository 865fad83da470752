"""Backward Euler: drop-in for simkit/integrators/backward_euler.py:27-91."""

from ..energies.kinetic import be_target, kinetic_energy_be, kinetic_gradient_be, kinetic_hessian_be
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

    def energy(x):
        return energy_func(x) + kinetic_energy_be(x, x_curr, x_prev, M, h)

    def gradient(x):
        return gradient_func(x) + kinetic_gradient_be(x, x_curr, x_prev, M, h)

    def hessian(x):
        return hessian_func(x) + kinetic_hessian_be(M, h)

    x0 = be_target(x_curr, x_prev, h)
    return newton_solver(x0, energy, gradient, hessian, tolerance=tolerance, max_iter=max_iter,
                         do_line_search=do_line_search, return_info=return_info, **solver_kw)
