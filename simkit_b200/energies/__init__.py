"""Energy modules of the hot path, same star-import surface as simkit/energies/__init__.py:1-16
restricted to the seven materials and the dispatcher this library implements."""
from .elastic import *  # noqa: F401,F403
from .arap import *  # noqa: F401,F403
from .fcr import *  # noqa: F401,F403
from .linear_elasticity import *  # noqa: F401,F403
from .macklin_mueller_neo_hookean import *  # noqa: F401,F403
from .stable_neo_hookean import *  # noqa: F401,F403
from .stvk import *  # noqa: F401,F403
from .neo_hookean import *  # noqa: F401,F403
from .kinetic import *  # noqa: F401,F403
from .contact_springs_plane import (contact_springs_plane_energy, contact_springs_plane_gradient,  # noqa: F401
                                    contact_springs_plane_hessian)
from .contact_springs_sphere import (contact_springs_sphere_energy, contact_springs_sphere_gradient,  # noqa: F401
                                     contact_springs_sphere_hessian)
from .quadratic import quadratic_energy, quadratic_gradient, quadratic_hessian  # noqa: F401
