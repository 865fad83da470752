"""simkit_b200: B200-native implementation of simkit's per-element FEM elasticity hot path.

Same Python call surface as the reference for that path; all numerics run in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/simkit_b200.h``.  Importing the
package does not need a GPU; calling into it does, and fails loudly without one (no CPU fallback).
"""

from . import energies, integrators, solvers  # noqa: F401
from .backtracking_line_search import backtracking_line_search  # noqa: F401
from .deformation_jacobian import deformation_jacobian  # noqa: F401
from .dirichlet_laplacian import dirichlet_laplacian  # noqa: F401
from .dirichlet_penalty import dirichlet_penalty  # noqa: F401
from .energies import *  # noqa: F401,F403
from .fast_sandwich_transform_clustered import fast_sandwich_transform_clustered  # noqa: F401
from .integrators import backward_euler, bdf2, forward_euler  # noqa: F401
from .lbs_jacobian import lbs_jacobian  # noqa: F401
from .skinning_eigenmodes import skinning_eigenmodes  # noqa: F401
from .linear_solve import solve_dense, solve_sparse  # noqa: F401
from .orthonormalize import orthonormalize  # noqa: F401
from .project_into_subspace import project_into_subspace  # noqa: F401
from .average_onto_simplex import average_onto_simplex  # noqa: F401
from .spectral_clustering import spectral_clustering  # noqa: F401
from .spectral_cubature import spectral_cubature  # noqa: F401
from .operators import gravity_force, massmatrix, volume, ympr_to_lame  # noqa: F401
from .plan import MeshPlan, plan_from_operator  # noqa: F401
from .potential import ElasticPotential  # noqa: F401
from .smallmat import polar_svd, psd_project, rotation_gradient_F, svd_rv  # noqa: F401
from .solvers import newton_solver, sqp_mfem  # noqa: F401
from .stretch import (stretch, stretch_gradient, stretch_gradient_dF, stretch_gradient_dx, stretch_gradient_dz,  # noqa: F401
                      symmetric_stretch_map)

__version__ = "0.1.0"
