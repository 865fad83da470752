from .newton import newton_solver  # noqa: F401
from .sqpmfem import sqp_mfem  # noqa: F401
