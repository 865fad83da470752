from .newton import newton_solver  # noqa: F401
