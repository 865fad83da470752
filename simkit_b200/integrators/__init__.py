from .backward_euler import backward_euler  # noqa: F401
from .bdf2 import bdf2  # noqa: F401
from .forward_euler import forward_euler  # noqa: F401
