"""rotation-Laplace head (mirrors the reference's ``src/laplace`` package)."""
from . import rotation_laplace  # noqa: F401
