"""matrix-Fisher head (mirrors the reference's ``src/fisher`` package)."""
from . import fisher_utils, torch_norm_factor  # noqa: F401
