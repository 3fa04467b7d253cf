"""Mirror of the reference's `models` package, Mamba only (reference models/__init__.py:1)."""
from . import mamba  # noqa: F401
from .mamba import *  # noqa: F401,F403
