from .mamba import *  # noqa: F401,F403  (reference models/mamba/__init__.py:1)
from .mamba import __all__  # noqa: F401
