"""Host-side mirror of the reference's `configs` package (configs/common, configs/mamba).

The reference loads yaml at import time and opens a hard-coded /scratch path
(configs/common/__init__.py:23); this mirror exposes the same attribute names with the same values,
has no import-time I/O and never touches a device.
"""
from . import common, mamba  # noqa: F401
