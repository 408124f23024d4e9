"""B200-native guided path tracer -- Python host mirror of the C-ABI in include/b200pg.h.

The directory name contains a dash, so import it through ``load_package()`` in
``__graft_entry__.py`` / ``tests/conftest.py`` (module name ``b200pg``).
The CUDA library is loaded lazily by :mod:`b200pg.api`; there is NO CPU fallback:
if ``libb200pg.so`` is missing or no CUDA device is present, calls raise.
"""
from . import _abi, scenes  # noqa: F401
