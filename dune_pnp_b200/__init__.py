"""dune_pnp_b200 -- B200-native backend for dune-pnp's Newton-step hot path.

The product is the CUDA library dune_pnp_b200/libpnp_b200.so (C ABI: include/pnp_b200.h, sources in
dune_pnp_b200/csrc/).  `capi` is a thin ctypes binding used by tests/ and bench.py.
"""
from . import capi  # noqa: F401
