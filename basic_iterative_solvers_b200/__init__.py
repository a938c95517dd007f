"""B200-native (sm_100a) iteration-loop hot path of basic_iterative_solvers.

  capi  -- ctypes binding of the C-ABI (include/bis_b200.h, lib/libbis_b200.so)
  host  -- ctypes binding of the C++ host stack (Solver / harness / methods,
           lib/libbis_host.so) that drives the device through the C-ABI

The libraries are built in-tree by `__graft_entry__.build()`.  There is no CPU
fallback: every compute entry point needs a B200 and fails loudly otherwise.
"""
from . import capi, host  # noqa: F401

__all__ = ["capi", "host"]
