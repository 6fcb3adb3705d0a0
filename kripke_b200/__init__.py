"""kripke_b200 -- B200-native (sm_100a) implementation of LLNL/Kripke's source-iteration hot path.

The product is native: `lib/libkripke_b200.so` (C ABI of include/kripke_b200.h + hand-written CUDA)
and `lib/libkripke_host.so` / `bin/kripke.exe` (C++ host layer with the reference's Kripke::
interface).  This package is only the ctypes doorway used by the tests, bench.py and
__graft_entry__; it contains no compute and no CPU fallback: every numerical call ends in the
CUDA library and fails loudly when that library or a GPU is missing.
"""
from .api import (Problem, abi, host, have_gpu, init_device, lib_paths, build,  # noqa: F401
                  LAYOUTS, KB200Error)
