"""mbavo_b200 — B200-native blur-aware photometric tracking hot path of MBA-VO (src/ba_tracker).

The product is the CUDA library `lib/libmbavo_b200.so` behind the C-ABI of `include/mbavo.h`; this package is a thin
ctypes front end used by the tests and the benchmark.  There is NO CPU fallback: every compute entry point goes to
the CUDA library and raises if it is missing or if CUDA reports an error.

The directory is called `mba-vo_b200`, which is not a Python identifier; load it with `__graft_entry__.load_package()`
(it registers the package as `mbavo_b200`).
"""
from .api import (Context, Limits, MbavoError, library_path, load_library, optimize_trajectory, EXPORTED_SYMBOLS)  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["Context", "Limits", "MbavoError", "library_path", "load_library", "optimize_trajectory", "synth",
           "EXPORTED_SYMBOLS"]
