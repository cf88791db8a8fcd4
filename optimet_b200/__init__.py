"""optimet_b200 -- B200-native (sm_100a) multiple-scattering hot path for OPTIMET-3D.

The product is the C-ABI library ``liboptimet_b200.so`` (hand-written CUDA kernels, declared in
``include/optimet_b200.h``) plus the C++11 host layer under ``optimet_b200/host``.  This Python
package is only a thin ctypes loader used by tests and ``bench.py``; there is no CPU fallback:
every compute entry point fails loudly when no CUDA device / extension is available.
"""
from .capi import (Library, Context, GmresOpts, OB_GMRES_ZCOMP, OB_GMRES_BELOS, OB_SOLVE_DIRECT, lib_path, load)  # noqa: F401

__all__ = ["Library", "Context", "GmresOpts", "OB_GMRES_ZCOMP", "OB_GMRES_BELOS", "OB_SOLVE_DIRECT", "lib_path", "load"]
