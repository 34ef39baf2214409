"""famsa_b200 -- B200-native replacement for FAMSA's two data-parallel hot paths.

The product is the C-ABI shared library ``famsa_b200/lib/libfamsa_b200.so`` (include/famsa_b200.h),
hand-written CUDA for sm_100a.  This package is the thin Python binding used by the tests, bench.py
and torch.distributed plumbing; there is NO CPU fallback: importing works without a GPU (so the
symbol table can be checked), every compute call fails loudly without one.
"""
from .binding import Engine, FamsaError, lib_path, load_library, EXPORTED_SYMBOLS  # noqa: F401
from . import seqio  # noqa: F401

__all__ = ["Engine", "FamsaError", "lib_path", "load_library", "EXPORTED_SYMBOLS", "seqio"]
