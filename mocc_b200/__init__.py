"""mocc_b200 -- B200-native MoC transport sweep for MOCC (youngmit/mocc).

The product is the C-ABI CUDA library ``mocc_b200/csrc/libmocc_b200.so``
(declared in ``include/mocc_b200.h``) plus the C++ ``TransportSweeper`` plugin in
``mocc_b200/host/``.  This Python package is a thin FFI layer over the same C ABI
(used by ``bench.py``, the tests and ``__graft_entry__.py``); it contains no
compute path of its own and fails loudly when the CUDA library is missing.
"""
from .flatfile import load_arrays, save_arrays  # noqa: F401
from .capi import Sweeper, load_library, LIB_PATH, problem_from_arrays  # noqa: F401

__all__ = ["load_arrays", "save_arrays", "Sweeper", "load_library", "LIB_PATH", "problem_from_arrays"]
