"""mtg-b200: B200-native engine for the data-parallel core of `MindTheGap find`.

The product is the CUDA library `libmtg_b200.so` (C ABI in include/mtg_b200.h) plus the C++ `mtg_find` CLI.
This package is the Python host mirror of that ABI (ctypes), used by the tests and by bench.py.
"""
from .api import Finder, FindParams, MtgError, load_library, build_library  # noqa: F401
