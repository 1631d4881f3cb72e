"""syngular_b200 -- B200-native (sm_100a) implementation of Syngular's matrix-product hot path.

The package holds the CUDA library (csrc/ -> libsyngular_b200.so, C ABI in include/syngular_b200.h), its
ctypes binding and thin op wrappers.  The drop-in `syngular.tensor` API lives in the sibling `syngular` package.
Importing this package loads the shared library and FAILS LOUDLY if it is missing: there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises SynError when the .so is absent)
from ._lib import SynError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
