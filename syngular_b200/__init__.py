"""syngular_b200 -- B200-native (sm_100a) implementation of Syngular's matrix-product hot path.

The package holds the CUDA library (csrc/ -> libsyngular_b200.so, C ABI in include/syngular_b200.h), its ctypes binding
(`_lib`) and thin op wrappers (`ops`).  The drop-in `syngular.tensor` API lives in the sibling `syngular` package.

`syngular_b200.ops` (and therefore `import syngular`) loads the shared library and FAILS LOUDLY if it is missing: there is
no CPU fallback.  Only `syngular_b200.build` can be imported without the library (it is what produces it).
"""
__version__ = "0.1.0"


def load():
    """Load the shared library now (raises SynError when it has not been built)."""
    from . import _lib
    return _lib.lib
