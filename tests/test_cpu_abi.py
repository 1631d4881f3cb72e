"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/syngular_b200.h declares.
No compute calls are made (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    header = open(os.path.join(ROOT, "include", "syngular_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(syn_[a-z0-9_]+)\s*\(", header)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    lib_path = g.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    names = declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (syn_[a-z0-9_]+)", out)))
    assert exported == names, (set(exported) ^ set(names))
    assert lib.syn_version() >= 100
    lib.syn_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.syn_last_error(), bytes)


def test_sass_is_sm100a_and_uses_the_fp64_tensor_pipe():
    lib_path = os.path.join(ROOT, "syngular_b200", "libsyngular_b200.so")
    r = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in r.stdout
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "DMMA" in sass and "LDGSTS" in sass


def test_argument_errors_do_not_need_a_gpu():
    from syngular_b200 import _lib
    rc = _lib.lib.syn_gemm_f64(None, None, None, None, None)
    assert rc != 0 and b"null descriptor" in _lib.lib.syn_last_error()
    d = _lib.GemmDesc(4, 4, 4, 1, _lib.ix(4), _lib.ix(1), _lib.ix(0), _lib.ix(4), _lib.ix(1), _lib.ix(0), _lib.ix(4), _lib.ix(1), _lib.ix(0), 1.0, 0.0)
    rc = _lib.lib.syn_gemm_f64(ctypes.byref(d), None, None, None, None)
    assert rc != 0 and b"null operand" in _lib.lib.syn_last_error()


def test_product_has_no_cpu_fallback():
    """The hot path must fail loudly without a CUDA device instead of silently computing on the host."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from syngular.tensor import MatrixProductState
    from syngular_b200._lib import SynError
    with pytest.raises(SynError):
        MatrixProductState.from_sites([np.zeros((1, 2, 1))])
    # and the product packages never import the oracle
    import sys
    for mod in ("syngular", "syngular.tensor", "syngular_b200.ops"):
        src = open(sys.modules[mod].__file__).read()
        assert "oracle" not in src
