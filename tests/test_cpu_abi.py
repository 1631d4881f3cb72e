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


def test_jacobi_ring_schedule_is_a_complete_consistent_ordering():
    """Host logic of the multi-CTA Jacobi: every pair of row blocks meets exactly once per sweep, no block is in two places in
    one round, a slot that is not reloaded still holds the block from the previous round, and the version a load waits for is
    exactly the number of earlier stores of that block."""
    import ctypes
    import numpy as np
    from syngular_b200.build import LIB
    lib = ctypes.CDLL(LIB)
    for P in (2, 4, 8, 16, 32, 64):
        NB, R = 2 * P, 2 * P - 1
        table = np.zeros((R * P, 2), dtype=np.uint32)
        cnt = np.zeros(NB, dtype=np.uint32)
        rc = lib.syn_jacobi_ring_schedule(P, table.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        pairs, ver = set(), [0] * NB
        hold_a, hold_b = [-1] * P, [-1] * P
        loads = 0
        for r in range(R):
            used = set()
            for q in range(P):
                x, y = int(table[r * P + q, 0]), int(table[r * P + q, 1])
                a, b, fl = x & 0xFF, (x >> 8) & 0xFF, (x >> 16) & 0xFF
                assert a != b and a < NB and b < NB and a not in used and b not in used
                used.update((a, b))
                key = (min(a, b), max(a, b))
                assert key not in pairs
                pairs.add(key)
                if fl & 1:
                    assert ver[a] == (y & 0xFFFF); loads += 1
                else:
                    assert hold_a[q] == a
                if fl & 2:
                    assert ver[b] == (y >> 16); loads += 1
                else:
                    assert hold_b[q] == b
                assert bool(fl & 16) == (r == R - 1)
                hold_a[q], hold_b[q] = a, b
            assert len(used) == NB
            for q in range(P):
                x = int(table[r * P + q, 0]); a, b, fl = x & 0xFF, (x >> 8) & 0xFF, (x >> 16) & 0xFF
                if fl & 4:
                    ver[a] += 1; hold_a[q] = -1
                if fl & 8:
                    ver[b] += 1; hold_b[q] = -1
        assert len(pairs) == NB * (NB - 1) // 2
        assert ver == [int(c) for c in cnt]
        assert loads < (0.6 if P >= 8 else 0.8) * 2 * P * R        # about half the block traffic of the circle method
    assert lib.syn_jacobi_ring_schedule(3, table.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p)) != 0
