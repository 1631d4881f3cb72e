"""TEST INFRASTRUCTURE ONLY -- can the reference's dead-code SVD split pin the SVD oracle?  ("only lever")

The reference has no working SVD truncation; the one place that calls np.linalg.svd on a chain is the abandoned TT-SVD of
/root/reference/trash/mpo.py:51-190 (`MatrixProductOperator.decompose` / `__create_svd_core`).  This script imports that file UNMODIFIED in
this container (it needs /root/reference, so it is not part of the test suite) and checks it against its own `retrieve` (trash/mpo.py:33-49):

    python oracle/check_trash_svd.py

Finding (numpy 2.3, this container): the file imports and runs, but it is not a tensor-train decomposition -- with bonds large enough that
nothing is truncated, `retrieve(i, o)` differs from the input tensor by O(1) on N(0,1) entries (3.95 / 2.88 / 4.34 on the three cases below):
the order='F' reshapes of U and of S·V do not match the C-order axis bookkeeping of the leftover.  It therefore cannot serve as a golden
for oracle/svd_numpy.py, whose header keeps saying "parity unpinned"; what pins the SVD path instead is listed in DESIGN.md section 2
(textbook SVD rounding vs density-matrix rounding vs LAPACK at BASELINE's full size, optimality of the discarded weight, idempotence)."""
import contextlib
import importlib.util
import io
import itertools
import warnings

import numpy as np

if __name__ == "__main__":
    warnings.simplefilter("ignore")
    spec = importlib.util.spec_from_file_location("ref_trash_mpo", "/root/reference/trash/mpo.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    for ins, outs, bonds in [((2, 2, 2), (2, 2, 2), (4, 4)), ((2, 2, 2), (2, 2, 2), (2, 2)), ((2, 3, 2), (2, 2, 3), (4, 6))]:
        T = rng.normal(size=ins + outs)
        tt = mod.MatrixProductOperator(T, ins, outs, bonds)
        err = 0.0
        with contextlib.redirect_stdout(io.StringIO()):           # the file prints every step
            tt.decompose()
            for ii in itertools.product(*[range(d) for d in ins]):
                for oo in itertools.product(*[range(d) for d in outs]):
                    err = max(err, abs(tt.retrieve(ii, oo) - T[ii + oo]))
        print("in %s out %s bonds %s: cores %s, max |retrieve - tensor| = %.3e" % (ins, outs, bonds, [c.shape for c in tt.cores], err))
