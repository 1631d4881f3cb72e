"""TEST INFRASTRUCTURE ONLY -- golden vectors for the DMRG environment blocks (reference variational/dmrg.py:65-112), produced by
the UNMODIFIED reference in the build container (oracle/shim.py).  Writes tests/golden/dmrg_blocks.npz: chains and the left / right
blocks the reference's (name-mangled private) DMRG.__left_blocks / DMRG.__right_blocks return for them.

Run:  python oracle/gen_golden_dmrg.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import shim
    shim.load_reference()
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    from syngular.variational.dmrg import DMRG
    rng = np.random.default_rng(77)

    def chain(bonds, dims, mpo):
        b = [1] + list(bonds) + [1]
        if mpo:
            return [rng.normal(size=(b[k], dims[k], dims[k], b[k + 1])) for k in range(len(dims))]
        return [rng.normal(size=(b[k], dims[k], b[k + 1])) for k in range(len(dims))]

    cases = {
        "n5_d2": (chain((2, 4, 4, 2), (2,) * 5, False), chain((3, 3, 3, 3), (2,) * 5, True)),
        "n4_d3": (chain((3, 5, 3), (3,) * 4, False), chain((2, 4, 2), (3,) * 4, True)),
        "n6_d2": (chain((2, 4, 8, 4, 2), (2,) * 6, False), chain((4, 5, 5, 5, 4), (2,) * 6, True)),
    }
    store = {"names": np.array(sorted(cases))}
    for name, (xs, ws) in cases.items():
        X, W = MPS.from_sites([c.copy() for c in xs]), MPO.from_sites([c.copy() for c in ws])
        with contextlib.redirect_stdout(io.StringIO()):
            right = DMRG._DMRG__right_blocks(operator=W, state=X)
            left = DMRG._DMRG__left_blocks(operator=W, state=X)
        n = len(xs)
        store[name + "/n"] = np.array(n)
        for k in range(n):
            store["%s/X/site%d" % (name, k)] = xs[k]
            store["%s/W/site%d" % (name, k)] = ws[k]
            if right[k] is not None:
                store["%s/right%d" % (name, k)] = np.asarray(right[k])
            if left[k] is not None:
                store["%s/left%d" % (name, k)] = np.asarray(left[k])
        print(name, "right", [None if r is None else r.shape for r in right], "left", [None if l is None else l.shape for l in left])
    np.savez_compressed(os.path.join(OUT, "dmrg_blocks.npz"), **store)


if __name__ == "__main__":
    main()
