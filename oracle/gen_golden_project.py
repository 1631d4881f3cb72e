"""TEST INFRASTRUCTURE ONLY -- golden vectors for DifferentialMatrixProductOperator.project (reference
tensor/differential_matrix_product_operator.py:79-173), produced by the UNMODIFIED reference in the build container
(oracle/shim.py).  Writes tests/golden/dmpo_project.npz.

Run:  python oracle/gen_golden_project.py
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
    from syngular.tensor import MatrixProductState as MPS
    from syngular.tensor.differential_matrix_product_operator import DifferentialMatrixProductOperator as DMPO
    rng = np.random.default_rng(91)

    def chain(bonds, ins, outs=None):
        b = [1] + list(bonds) + [1]
        if outs is not None:
            return [rng.normal(size=(b[k], ins[k], outs[k], b[k + 1])) for k in range(len(ins))]
        return [rng.normal(size=(b[k], ins[k], b[k + 1])) for k in range(len(ins))]

    cases = {
        # name: (mps cores, mpo cores, indices to project on)      valid indices: 1 .. n-3 (the reference's wing einsums need >= 1 core)
        "n5_d2": (chain((2, 3, 3, 2), (2,) * 5), chain((2, 3, 3, 2), (2,) * 5, (2,) * 5), (1, 2)),
        "n6_mixed": (chain((2, 4, 5, 4, 2), (2, 3, 2, 3, 2, 2)), chain((3, 2, 4, 2, 3), (2, 3, 2, 3, 2, 2), (3, 2, 2, 2, 3, 2)), (1, 2, 3)),
        "n4_d3": (chain((3, 4, 3), (3,) * 4), chain((2, 3, 2), (3,) * 4, (2,) * 4), (1,)),
    }
    store = {"names": np.array(sorted(cases))}
    for name, (xs, ws, indices) in cases.items():
        store[name + "/n"] = np.array(len(xs))
        store[name + "/indices"] = np.array(indices)
        for k in range(len(xs)):
            store["%s/X/site%d" % (name, k)] = xs[k]
            store["%s/W/site%d" % (name, k)] = ws[k]
        for index in indices:
            X, W = MPS.from_sites([c.copy() for c in xs]), DMPO.from_sites([c.copy() for c in ws])
            W.__class__ = DMPO          # from_sites is the base class's static constructor (matrix_product_operator.py:357-384): re-class the instance
            with contextlib.redirect_stdout(io.StringIO()):
                res = W.project(index, X)
            for key, val in res.items():
                store["%s/i%d/%s" % (name, index, key)] = np.asarray(val)
            print(name, index, {k: np.asarray(v).shape for k, v in res.items()})
    np.savez_compressed(os.path.join(OUT, "dmpo_project.npz"), **store)


if __name__ == "__main__":
    main()
