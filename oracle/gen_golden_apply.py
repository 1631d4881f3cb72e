"""TEST INFRASTRUCTURE ONLY -- golden vectors for `MatrixProductOperator.apply(operator, indices)` (reference
tensor/matrix_product_operator.py:582-651), produced by the UNMODIFIED reference in the build container (oracle/shim.py).

Writes tests/golden/mpo_apply.npz: for each case the cores of the target chain, the cores of the operator chain, the site
indices, and the dense tensor of the target AFTER the in-place application.  Note what the reference actually computes
(probed, kept as the contract): for one site it is the exact product; for m >= 2 sites the contracted block has its legs
ordered (l, in..., out..., r) but is re-split as if they were interleaved (MPO:609-626), so the result is a deterministic
scramble + QR truncation, not the operator product.  Both are pinned here.

Run:  python oracle/gen_golden_apply.py
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def dense(sites):
    T = sites[0]
    for c in sites[1:]:
        T = np.tensordot(T, c, axes=(T.ndim - 1, 0))
    return T


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import shim
    shim.load_reference()
    from syngular.tensor import MatrixProductOperator as MPO
    rng = np.random.default_rng(41)

    def rand(bonds, dims):
        b = [1] + list(bonds) + [1]
        return [rng.normal(size=(b[k], dims[k][0], dims[k][1], b[k + 1])) for k in range(len(dims))]

    cases = {
        "one_site": (rand((3, 3, 3), [(2, 2)] * 4), rand((), [(2, 2)]), [1]),
        "one_site_last": (rand((2, 3), [(2, 3), (2, 2), (3, 2)]), rand((), [(2, 4)]), [2]),
        "two_sites": (rand((3, 3, 3), [(2, 2)] * 4), rand((2,), [(2, 2)] * 2), [1, 2]),
        "three_sites": (rand((3, 4, 3), [(2, 2)] * 4), rand((2, 2), [(2, 2)] * 3), [0, 1, 2]),
        "two_sites_d3": (rand((4, 4), [(3, 3)] * 3), rand((3,), [(3, 3)] * 2), [0, 1]),
    }
    store = {"names": np.array(sorted(cases))}
    for name, (a, op, idx) in cases.items():
        A, O = MPO.from_sites([c.copy() for c in a]), MPO.from_sites([c.copy() for c in op])
        with contextlib.redirect_stdout(io.StringIO()):
            ret = A.apply(O, list(idx))
        assert ret is None
        for k, c in enumerate(a):
            store["%s/A/site%d" % (name, k)] = c
        for k, c in enumerate(op):
            store["%s/op/site%d" % (name, k)] = c
        store[name + "/A/n"], store[name + "/op/n"] = np.array(len(a)), np.array(len(op))
        store[name + "/indices"] = np.array(idx)
        store[name + "/after_dense"] = dense(A.sites)
        store[name + "/after_shapes"] = np.array([s.shape for s in A.sites])
        print(name, [s.shape for s in A.sites], float(np.abs(store[name + "/after_dense"]).max()))
    np.savez_compressed(os.path.join(OUT, "mpo_apply.npz"), **store)


if __name__ == "__main__":
    main()
