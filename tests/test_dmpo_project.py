"""DifferentialMatrixProductOperator.project (reference tensor/differential_matrix_product_operator.py:79-173) against goldens
made by the unmodified reference (oracle/gen_golden_project.py): the numpy oracle on the CPU, the CUDA product on the GPU."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dmpo_project.npz")
KEYS = ("center_site", "left_wing", "left_center", "right_center", "right_wing")


def _cases():
    g = np.load(GOLD)
    for name in [str(x) for x in g["names"]]:
        n = int(g[name + "/n"])
        xs = [g["%s/X/site%d" % (name, k)] for k in range(n)]
        ws = [g["%s/W/site%d" % (name, k)] for k in range(n)]
        for index in [int(i) for i in g[name + "/indices"]]:
            yield g, name, index, xs, ws


def _check(g, name, index, res, to_np):
    for key in KEYS:
        ref = g["%s/i%d/%s" % (name, index, key)]
        got = to_np(res[key])
        assert got.shape == ref.shape, (name, index, key, got.shape, ref.shape)
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref)), (name, index, key)


def test_oracle_project_matches_reference_golden():
    from oracle import ref_numpy as R
    for g, name, index, xs, ws in _cases():
        _check(g, name, index, R.dmpo_project(xs, ws, index), np.asarray)
    with pytest.raises(Exception):
        R.dmpo_project(xs, ws, 0)


@pytest.mark.gpu
def test_cuda_project_matches_reference_golden():
    from syngular.tensor import DifferentialMatrixProductOperator as DMPO, MatrixProductOperator as MPO, MatrixProductState as MPS
    for g, name, index, xs, ws in _cases():
        W, X = DMPO.from_sites(ws), MPS.from_sites(xs)
        assert isinstance(W, DMPO) and isinstance(W, MPO)
        _check(g, name, index, W.project(index, X), lambda t: t.cpu().numpy())
    with pytest.raises(Exception):
        W.project(0, X)                                   # empty left wing: the reference crashes inside opt_einsum
    with pytest.raises(Exception):
        W.project(1, W)                                   # "projected wings should come from a matrix product state input"
    with pytest.raises(Exception):
        W.project(len(ws) - 1, X)
