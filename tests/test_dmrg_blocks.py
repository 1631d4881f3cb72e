"""DMRG environment blocks (reference variational/dmrg.py:65-112) against goldens made by the unmodified reference
(oracle/gen_golden_dmrg.py): the numpy oracle and the product's host logic on the CPU (kernels emulated), the CUDA product on the GPU."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dmrg_blocks.npz")


def _cases():
    g = np.load(GOLD)
    for name in [str(x) for x in g["names"]]:
        n = int(g[name + "/n"])
        xs = [g["%s/X/site%d" % (name, k)] for k in range(n)]
        ws = [g["%s/W/site%d" % (name, k)] for k in range(n)]
        yield g, name, n, xs, ws


def _check(g, name, n, right, left, to_np):
    for k in range(n):
        for tag, blocks in (("right", right), ("left", left)):
            key = "%s/%s%d" % (name, tag, k)
            if key in g.files:
                ref = g[key]
                got = to_np(blocks[k])
                assert got.shape == ref.shape
                assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref)), key
            else:
                assert blocks[k] is None, key          # the reference leaves these entries None (dmrg.py:67,92)


def test_oracle_blocks_match_reference():
    from oracle import ref_numpy as R
    for g, name, n, xs, ws in _cases():
        _check(g, name, n, R.dmrg_right_blocks(xs, ws), R.dmrg_left_blocks(xs, ws), np.asarray)


def test_product_host_logic_with_emulated_kernels():
    import cpu_ops
    with cpu_ops.patched():
        from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
        from syngular.variational import DMRG
        for g, name, n, xs, ws in _cases():
            W, X = MPO.from_sites(ws), MPS.from_sites(xs)
            _check(g, name, n, DMRG.right_blocks(W, X), DMRG.left_blocks(W, X), lambda t: t.numpy())


@pytest.mark.gpu
def test_cuda_blocks_match_reference():
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    from syngular.variational import DMRG
    for g, name, n, xs, ws in _cases():
        W, X = MPO.from_sites(ws), MPS.from_sites(xs)
        _check(g, name, n, DMRG.right_blocks(W, X), DMRG.left_blocks(W, X), lambda t: t.cpu().numpy())
    with pytest.raises(NotImplementedError):
        DMRG.solve(None)


@pytest.mark.gpu
def test_cuda_blocks_give_the_expectation_value_of_a_larger_chain():
    """<S|W|S> from the blocks of a 16-site chi=32 chain equals the overlap of S with W applied to S (materialised, no rounding)."""
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS, _sweeps as sw
    from syngular.variational import DMRG
    import bench
    xs, ws = bench.make_chain(21, n=16, chi=32, chiw=8)
    X, W = MPS.from_sites(xs), MPO.from_sites(ws)
    right = DMRG.right_blocks(W, X)
    left = DMRG.left_blocks(W, X)
    # close the chain at bond 2: L_1 (b, v, b') with the remaining site-2.. block R_2 (a, w, a')
    val = float((left[1] * right[2]).sum().item())
    prod = [sw.site_mpo_mps(x, w) for x, w in zip(X.sites, W.sites)]
    ref = float(sw.overlap(prod, list(X.sites)).item())
    assert abs(val - ref) <= 1e-10 * abs(ref)
