"""GPU: the whole-chain entry points of the C ABI (csrc/chain.cu: syn_apply_round_chain_f64, syn_round_chain_f64) against the numpy
oracle of the reference semantics (oracle/ref_numpy.py, pinned by the reference's goldens) and against the site-by-site host sweep."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _chain(seed, n, d, chi, chiw):
    import bench
    return bench.make_chain(seed, n=n, d=d, chi=chi, chiw=chiw)


@pytest.mark.parametrize("n,d,chi,chiw,dim", [(6, 2, 8, 4, 6), (8, 3, 9, 3, 20), (10, 2, 16, 4, 16), (5, 4, 6, 2, 100), (12, 2, 64, 8, 64)])
def test_apply_round_chain_matches_the_oracle(n, d, chi, chiw, dim):
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R
    X, W = _chain(n * 7 + d, n, d, chi, chiw)
    out = sw.apply_round_qr([sw.as_core(x) for x in X], [sw.as_core(w) for w in W], dim)
    ref = R.round_qr([R.site_mpo_mps(x, w) for x, w in zip(X, W)], dim)
    assert [tuple(c.shape) for c in out] == [tuple(c.shape) for c in ref]
    got = R.to_dense([c.cpu().numpy() for c in out])
    want = R.to_dense(ref)
    assert np.max(np.abs(got - want)) < 1e-10 * np.max(np.abs(want))


@pytest.mark.parametrize("n,d,chi,dim", [(6, 2, 8, 5), (7, 3, 12, 30), (9, 2, 16, 7)])
def test_round_chain_matches_the_oracle_mps_and_mpo(n, d, chi, dim):
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R
    X, W = _chain(n + 31 * d, n, d, chi, 4)
    for cores in (X, W):
        out = sw.round_qr([sw.as_core(c) for c in cores], dim)
        ref = R.round_qr(cores, dim)
        assert [tuple(c.shape) for c in out] == [tuple(c.shape) for c in ref]
        got = R.to_dense([c.cpu().numpy() for c in out])
        want = R.to_dense(ref)
        assert np.max(np.abs(got - want)) < 1e-10 * np.max(np.abs(want))


def test_chain_call_equals_the_site_by_site_sweep_at_plateau_size():
    """A C2-shaped sub-chain that reaches the 256 x 16 plateau: one library call vs the host-driven sweep (same kernels, same order)."""
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    import bench
    X, W = bench.make_chain(2)
    Xs = [sw.as_core(c) for c in X[:13]] + [sw.as_core(X[13][:, :, :1])]
    Ws = [sw.as_core(c) for c in W[:13]] + [sw.as_core(W[13][:, :, :, :1])]
    l0 = ops.lib.syn_launch_count()
    one = sw.apply_round_qr(Xs, Ws, 256)
    launches = ops.lib.syn_launch_count() - l0
    steps = sw.apply_round_qr_steps(Xs, Ws, 256)
    assert launches > 0
    for a, b in zip(one, steps):
        # the same kernels in the same order; the Newton-Schulz kernel accumulates its norms with floating-point atomics, so two runs
        # agree to roundoff, not bit for bit (the polar factor it converges to is unique: no gauge freedom between the two)
        assert a.shape == b.shape and float((a - b).abs().max().item()) < 1e-11
    rq = sw.round_qr(one, 128)
    saved = sw.CHAIN_CALLS
    try:
        sw.CHAIN_CALLS = False
        rs = sw.round_qr(one, 128)
    finally:
        sw.CHAIN_CALLS = saved
    for a, b in zip(rq, rs):
        assert a.shape == b.shape and float((a - b).abs().max().item()) < 1e-11


def test_chain_entry_points_reject_bad_arguments():
    from syngular_b200 import ops
    from syngular_b200._lib import SynError
    X = [torch.randn(1, 2, 3, dtype=torch.float64, device="cuda"), torch.randn(4, 2, 1, dtype=torch.float64, device="cuda")]      # bonds do not match
    W = [torch.randn(1, 2, 2, 2, dtype=torch.float64, device="cuda"), torch.randn(2, 2, 2, 1, dtype=torch.float64, device="cuda")]
    with pytest.raises(SynError):
        ops.apply_round_chain(X, W, 4)


def test_chain_call_repeats_the_sweep_when_a_speculative_step_is_rejected():
    """The chain call queues its Newton-Schulz truncation steps without waiting for their verdicts and reads them once at the end.  A plateau
    site whose leading columns are dependent (a duplicated bond column) fails that iteration: the call must notice it afterwards, repeat
    the sweep with immediate verdicts (Householder at that site) and return what the site-by-site sweep returns."""
    from syngular.tensor import _sweeps as sw
    import bench
    X, W = bench.make_chain(2)
    X = [np.array(c, copy=True) for c in X[:13]] + [np.array(X[13][:, :, :1], copy=True)]
    Ws = [sw.as_core(c) for c in W[:13]] + [sw.as_core(W[13][:, :, :, :1])]
    X[10][:, :, 1] = X[10][:, :, 0]                     # site 10 (256 x 2 x 256): bond column 1 duplicates column 0
    Xs = [sw.as_core(c) for c in X]
    one = sw.apply_round_qr(Xs, Ws, 256)
    steps = sw.apply_round_qr_steps(Xs, Ws, 256)
    for k, (a, b) in enumerate(zip(one, steps)):
        assert a.shape == b.shape and bool(torch.isfinite(a).all())
        if k < 10:                                      # before the deficient site both sweeps are the same kernels on the same data
            assert float((a - b).abs().max().item()) < 1e-10, k
    # at the deficient site the 16 dependent columns are replaced by completion directions, which are as arbitrary as LAPACK's in the
    # reference (the two Householder kernels complete differently): the columns that ARE determined agree, every core is left-orthonormal
    a, b = one[10].reshape(-1, 256), steps[10].reshape(-1, 256)
    assert float((a[:, :16] - b[:, :16]).abs().max().item()) < 1e-10
    for k in range(13):
        L = one[k].reshape(-1, one[k].shape[-1])
        assert float((L.t() @ L - torch.eye(L.shape[1], dtype=torch.float64, device=L.device)).abs().max()) < 1e-12, k
