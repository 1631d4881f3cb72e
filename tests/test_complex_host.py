"""CPU: host logic of the complex128 path (syngular_b200/cplx.py and the complex-aware sweeps) against numpy, with the kernel
wrappers swapped for their numpy restatements (tests/cpu_ops.py).  This checks the ORCHESTRATION -- planar 4-GEMM complex
products with the real descriptor, the interleaved embedding behind the complex qrt, the pairing of the doubled Jacobi spectrum
(including degenerate singular values) -- not the kernels; tests/test_gpu_complex.py runs the same scenarios on the B200."""
import numpy as np
import pytest
import torch

import cpu_ops
import complex_cases as cc


@pytest.fixture(autouse=True)
def _emulated_kernels():
    with cpu_ops.patched():
        yield


def test_emulated_gemm_follows_the_two_level_index_contract():
    # the emulation itself: the site contraction issued by the product equals the oracle's einsum
    from oracle import ref_numpy as R
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(0)
    X, W = rng.normal(size=(3, 2, 4)), rng.normal(size=(2, 2, 3, 5))
    got = sw.site_mpo_mps(torch.from_numpy(X), torch.from_numpy(W)).numpy()
    assert np.allclose(got, R.site_mpo_mps(X, W), atol=1e-13)


@pytest.mark.parametrize("name", sorted(cc.CASES))
def test_complex_case(name):
    cc.CASES[name]()


def test_real_golden_cases_through_the_emulated_kernels():
    """The reference-pinned golden scenarios (tests/golden_cases.py) drive the same host code on the CPU: guards the refactor
    that made the sweeps operand-generic (real tensors / planar complex)."""
    import backends
    import golden_cases
    be = backends.ProductBackend()
    for case in golden_cases.ALL_CASES:
        case(be)


def test_truncation_step_host_logic_polar_and_fallback():
    """qrt_step: large full-rank blocks take the polar orthonormalisation, dependent leading columns fall back to Householder (emulated
    kernels: only the branch logic and the projection identity are checked here; the kernels run in tests/test_gpu_purify.py)."""
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    rng = np.random.default_rng(0)
    L = rng.normal(size=(256, 300))
    calls = {"polar": 0, "house": 0}
    polar, house = ops.orthonormalize_columns, ops.qrt

    def spy_polar(*a, **k):
        calls["polar"] += 1
        return polar(*a, **k)

    def spy_house(*a, **k):
        calls["house"] += 1
        return house(*a, **k)
    ops.orthonormalize_columns, ops.qrt = spy_polar, spy_house
    try:
        Q, S = sw.qrt_step(torch.from_numpy(L), 64)
        assert calls == {"polar": 1, "house": 0}
        qr, _ = np.linalg.qr(L[:, :64])
        assert np.max(np.abs(Q.numpy() @ S.numpy() - qr @ (qr.T @ L))) < 1e-11
        L2 = L.copy()
        L2[:, 3] = 2.0 * L2[:, 1]
        Q, S = sw.qrt_step(torch.from_numpy(L2), 64)
        assert calls == {"polar": 2, "house": 1}
        assert np.max(np.abs((Q.numpy() @ S.numpy())[:, :64] - L2[:, :64])) < 1e-11
        Q, S = sw.qrt_step(torch.from_numpy(L[:100]), 40)          # too few rows: Householder directly
        assert calls == {"polar": 2, "house": 2}
    finally:
        ops.orthonormalize_columns, ops.qrt = polar, house


@pytest.mark.parametrize("fail_calls", [(1,), (0, 1), (3,), (2, 5), tuple(range(20))])
def test_density_matrix_sweep_rolls_back_a_rejected_projection(fail_calls):
    """apply_round_dm reads the projection solver's verdict one site late; a rejected bond must roll the sweep back to that site and
    redo it (and everything queued after it) with the Jacobi path -- same state as the oracle whichever calls are rejected."""
    import bench
    from oracle import ref_numpy as R, svd_numpy as S
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    X, W = bench.make_chain(5, n=14, chi=16, chiw=4)
    ref, spectra, discarded = S.apply_round_svd(X, W, 16)
    dense_ref = R.to_dense(ref)
    real = ops.dominant_subspace
    calls = [0]

    def flaky(*a, **k):
        U, info = real(*a, **k)
        if calls[0] in fail_calls:
            info = info.clone()
            info[0] += 0.5                                     # trace off: the verdict rejects this bond
        calls[0] += 1
        return U, info
    saved = sw.PURIFY_MIN_N
    ops.dominant_subspace = flaky
    try:
        sw.PURIFY_MIN_N = 16
        sw.PURIFY_STATS.update(taken=0, fallback=0)
        out, trunc = sw.apply_round_dm([sw.as_core(c) for c in X], [sw.as_core(c) for c in W], 16)
        stats = dict(sw.PURIFY_STATS)
    finally:
        ops.dominant_subspace = real
        sw.PURIFY_MIN_N = saved
    assert stats["fallback"] >= 1 and [tuple(c.shape) for c in out] == [c.shape for c in ref]
    got = R.to_dense([c.numpy() for c in out])
    assert np.max(np.abs(got - dense_ref)) < 1e-10 * np.max(np.abs(dense_ref))
    sig, keep, disc = trunc.host()
    assert len(keep) == len(ref) - 1 and all(d is not None for d in disc)
    for k in range(len(ref) - 1):
        assert keep[k] == ref[k].shape[-1]
        assert abs(disc[k] - discarded[k]) < 1e-10 * spectra[k][0] ** 2 * len(spectra[k])
