"""CPU: invariants of the SVD-rounding oracle (oracle/svd_numpy.py).  The reference has no SVD, so this oracle is pinned by
mathematics instead of goldens: canonical form, spectra equal to the dense SVD of the unfolding, discarded weight == error^2,
optimality against the reference's QR truncation, and agreement of the density-matrix variant."""
import numpy as np

from oracle import ref_numpy as R
from oracle import svd_numpy as S


def rand_chain(rng, n, d, chi, phys=1):
    b = [1] + [chi] * (n - 1) + [1]
    if phys == 1:
        return [rng.normal(size=(b[k], d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(n)]
    return [rng.normal(size=(b[k], d, d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(n)]


def test_round_svd_invariants():
    rng = np.random.default_rng(0)
    cores = rand_chain(rng, 6, 3, 9)
    dense = R.to_dense(cores)
    out, spectra, discarded = S.round_svd(cores, 4)
    assert [c.shape for c in out] == [(1, 3, 3), (3, 3, 4), (4, 3, 4), (4, 3, 4), (4, 3, 3), (3, 3, 1)]
    for c in out[:-1]:
        L = c.reshape(-1, c.shape[-1])
        assert np.allclose(L.T @ L, np.eye(L.shape[1]), atol=1e-13)
    # first bond: spectrum == dense SVD of the first unfolding
    s_dense = np.linalg.svd(dense.reshape(3, -1), compute_uv=False)
    assert np.allclose(spectra[0][:3], s_dense, rtol=1e-12)
    err2 = np.sum((R.to_dense(out) - dense) ** 2)
    assert err2 <= sum(discarded) * (1 + 1e-9) + 1e-20       # TT-rounding bound: error^2 <= sum of discarded weights
    # lossless when chi_max is large enough
    out2, _, disc2 = S.round_svd(cores, 100)
    assert np.allclose(R.to_dense(out2), dense, atol=1e-13) and max(disc2) < 1e-24
    # SVD rounding beats (or ties) the reference's QR truncation at equal bond
    qr = R.round_qr(cores, 4)
    assert err2 <= np.sum((R.to_dense(qr) - dense) ** 2) + 1e-15


def test_density_matrix_variant_agrees():
    rng = np.random.default_rng(1)
    X = rand_chain(rng, 6, 2, 6)
    W = rand_chain(rng, 6, 2, 3, phys=2)
    a, sa, da = S.apply_round_svd(X, W, 5)
    b, sb, db = S.apply_round_density_matrix(X, W, 5)
    assert np.allclose(R.to_dense(a), R.to_dense(b), atol=1e-11)
    for x, y in zip(sa, sb):
        k = min(len(x), len(y), 5)
        assert np.allclose(x[:k], y[:k], rtol=1e-9)
    assert np.allclose(da, db, atol=1e-12)


def test_cutoff_and_rank_deficiency():
    x = np.arange(4 ** 4, dtype=float).reshape(4, 4, 4, 4)
    X = R.MPS.dense(x, (4, 4, 4))
    out, spectra, _ = S.round_svd(X.sites, 4, cutoff=1e-10)
    assert [c.shape[-1] for c in out[:-1]] == [2, 2, 2]       # arange tensors have TT-rank 2
    assert np.allclose(R.to_dense(out), x, atol=1e-9)
