"""GPU: batched overlaps and batched apply+round against the oracle, state by state."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand_chain(rng, n, d, chi, phys=1):
    b = [1] + [min(chi, d ** (phys * min(k, n - k))) for k in range(1, n)] + [1]
    return [rng.normal(size=(b[k],) + (d,) * phys + (b[k + 1],)) / np.sqrt(b[k] * d) for k in range(n)]


def test_batched_overlap_matches_oracle():
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    from oracle import ref_numpy as R
    rng = np.random.default_rng(5)
    B, n, d, chi = 6, 10, 2, 8
    As = [rand_chain(rng, n, d, chi) for _ in range(B)]
    Bs = [rand_chain(rng, n, d, chi) for _ in range(B)]
    got = BMPS.from_states(As).overlap(BMPS.from_states(Bs)).cpu().numpy()
    ref = np.array([R.overlap(a, b) for a, b in zip(As, Bs)])
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))


def test_batched_apply_round_matches_oracle():
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R
    rng = np.random.default_rng(6)
    B, n, d, chi, chiw, dim = 4, 8, 2, 8, 4, 6
    Xs = [rand_chain(rng, n, d, chi) for _ in range(B)]
    W = rand_chain(rng, n, d, chiw, phys=2)
    out = BMPS.from_states(Xs).apply_round([sw.as_core(w) for w in W], dim)
    for b in range(B):
        ref = R.round_qr([R.site_mpo_mps(x, w) for x, w in zip(Xs[b], W)], dim)
        got = [c[b].cpu().numpy() for c in out.sites]
        assert [c.shape for c in got] == [c.shape for c in ref]
        dr, dg = R.to_dense(ref), R.to_dense(got)
        assert np.max(np.abs(dr - dg)) < 1e-10 * np.max(np.abs(dr))


def test_batched_apply_round_svd_matches_oracle():
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(8)
    B, n, d, chi, chiw, target = 5, 8, 2, 8, 4, 6
    Xs = [rand_chain(rng, n, d, chi) for _ in range(B)]
    W = rand_chain(rng, n, d, chiw, phys=2)
    out = BMPS.from_states(Xs).apply_round_svd([sw.as_core(w) for w in W], target, chunk=2)     # chunk < B: exercises the stitching
    for b in range(B):
        ref, _, _ = S.apply_round_svd(Xs[b], W, target)
        got = [c[b].cpu().numpy() for c in out.sites]
        assert [c.shape for c in got] == [c.shape for c in ref]
        dr, dg = R.to_dense(ref), R.to_dense(got)
        assert np.max(np.abs(dr - dg)) < 1e-10 * np.max(np.abs(dr))
