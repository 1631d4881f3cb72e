"""CPU: host-side decisions of the batched SVD sweep (syngular_b200/batched.py) -- the per-member verdict on the projection kernel's info doubles
and the `fits` predicates of the C ABI that need no GPU."""
import numpy as np


def _info(ne, tr=None, f2=None, dev=1e-15, idem=0.0, lift=8, sp2=40, ns=12):
    h = np.zeros(8)
    h[0] = ne if tr is None else tr
    h[1] = ne if f2 is None else f2
    h[2], h[3], h[4], h[5], h[6] = 1.0, 1.0, dev, 1.1, idem
    h[7] = sp2 + 1000 * ns + 1e6 * lift
    return h


def test_rejected_members_are_exactly_those_that_fail_the_single_chain_verdict():
    from syngular_b200.batched import _rejected
    from syngular.tensor import _sweeps as sw
    ne = 64
    rows = [_info(ne),                                   # accepted
            _info(ne, tr=ne + 1e-3),                     # trace off: no projector of rank ne (no gap at the cut)
            _info(ne, dev=1e-9),                         # basis not orthonormal
            _info(ne, idem=1e-6),                        # not idempotent
            _info(ne, lift=sw.PURIFY_MAX_LIFT + 1),      # cut too deep in the spectrum for the accuracy bound
            _info(ne, lift=sw.PURIFY_MAX_LIFT),          # just inside
            _info(ne, f2=np.nan)]                        # diverged
    h = np.stack(rows)
    assert list(_rejected(h, ne, False)) == [1, 2, 3, 4, 6]
    # at the edge of the null space (kept rank = structural rank) deeper cuts are allowed
    h2 = np.stack([_info(ne, lift=sw.PURIFY_MAX_LIFT_RANK_GAP), _info(ne, lift=sw.PURIFY_MAX_LIFT_RANK_GAP + 1)])
    assert list(_rejected(h2, ne, True)) == [1]
    # and the verdict agrees with the single-chain function member by member
    for k, row in enumerate(rows):
        assert sw._projection_verdict(row, ne, False)[0] == (k not in (1, 2, 3, 4, 6))


def test_fits_predicates_need_no_gpu():
    from syngular_b200 import ops
    assert ops.dominant_subspace_batched_fits(128, 64) and ops.dominant_subspace_batched_fits(64, 32) and ops.dominant_subspace_batched_fits(128, 96)
    assert not ops.dominant_subspace_batched_fits(128, 128) and not ops.dominant_subspace_batched_fits(256, 64)
    assert not ops.dominant_subspace_batched_fits(100, 32) and not ops.dominant_subspace_batched_fits(64, 16)
    assert ops.small_core_fits(8, 8) and ops.small_core_fits(2, 16) and not ops.small_core_fits(3, 8) and not ops.small_core_fits(32, 8)
    assert ops.dominant_subspace_c128_fits(1024, 512) and not ops.dominant_subspace_c128_fits(1000, 500) and not ops.dominant_subspace_c128_fits(64, 64)


def test_gauss_mix_is_reproducible():
    import torch
    from syngular_b200 import batched
    a = batched._gauss(16, 4, torch.device("cpu"))
    batched._GAUSS.clear()
    b = batched._gauss(16, 4, torch.device("cpu"))
    assert torch.equal(a, b) and a.shape == (16, 4) and float(a.abs().max()) > 0
