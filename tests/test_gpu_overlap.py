"""GPU: the fused transfer-matrix kernel (csrc/overlap.cu, syn_overlap_batched_f64) against the numpy oracle and against the
GEMM-per-site route, through the C ABI.  Bit-for-bit agreement is not expected (different summation order): 1e-12 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def chain(rng, bonds, d):
    b = [1] + list(bonds) + [1]
    return [rng.normal(size=(b[k], d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(len(b) - 1)]


def oracle_overlap(a, b):
    from oracle import ref_numpy as R
    return R.overlap(a, b)


def stack(states):
    return [torch.from_numpy(np.stack([s[k] for s in states])).cuda() for k in range(len(states[0]))]


def capped(n, d, chi):
    return [min(chi, d ** min(k, n - k)) for k in range(1, n)]


CASES = [
    # (name, bonds of A, bonds of B, d, batch)
    ("c4_like", capped(12, 2, 64), capped(12, 2, 64), 2, 5),
    ("ragged_odd_bonds", [3, 5, 7, 9, 6, 1, 2], [2, 7, 11, 4, 3, 5, 1], 3, 4),
    ("different_widths", [8, 33, 64, 17, 2], [5, 40, 31, 64, 8], 2, 3),
    ("d4_chi32", capped(8, 4, 32), capped(8, 4, 32), 4, 3),
    ("d1", [4, 9, 4], [3, 3, 3], 1, 2),
    ("persistent_grid", [2, 4, 8, 4, 2], [2, 4, 8, 4, 2], 2, 333),     # batch > #SM: every CTA walks several states
    ("long_chain_is_split", [3] * 69, [2] * 69, 2, 2),                 # 70 sites > SYN_OVERLAP_MAX_SITES: E_out -> E_in
]


@pytest.mark.parametrize("name,ba,bb,d,batch", CASES, ids=[c[0] for c in CASES])
def test_fused_overlap_matches_oracle(name, ba, bb, d, batch):
    from syngular_b200 import ops
    rng = np.random.default_rng([c[0] for c in CASES].index(name))
    As = [chain(rng, ba, d) for _ in range(batch)]
    Bs = [chain(rng, bb, d) for _ in range(batch)]
    a, b = stack(As), stack(Bs)
    assert ops.overlap_fits(a, b)
    n0 = ops.lib.syn_launch_count()
    got = ops.overlap_batched(a, b).reshape(batch).cpu().numpy()
    assert ops.lib.syn_launch_count() - n0 == (len(a) + 63) // 64
    ref = np.array([oracle_overlap(x, y) for x, y in zip(As, Bs)])
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), (got, ref)


def test_shared_chain_and_single_chain_route():
    """One chain shared by the whole batch (batch stride 0) and the unbatched route `A | B` takes (syngular.tensor._sweeps.overlap)."""
    from syngular_b200 import ops
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(11)
    bonds = capped(9, 2, 16)
    As = [chain(rng, bonds, 2) for _ in range(7)]
    Bsh = chain(rng, bonds, 2)
    a, b = stack(As), stack([Bsh])
    got = ops.overlap_batched(a, b).reshape(7).cpu().numpy()
    ref = np.array([oracle_overlap(x, Bsh) for x in As])
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    one = float(sw.overlap([sw.as_core(c) for c in As[3]], [sw.as_core(c) for c in Bsh]).item())
    assert abs(one - ref[3]) <= 1e-12 * abs(ref[3])


def test_wide_bonds_take_the_gemm_route_and_agree():
    from syngular_b200 import ops
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    rng = np.random.default_rng(12)
    bonds = [2, 4, 80, 4, 2]
    As = [chain(rng, bonds, 2) for _ in range(3)]
    Bs = [chain(rng, bonds, 2) for _ in range(3)]
    a, b = stack(As), stack(Bs)
    assert not ops.overlap_fits(a, b)
    with pytest.raises(Exception):
        ops.overlap_batched(a, b)                      # the ABI refuses loudly, it does not fall back by itself
    got = BMPS(a).overlap(BMPS(b)).cpu().numpy()
    ref = np.array([oracle_overlap(x, y) for x, y in zip(As, Bs)])
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_full_size_c4_fused_equals_gemm_route():
    """BASELINE configs[3] shapes (N=32, d=2, chi=64), 512 pairs: fused kernel vs the GEMM-per-site route, state by state."""
    from syngular_b200 import ops
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    import bench
    bonds = bench.capped_bonds(32, 2, 64)[1:-1]
    A = BMPS.random(512, (2,) * 32, bonds, seed=3)
    B = BMPS.random(512, (2,) * 32, bonds, seed=4)
    fused = A.overlap(B)
    ops.OVERLAP_FUSED = False
    try:
        gemm = A.overlap(B)
    finally:
        ops.OVERLAP_FUSED = True
    scale = float(gemm.abs().max().item())
    assert float((fused - gemm).abs().max().item()) <= 1e-12 * scale
    nn = A.norms2()
    assert bool((nn > 0).all())
