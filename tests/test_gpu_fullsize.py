"""GPU: the hot path at BASELINE.json's FULL sizes (configs[1]: N=64, d=2, chi=256, chi_W=16; configs[3]: N=32, chi=64 batch),
checked (i) against ONE full-size sweep of the like-for-like CPU oracle (about a minute of BLAS/LAPACK on the host cores: the last
test of this file) and (ii) through size-independent properties:
  * every rounded core is left-orthonormal;
  * sequential orthogonal projections:  |W X|^2 - |result|^2 == sum of the per-bond discarded weights   (SVD mode);
  * optimality: the SVD-rounded state is at least as close to W X as the reference's QR-rounded state;
  * idempotence: rounding the result again to the same bond changes nothing;
  * fused apply+round == materialise-then-round on a full-size sub-chain;
  * linearity / symmetry of the batched overlap."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import bench
    from syngular.tensor import _sweeps as sw
    X, W = bench.make_chain(2)
    return [sw.as_core(x) for x in X], [sw.as_core(w) for w in W]


def _norm2_of_product(X, W):
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    E = sw.right_environments(X, W)
    T = torch.ones((1, 1, 1), dtype=torch.float64, device=X[0].device)
    M0 = sw.contract_carry(T, X[0], W[0])                      # (1, o, D1)
    M2 = M0.reshape(M0.shape[1], M0.shape[2])
    return float(sw.gram_with_environment(M2, E[1], X[0].shape[2], W[0].shape[3]).diagonal().sum().item())


def _gram_err(core):
    from syngular_b200 import ops
    L = core.reshape(-1, core.shape[-1])
    G = ops.matmul(L.t(), L)
    return float((G - torch.eye(G.shape[0], dtype=torch.float64, device=G.device)).abs().max().item())


def test_c2_svd_round_invariants(c2):
    from syngular.tensor import _sweeps as sw
    X, W = c2
    out, trunc = sw.apply_round_dm(X, W, 256)
    assert [tuple(c.shape) for c in out][30] == (256, 2, 256)
    assert max(_gram_err(c) for c in out[:-1]) < 1e-11
    n2_exact = _norm2_of_product(X, W)
    n2_out = float(sw.overlap(out, out).item())
    _, keep, disc = trunc.host()
    assert abs((n2_exact - n2_out) - sum(disc)) < 1e-9 * n2_exact
    assert 0.0 < n2_out < n2_exact and max(keep) == 256
    # optimality against the reference-semantic QR truncation at the same bond: <result|WX> = |result|^2 for projections,
    # so the error is |WX|^2 - |result|^2: SVD must lose less
    out_qr = sw.apply_round_qr(X, W, 256)
    assert max(_gram_err(c) for c in out_qr[:-1]) < 1e-11
    n2_qr = float(sw.overlap(out_qr, out_qr).item())
    assert n2_exact - n2_out <= (n2_exact - n2_qr) * (1 + 1e-9)
    # idempotence of both roundings at the same bond
    again = sw.round_qr(out_qr, 256)
    ov = float(sw.overlap(again, out_qr).item())
    assert abs(ov - n2_qr) < 1e-10 * n2_qr
    again_svd, _ = sw.round_svd(out, 256)
    ov = float(sw.overlap(again_svd, out).item())
    assert abs(ov - n2_out) < 1e-9 * n2_out


def test_c2_fused_equals_materialised_on_a_full_size_subchain(c2):
    """Sites 0..11 of C2 reach the full plateau bond (256 x 16 = 4096): fused vs literal apply + QR round agree."""
    from syngular.tensor import _sweeps as sw
    X, W = c2
    Xs = [c for c in X[:11]] + [X[11][:, :, :1].contiguous()]
    Ws = [c for c in W[:11]] + [W[11][:, :, :, :1].contiguous()]
    fused = sw.apply_round_qr(Xs, Ws, 256)
    lit = sw.round_qr([sw.site_mpo_mps(x, w) for x, w in zip(Xs, Ws)], 256)
    a = float(sw.overlap(fused, fused).item()); b = float(sw.overlap(lit, lit).item()); c = float(sw.overlap(fused, lit).item())
    assert abs(a - b) < 1e-10 * a and abs(a - c) < 1e-10 * a


def test_c4_batched_overlap_properties():
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    import bench
    bonds = bench.capped_bonds(32, 2, 64)[1:-1]
    A = BMPS.random(256, (2,) * 32, bonds, seed=1)
    B = BMPS.random(256, (2,) * 32, bonds, seed=2)
    ab, ba = A.overlap(B), B.overlap(A)
    assert float((ab - ba).abs().max().item()) < 1e-12 * float(ab.abs().max().item())       # bilinear form is symmetric
    aa, bb = A.norms2(), B.norms2()
    assert bool((aa > 0).all()) and bool((ab * ab <= aa * bb * (1 + 1e-12)).all())            # Cauchy-Schwarz, state by state
    # single-chain path on one member agrees with the batched kernel
    got = A.state(7) | B.state(7)
    assert abs(got - ab[7].item()) < 1e-12 * abs(got)


# ---------------------------------------------------------------------------------------------------------------------
# the headline path against the oracle AT HEADLINE SIZE: one full C2 sweep of the like-for-like CPU algorithm
# (oracle/svd_numpy.StructuredDensityMatrixSweep, ~1 minute of BLAS/LAPACK on the host cores) versus the CUDA sweep
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_oracle_sweep():
    import bench
    from oracle import svd_numpy as S
    X, W = bench.make_chain(2)
    bench.use_all_host_threads()
    return S.apply_round_density_matrix_structured(X, W, 256)


def test_c2_svd_round_matches_the_oracle_at_full_size(c2, c2_oracle_sweep):
    """Per-bond kept ranks, kept singular spectra and discarded weights, the norm, the rounded STATE itself (overlap of the two
    results) and overlaps with a fixed probe state: 1e-10 relative (FP64 bar of BASELINE.json's north_star)."""
    import bench
    from syngular.tensor import _sweeps as sw
    X, W = c2
    ref_cores, ref_spec, ref_disc = c2_oracle_sweep
    sw.PURIFY_STATS.update(taken=0, fallback=0)
    capture = {31: None}
    out, trunc = sw.apply_round_dm(X, W, 256, capture=capture)
    assert sw.PURIFY_STATS["taken"] >= 40                      # the plateau bonds really took the spectral-projection solver
    sig, keep, disc = trunc.host()
    # kept rank = min(chi_max, numerical rank): the last bonds of the chain are rank-deficient (2^(64-k-1) < 256) and the oracle's
    # eigh returns roundoff-sized eigenvalues there, which its cutoff = 0 rule counts; the CUDA path drops them (rank_tol 3.2e-7)
    assert [int(k) for k in keep] == [min(256, int(np.count_nonzero(s > 3.2e-7 * s[0]))) for s in ref_spec]
    ref = [sw.as_core(c) for c in ref_cores]
    n_gg = float(sw.overlap(out, out).item()); n_cc = float(sw.overlap(ref, ref).item()); n_gc = float(sw.overlap(out, ref).item())
    assert abs(n_gg - n_cc) < 1e-10 * n_cc
    # |out_gpu - out_cpu|^2 / |out|^2 from three overlaps: the same STATE in any gauge.  The difference of O(1) numbers resolves
    # ~1e-15 at best, i.e. this bounds the distance itself by ~3e-8; the 1e-10 statements are the linear functionals below
    assert abs(n_gg + n_cc - 2.0 * n_gc) < 1e-14 * n_cc
    for k, (s_g, s_c, d_g, d_c) in enumerate(zip(sig, ref_spec, disc, ref_disc)):
        kk = keep[k]
        assert len(s_g) >= kk, k
        assert np.max(np.abs(np.sort(s_g)[::-1][:kk] - s_c[:kk])) < 1e-10 * s_c[0], k
        assert abs(d_g - d_c) < 1e-10 * float(np.sum(s_c ** 2)), k
    # overlaps with a fixed probe state (bond 8): a gauge-invariant linear functional of the rounded state
    Xp, _ = bench.make_chain(77, chi=8)
    probe = [sw.as_core(x) for x in Xp]
    n_pp = float(sw.overlap(probe, probe).item())
    o_g = float(sw.overlap(out, probe).item()); o_c = float(sw.overlap(ref, probe).item())
    assert abs(o_g - o_c) < 1e-10 * np.sqrt(n_cc * n_pp)
    # one plateau bond: the kept projector of the spectral-projection solver against LAPACK's eigh of the SAME Gram matrix
    A, U = capture[31]
    lam, V = np.linalg.eigh(A.cpu().numpy())
    Vk = V[:, -256:]
    Ug = U.cpu().numpy()
    gap = lam[-256] - lam[-257]
    dist = np.linalg.norm(Ug @ Ug.T - Vk @ Vk.T)
    assert gap > 0
    # Davis-Kahan: a backward-stable eigen-solver tilts the kept space by ~ eps |A| / gap; 100x head-room, and never worse than 1e-8
    assert dist < min(100 * 2.2e-16 * lam[-1] / gap * np.sqrt(256), 1e-8), (dist, gap / lam[-1])
