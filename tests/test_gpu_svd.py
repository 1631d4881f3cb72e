"""GPU: SVD bond rounding (Householder + one-sided Jacobi kernels) against the numpy oracle oracle/svd_numpy.py.
Gauge-invariant comparisons only: dense tensors, per-bond singular spectra, kept ranks and discarded weights; FP64
tolerance 1e-10 relative (north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rand_chain(rng, n, d, chi, phys=1, cap=True):
    b = [1] + [chi] * (n - 1) + [1]
    if cap:
        for k in range(1, n):
            b[k] = min(chi, d ** (phys * k), d ** (phys * (n - k)))
    if phys == 1:
        return [rng.normal(size=(b[k], d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(n)]
    return [rng.normal(size=(b[k], d, d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(n)]


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,d,chi,target", [(6, 3, 9, 4), (8, 2, 12, 5), (5, 4, 16, 16), (10, 2, 20, 3), (4, 6, 30, 7)])
def test_round_svd_matches_oracle(n, d, chi, target):
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(n * 100 + chi)
    cores = rand_chain(rng, n, d, chi, cap=False)
    ref, spectra, discarded = S.round_svd(cores, target)
    out, trunc = sw.round_svd([sw.as_core(c) for c in cores], target)
    got = [c.cpu().numpy() for c in out]
    assert [c.shape for c in got] == [c.shape for c in ref]
    assert rel(R.to_dense(got), R.to_dense(ref)) < 1e-10
    sig, keep, disc = trunc.host()
    for k in range(n - 1):
        m = len(spectra[k])
        assert rel(sig[k][:m], spectra[k]) < 1e-10 * 1.0 or np.max(np.abs(sig[k][:m] - spectra[k])) < 1e-12 * spectra[k][0]
        assert keep[k] == ref[k].shape[-1]
        assert abs(disc[k] - discarded[k]) < 1e-10 * max(np.sum(spectra[k] ** 2), 1e-300)
    for c in got[:-1]:
        L = c.reshape(-1, c.shape[-1])
        assert np.max(np.abs(L.T @ L - np.eye(L.shape[1]))) < 1e-11


@pytest.mark.parametrize("n,d,chi,chiw,target", [(6, 2, 6, 3, 5), (8, 2, 8, 4, 8), (5, 3, 5, 2, 4), (12, 2, 16, 4, 16)])
def test_apply_round_density_matrix_matches_oracle(n, d, chi, chiw, target):
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(n * 10 + chi)
    X = rand_chain(rng, n, d, chi)
    W = rand_chain(rng, n, d, chiw, phys=2)
    ref, spectra, discarded = S.apply_round_svd(X, W, target)
    out, trunc = sw.apply_round_dm([sw.as_core(c) for c in X], [sw.as_core(c) for c in W], target)
    got = [c.cpu().numpy() for c in out]
    assert [c.shape for c in got] == [c.shape for c in ref]
    assert rel(R.to_dense(got), R.to_dense(ref)) < 1e-10
    sig, keep, disc = trunc.host()
    for k in range(n - 1):
        kk = ref[k].shape[-1]
        assert np.max(np.abs(sig[k][:kk] - spectra[k][:kk])) < 1e-10 * spectra[k][0]
    # and the textbook path on the materialised product gives the same state
    prod = [sw.site_mpo_mps(sw.as_core(x), sw.as_core(w)) for x, w in zip(X, W)]
    out2, _ = sw.round_svd(prod, target)
    assert rel(R.to_dense([c.cpu().numpy() for c in out2]), R.to_dense(ref)) < 1e-10


def test_svd_mode_through_the_api():
    import syngular.tensor as st
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(3)
    X = rand_chain(rng, 7, 2, 8)
    W = rand_chain(rng, 7, 2, 4, phys=2)
    st.set_rounding("svd")
    try:
        Xg, Wg = MPS.from_sites(X), MPO.from_sites(W)
        Y = Wg @ Xg                                    # min_bond = min(2, 2) = 2 from the capped edge bonds
        ref, _, _ = S.apply_round_svd(X, W, 2)
        assert rel(Y.to_tensor().real, R.to_dense(ref)) < 1e-10
        assert (Xg >> 3) is Xg                         # guard: 3 >= min(bond_shape) = 2 -> the operand itself
        X2 = rand_chain(rng, 6, 3, 8, cap=False)
        X2g = MPS.from_sites(X2)
        Z = X2g >> 3
        ref2, _, _ = S.round_svd(X2, 3)
        assert rel(Z.to_tensor().real, R.to_dense(ref2)) < 1e-10
        assert Z.truncation is not None and Z.bond_shape == X2g.bond_shape     # stale metadata rule kept
        # rank-deficient input: arange tensors have TT rank 2 -> SVD mode shrinks the bond
        A = MPS(np.arange(4 ** 4, dtype=float).reshape(4, 4, 4, 4), bond_shape=(4, 4, 4)).decompose()
        B = A >> 3
        assert [s.shape[-1] for s in B.sites[:-1]] == [2, 2, 2]
        assert rel(B.to_tensor().real, np.arange(4 ** 4, dtype=float).reshape(4, 4, 4, 4)) < 1e-10
    finally:
        st.set_rounding("qr")


def test_site_contractions_and_overlap_vs_oracle():
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R
    rng = np.random.default_rng(11)
    X = rng.normal(size=(5, 3, 7)); W = rng.normal(size=(4, 3, 2, 6)); A = rng.normal(size=(3, 2, 4, 5)); B = rng.normal(size=(2, 4, 3, 6))
    assert rel(sw.site_mpo_mps(sw.as_core(X), sw.as_core(W)).cpu().numpy(), R.site_mpo_mps(X, W)) < 1e-13
    assert rel(sw.site_mpo_mpo(sw.as_core(A), sw.as_core(B)).cpu().numpy(), R.site_mpo_mpo(A, B)) < 1e-13
    P = rand_chain(rng, 9, 2, 10); Q = rand_chain(rng, 9, 2, 7)
    got = sw.overlap([sw.as_core(c) for c in P], [sw.as_core(c) for c in Q]).item()
    assert abs(got - R.overlap(P, Q)) < 1e-12 * abs(R.overlap(P, Q)) + 1e-300


def test_eigh_gram_with_and_without_the_cholesky_start():
    """eigh_gram (Jacobi on the shifted Cholesky factor, or on the Gram matrix itself) against the singular values of the factor."""
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(21)
    B = rng.normal(size=(256, 1024))
    A = B @ B.T
    ref = np.linalg.svd(B, compute_uv=False)
    old_chol = sw.CHOLESKY_MIN_N
    try:
        for flag in (0, 129):
            sw.CHOLESKY_MIN_N = flag
            Ut, sigma, info, winfo = sw.eigh_gram(torch.from_numpy(A.copy()).cuda(), 100, 0.0, 3.2e-7)
            assert int(info[1]) == 256                                   # the convergence verdict of the rows kernel travels in info[1]
            assert np.max(np.abs(sigma.cpu().numpy() - ref)) < 1e-10 * ref[0]
            U = Ut.cpu().numpy()
            assert np.max(np.abs(U @ U.T - np.eye(256))) < 1e-11
            assert np.max(np.abs(U[:100] @ A @ U[:100].T - np.diag(ref[:100] ** 2))) < 1e-9 * ref[0] ** 2
    finally:
        sw.CHOLESKY_MIN_N = old_chol


@pytest.mark.parametrize("n,rank", [(1, 1), (5, 5), (64, 64), (100, 37), (128, 128), (200, 200), (512, 512), (320, 100), (130, 1), (1024, 700)])
def test_shifted_cholesky_factor(n, rank):
    """syn_chol_upper_f64: G + shift I = B^T B with B upper triangular, also for singular G (bond "inflation", short chains)."""
    from syngular_b200 import ops
    rng = np.random.default_rng(n + rank)
    F = rng.normal(size=(n, rank)) * np.exp(-np.arange(rank) / (rank / 6.0 + 1.0))[None, :]
    G = F @ F.T
    B, shift = ops.chol_upper(torch.from_numpy(G.copy()).cuda())
    B, delta = B.cpu().numpy(), float(shift.item())
    assert 0.0 < delta <= 4.0 * n * 2.3e-16 * np.max(np.diag(G))
    assert np.max(np.abs(np.tril(B, -1))) == 0.0
    assert np.max(np.abs(B.T @ B - G - delta * np.eye(n))) < 2e-13 * np.max(np.diag(G))


def test_shifted_cholesky_factor_batched():
    """One launch, one CTA per problem (n <= 128): every member gets its own shift and factor."""
    from syngular_b200 import ops
    rng = np.random.default_rng(5)
    nb, n = 37, 96
    F = rng.normal(size=(nb, n, 40)) * rng.uniform(0.1, 10.0, size=(nb, 1, 1))
    G = F @ F.transpose(0, 2, 1)
    B, shift = ops.chol_upper(torch.from_numpy(G.copy()).cuda())
    B, shift = B.cpu().numpy(), shift.cpu().numpy()
    for b in range(nb):
        dmax = np.max(np.diag(G[b]))
        assert 0.0 < shift[b] <= 4.0 * n * 2.3e-16 * dmax
        assert np.max(np.abs(np.tril(B[b], -1))) == 0.0
        assert np.max(np.abs(B[b].T @ B[b] - G[b] - shift[b] * np.eye(n))) < 2e-13 * dmax


@pytest.mark.parametrize("n,rank", [(256, 256), (512, 512), (384, 150), (512, 256)])
def test_eigh_gram_on_cholesky_factor(n, rank):
    """eigh_gram through the shifted Cholesky factor (n > 128) against numpy and against the direct Jacobi-on-G path:
    singular values to 1e-10 sigma_0, orthonormal eigenvectors, kept rank = numerical rank for singular Gram matrices."""
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    rng = np.random.default_rng(7 * n + rank)
    F = rng.normal(size=(n, rank)) * np.exp(-np.arange(rank) / (rank / 8.0))[None, :]
    A = F @ F.T
    ref = np.zeros(n)
    ref[:rank] = np.linalg.svd(F, compute_uv=False)
    keep_ref = int(np.sum(ref[:n // 2] > 3.2e-7 * ref[0]))
    sweeps = {}
    old = sw.CHOLESKY_MIN_N
    try:
        for flag in (129, 0):
            sw.CHOLESKY_MIN_N = flag
            Ut, sigma, info, winfo = sw.eigh_gram(torch.from_numpy(A.copy()).cuda(), n // 2, 0.0, 3.2e-7)
            sweeps[flag] = ops.jacobi_sweeps_used()[0]
            keep = int(info[0].item())
            assert keep == keep_ref
            assert np.max(np.abs(sigma.cpu().numpy()[:keep] - ref[:keep])) < 1e-10 * ref[0]
            U = Ut.cpu().numpy()[:keep]
            assert np.max(np.abs(U @ U.T - np.eye(keep))) < 1e-11
            assert np.max(np.abs(U @ A @ U.T - np.diag(ref[:keep] ** 2))) < 1e-9 * ref[0] ** 2
            assert abs(winfo[0].item() - np.sum(ref[keep:] ** 2)) < 1e-10 * ref[0] ** 2
    finally:
        sw.CHOLESKY_MIN_N = old
    assert sweeps[129] <= sweeps[0]


def test_apply_round_dm_with_a_cutoff_below_the_gram_resolution_follows_the_textbook_oracle():
    """The density-matrix sweep drops singular values below rank_tol = 3.2e-7 sigma_0 (squared values are noise there).  A user cutoff
    below that must still be honoured: the call takes the textbook route.  State = a dominant part + a 1e-9 perturbation with its own
    bond space: with cutoff 1e-11 the oracle keeps both parts, with the default cutoff 0 the density-matrix mode keeps the dominant one."""
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(77)
    n, d = 8, 2

    def chain(chi, phys=1):
        b = [1] + [min(chi, d ** (phys * min(k, n - k))) for k in range(1, n)] + [1]
        return [rng.normal(size=(b[k],) + (d,) * phys + (b[k + 1],)) / np.sqrt(b[k] * d) for k in range(n)]
    A, Bp, W = chain(3), chain(2), chain(2, phys=2)
    Bp[3] = Bp[3] * 1e-9
    X = [R.site_add(a, b, k == 0, k == n - 1) for k, (a, b) in enumerate(zip(A, Bp))]
    Xd, Wd = [sw.as_core(x) for x in X], [sw.as_core(w) for w in W]
    ref, spectra, _ = S.apply_round_svd(X, W, 64, cutoff=1e-11)
    out, trunc = sw.apply_round_dm(Xd, Wd, 64, cutoff=1e-11)
    got = [c.cpu().numpy() for c in out]
    assert [c.shape for c in got] == [c.shape for c in ref]
    assert got[n // 2 - 1].shape[-1] == 7                    # middle bond: the dominant part's 3 x 2 plus one direction of the perturbation (4.9e-11 sigma_0)
    dr, dg = R.to_dense(ref), R.to_dense(got)
    assert np.max(np.abs(dr - dg)) < 1e-10 * np.max(np.abs(dr))
    # the small singular values themselves are resolved (relative accuracy of the Jacobi SVD on the triangular factor)
    mid = n // 2
    s_ref = spectra[mid - 1]
    s_got = np.sort(trunc.sigma[mid - 1].cpu().numpy())[::-1]
    k = got[mid - 1].shape[-1]
    assert np.max(np.abs(s_got[:k] - s_ref[:k]) / s_ref[:k]) < 1e-6
    # default cutoff: the density-matrix mode declares the 1e-9 part noise
    out0, _ = sw.apply_round_dm(Xd, Wd, 64)
    assert out0[n // 2 - 1].shape[-1] == 6


def test_syn_mul_host_out_streams_the_result_cores():
    """syn.mul(..., host_out=pinned buffers): the cores are copied to the host from inside the sweep (overlapped with it on the optimized
    route, plainly on the standard route); the buffers must hold exactly the cores the call returns."""
    import torch
    import syngular as syn
    from syngular.tensor import MatrixProductState as MPS, MatrixProductOperator as MPO
    import bench
    X, W = bench.make_chain(5, n=16, chi=64, chiw=8)
    for mode, bond in (("optimized", 64), ("standard", 32)):
        bufs = [torch.zeros(64 * 2 * 64, dtype=torch.float64).pin_memory() for _ in range(16)]
        Y = syn.mul(MPO.from_sites(W), MPS.from_sites(X), mode=mode, bond=bond, host_out=bufs)
        torch.cuda.synchronize()
        for k, c in enumerate(Y.sites):
            assert torch.equal(bufs[k][: c.numel()], c.reshape(-1).cpu()), (mode, k)


def test_contract_carry_small_core_path_equals_the_gemm_path():
    """chi_W = 8, d = 2: the carry x MPO-core contraction (16 x 16 core against s b >= 4096 columns) takes the streaming kernel, in the
    Python-driven sweeps and in the C chain call; same results as the GEMM route."""
    import torch
    from syngular.tensor import _sweeps as sw
    import bench
    X, W = bench.make_chain(9, n=14, chi=64, chiw=8)
    Xd, Wd = [sw.as_core(x) for x in X], [sw.as_core(w) for w in W]
    T = torch.randn(64, 8, 64, dtype=torch.float64, device="cuda")
    a = sw.contract_carry(T, Xd[7], Wd[7])
    sw.SMALL_CORE = False
    try:
        b = sw.contract_carry(T, Xd[7], Wd[7])
        ref_svd, _ = sw.apply_round_dm(Xd, Wd, 64)
        ref_qr = sw.apply_round_qr_steps(Xd, Wd, 64)
    finally:
        sw.SMALL_CORE = True
    assert float((a - b).abs().max()) < 1e-13 * float(b.abs().max())
    got_svd, _ = sw.apply_round_dm(Xd, Wd, 64)
    got_qr = sw.apply_round_qr(Xd, Wd, 64)                     # the C chain call
    for x, y in zip(got_qr, ref_qr):
        assert float((x - y).abs().max()) < 1e-10
    n2 = float(sw.overlap(ref_svd, ref_svd).item())
    d2 = n2 + float(sw.overlap(got_svd, got_svd).item()) - 2 * float(sw.overlap(got_svd, ref_svd).item())
    assert abs(d2) < 1e-12 * n2
