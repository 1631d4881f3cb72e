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


@pytest.mark.parametrize("n,ne,B", [(128, 64, 5), (64, 32, 300), (96, 32, 7), (128, 96, 3), (96, 64, 150)])
def test_batched_projection_kernel_matches_eigh(n, ne, B):
    """csrc/purify_batched.cu: one CTA per problem.  Projector U U^T against the top-ne eigenprojector of LAPACK, orthonormality, and the
    info doubles (trace, Frobenius norm, kept weight, |A|_F, tr A)."""
    from syngular_b200 import ops
    rng = np.random.default_rng(n + ne + B)
    A = np.empty((B, n, n))
    for b in range(B):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        lam = np.concatenate([np.exp(-rng.uniform(0.0, 4.0, size=ne)), 1e-3 * np.exp(-rng.uniform(0.0, 6.0, size=n - ne))])
        A[b] = (Q * lam) @ Q.T
        A[b] = 0.5 * (A[b] + A[b].T)
    assert ops.dominant_subspace_batched_fits(n, ne)
    U, info = ops.dominant_subspace_batched(torch.from_numpy(A).cuda(), ne)
    U, h = U.cpu().numpy(), info.cpu().numpy()
    for b in range(B):
        w, V = np.linalg.eigh(A[b])
        P = V[:, -ne:] @ V[:, -ne:].T
        assert np.max(np.abs(U[b].T @ U[b] - np.eye(ne))) < 1e-12
        assert np.max(np.abs(U[b] @ U[b].T - P)) < 1e-10
        assert abs(h[b, 0] - ne) < 1e-9 * ne and abs(h[b, 1] - ne) < 1e-9 * ne and h[b, 4] < 1e-12
        assert abs(h[b, 2] - w[-ne:].sum()) < 1e-11 * w.sum()
        assert abs(h[b, 3] - np.linalg.norm(A[b])) < 1e-12 * np.linalg.norm(A[b]) and abs(h[b, 5] - np.trace(A[b])) < 1e-12 * np.trace(A[b])


def test_batched_projection_kernel_reports_a_missing_gap():
    """No gap at the cut (a degenerate pair straddles it): the iteration cannot reach a projector of trace ne and the info doubles say so."""
    from syngular_b200 import ops
    from syngular_b200.batched import _rejected
    rng = np.random.default_rng(3)
    n, ne = 64, 32
    A = np.empty((3, n, n))
    for b in range(3):
        Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
        lam = np.sort(np.exp(-rng.uniform(0.0, 3.0, size=n)))[::-1].copy()
        if b == 1:
            lam[ne] = lam[ne - 1]                      # eigenvalues number ne-1 and ne coincide
        A[b] = (Q * lam) @ Q.T
        A[b] = 0.5 * (A[b] + A[b].T)
    U, info = ops.dominant_subspace_batched(torch.from_numpy(A).cuda(), ne)
    assert list(_rejected(info.cpu().numpy(), ne, False)) == [1]


def test_batched_apply_round_svd_through_the_projection_kernel():
    """Bonds with rows = 64 keeping 32 go through the batched projection kernel; state by state against the oracle, and against the
    Jacobi-only route."""
    from syngular_b200 import batched
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    rng = np.random.default_rng(18)
    B, n, d, chi, chiw, target = 3, 14, 2, 32, 4, 32
    Xs = [rand_chain(rng, n, d, chi) for _ in range(B)]
    W = rand_chain(rng, n, d, chiw, phys=2)
    Wd = [sw.as_core(w) for w in W]
    batched.PROJECTION_STATS.update(taken=0, fallback=0)
    out = BMPS.from_states(Xs).apply_round_svd(Wd, target)
    assert batched.PROJECTION_STATS["taken"] >= B           # at least one bond of every state went through the kernel
    batched.PROJECTION_BATCHED = False
    try:
        jac = BMPS.from_states(Xs).apply_round_svd(Wd, target)
    finally:
        batched.PROJECTION_BATCHED = True
    for b in range(B):
        ref, _, _ = S.apply_round_svd(Xs[b], W, target)
        got = [c[b].cpu().numpy() for c in out.sites]
        assert [c.shape for c in got] == [c.shape for c in ref]
        dr, dg, dj = R.to_dense(ref), R.to_dense(got), R.to_dense([c[b].cpu().numpy() for c in jac.sites])
        assert np.max(np.abs(dr - dg)) < 1e-10 * np.max(np.abs(dr))
        assert np.max(np.abs(dj - dg)) < 1e-10 * np.max(np.abs(dr))


@pytest.mark.parametrize("rin,rout", [(8, 8), (4, 8), (16, 2), (2, 4), (16, 16)])
def test_small_core_kernel_matches_einsum(rin, rout):
    """csrc/smallcore.cu: Y[q][ro][x] = sum_ri W[ro][ri] X[q][ri][x] with strided X and a two-level ro index on the output."""
    from syngular_b200 import ops
    rng = np.random.default_rng(rin * 17 + rout)
    Q, L = 37, 50
    X = rng.normal(size=(Q, rin, L))
    Wm = rng.normal(size=(rout, rin))
    ref = np.einsum("or,qrx->qox", Wm, X)
    Xd, Wd = torch.from_numpy(X).cuda(), torch.from_numpy(Wm).cuda()
    Y = torch.empty((Q, rout, L), dtype=torch.float64, device="cuda")
    ops.apply_small_core(Xd, Wd, Y, Q=Q, L=L, x_q=rin * L, x_r=L, x_l=1, y_q=rout * L, y_ro=(0, L, rout), y_l=1)
    assert np.max(np.abs(Y.cpu().numpy() - ref)) < 1e-13 * np.max(np.abs(ref))
    # output with ro = (o, r) split around x: Y2[q][o][x][r], r = rout / 2 values innermost
    ro_r = 2
    Y2 = torch.empty((Q, rout // ro_r, L, ro_r), dtype=torch.float64, device="cuda")
    ops.apply_small_core(Xd, Wd, Y2, Q=Q, L=L, x_q=rin * L, x_r=L, x_l=1, y_q=rout * L, y_ro=(L * ro_r, 1, ro_r), y_l=ro_r)
    want2 = ref.reshape(Q, rout // ro_r, ro_r, L).transpose(0, 1, 3, 2)
    assert np.max(np.abs(Y2.cpu().numpy() - want2)) < 1e-13 * np.max(np.abs(ref))
    # the small matrix read in place as the transpose of a (rin x rout) tensor
    Wt = torch.from_numpy(np.ascontiguousarray(Wm.T)).cuda()
    Y3 = torch.empty((Q, rout, L), dtype=torch.float64, device="cuda")
    ops.apply_small_core(Xd, Wt, Y3, Q=Q, L=L, x_q=rin * L, x_r=L, x_l=1, y_q=rout * L, y_ro=(0, L, rout), y_l=1, w=(rout, rin, 1, rout))
    assert np.max(np.abs(Y3.cpu().numpy() - ref)) < 1e-13 * np.max(np.abs(ref))


def test_batched_apply_round_svd_small_core_path_equals_the_gemm_path():
    """chi_W = 4, d = 2 cores (8 x 8 small matrices) take the streaming kernel; the same sweep with GEMMs must give the same state."""
    from syngular_b200 import batched
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R
    rng = np.random.default_rng(28)
    B, n, d, chi, chiw, target = 3, 10, 2, 16, 4, 12
    Xs = [rand_chain(rng, n, d, chi) for _ in range(B)]
    W = rand_chain(rng, n, d, chiw, phys=2)
    Wd = [sw.as_core(w) for w in W]
    out = BMPS.from_states(Xs).apply_round_svd(Wd, target)
    batched.SMALL_CORE = False
    try:
        ref = BMPS.from_states(Xs).apply_round_svd(Wd, target)
    finally:
        batched.SMALL_CORE = True
    for b in range(B):
        dg = R.to_dense([c[b].cpu().numpy() for c in out.sites])
        dr = R.to_dense([c[b].cpu().numpy() for c in ref.sites])
        assert np.max(np.abs(dr - dg)) < 1e-11 * np.max(np.abs(dr))
