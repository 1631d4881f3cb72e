"""GPU: Householder qrt and one-sided Jacobi kernels against numpy/LAPACK on the same inputs (gauge-invariant checks)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).normal(size=shape)).cuda()


@pytest.mark.parametrize("m,n,q", [(8, 5, 3), (16, 16, 16), (64, 40, 40), (100, 300, 37), (512, 4096, 256), (512, 300, 256),
                                   (6, 3, 5), (3, 7, 9), (1, 4, 1), (700, 33, 33), (2000, 64, 48), (256, 16, 16)])
def test_qrt_matches_reference_step(m, n, q):
    from syngular_b200 import ops
    from oracle.ref_numpy import qrt as qrt_ref
    A = _rand((m, n), m * 7 + n)
    Q, S = ops.qrt(A, q)
    qk = min(q, m)
    assert tuple(Q.shape) == (m, qk) and tuple(S.shape) == (qk, n)
    Qn, Sn, An = Q.cpu().numpy(), S.cpu().numpy(), A.cpu().numpy()
    assert np.max(np.abs(Qn.T @ Qn - np.eye(qk))) < 1e-13
    Qr, Sr = qrt_ref(An, q)
    # gauge-invariant: the projected matrix Q S (== Q Q^T A)
    ref = Qr @ Sr
    assert np.max(np.abs(Qn @ Sn - ref)) < 1e-12 * max(1.0, np.max(np.abs(ref)))
    # S must equal Q^T A
    assert np.max(np.abs(Sn - Qn.T @ An)) < 1e-12 * np.max(np.abs(An)) * np.sqrt(m)


def test_qrt_strided_transposed_and_batched():
    from syngular_b200 import ops
    R = _rand((40, 96), 3)                 # a right unfolding (l, d*r); QR of R^T as in right_orthonormalization
    Q, S = ops.qrt(R.t(), 40)
    Qn, Sn = Q.cpu().numpy(), S.cpu().numpy()
    assert np.max(np.abs(Qn @ Sn - R.cpu().numpy().T)) < 1e-12
    A = _rand((5, 70, 30), 4)
    Qb, Sb = ops.qrt(A, 12)
    for b in range(5):
        Qn, Sn, An = Qb[b].cpu().numpy(), Sb[b].cpu().numpy(), A[b].cpu().numpy()
        Qr, _ = np.linalg.qr(An[:, :12])
        assert np.max(np.abs(Qn @ Sn - Qr @ (Qr.T @ An))) < 1e-12
    Rr = ops.qr_r(A[0])
    Rn = np.linalg.qr(A[0].cpu().numpy(), mode="r")
    assert np.max(np.abs(np.abs(Rr.cpu().numpy()) - np.abs(Rn))) < 1e-12
    T = ops.copy_strided(A[1].t())
    assert torch.equal(T, A[1].t().contiguous())


def test_qrt_rank_deficient_and_zero_columns():
    from syngular_b200 import ops
    x = np.arange(64 * 48, dtype=np.float64).reshape(64, 48)      # rank 2, as in the reference's arange tests
    A = torch.from_numpy(x).cuda()
    Q, S = ops.qrt(A, 16)
    Qn, Sn = Q.cpu().numpy(), S.cpu().numpy()
    assert np.max(np.abs(Qn.T @ Qn - np.eye(16))) < 1e-12
    assert np.max(np.abs(Qn @ Sn - x)) < 1e-9 * np.max(x)
    Z = torch.zeros((10, 6), dtype=torch.float64, device="cuda")
    Q, S = ops.qrt(Z, 4)
    assert np.max(np.abs(Q.cpu().numpy().T @ Q.cpu().numpy() - np.eye(4))) < 1e-14 and S.abs().max().item() == 0.0


@pytest.mark.parametrize("n", [1, 2, 3, 8, 33, 64, 100, 128, 130, 192, 256, 300, 384, 512, 1024])
def test_jacobi_rows_singular_values(n):
    from syngular_b200 import ops
    G0 = _rand((n, n), n)
    G = G0.clone()
    ops.jacobi_rows(G)
    Ut, sigma, info, winfo = ops.jacobi_finalize(G, chi_max=n)
    s_ref = np.linalg.svd(G0.cpu().numpy(), compute_uv=False)
    s = sigma.cpu().numpy()
    assert np.max(np.abs(s - s_ref)) < 1e-12 * s_ref[0]
    U = Ut.cpu().numpy()
    keep = int(info[0].item())
    assert keep == n
    assert np.max(np.abs(U @ U.T - np.eye(n))) < 1e-11
    # rows are right singular vectors of G0: |U G0^T G0 U^T| = diag(s^2)
    M = U @ (G0.cpu().numpy().T @ G0.cpu().numpy()) @ U.T
    assert np.max(np.abs(M - np.diag(s_ref ** 2))) < 1e-10 * s_ref[0] ** 2


def test_jacobi_symmetric_psd_sqrt_mode_and_cut():
    from syngular_b200 import ops
    rng = np.random.default_rng(0)
    B = rng.normal(size=(96, 300))
    A = B @ B.T
    G = torch.from_numpy(A).cuda()
    ops.jacobi_rows(G)
    Ut, sigma, info, winfo = ops.jacobi_finalize(G, chi_max=40, cutoff=0.0, rank_tol=1e-7, sqrt_mode=True)
    s_ref = np.linalg.svd(B, compute_uv=False)
    assert np.max(np.abs(sigma.cpu().numpy() - s_ref)) < 1e-10 * s_ref[0]
    assert int(info[0].item()) == 40
    assert abs(winfo[0].item() - np.sum(s_ref[40:] ** 2)) < 1e-9 * np.sum(s_ref ** 2)
    U = Ut.cpu().numpy()[:40]
    Uref = np.linalg.svd(B)[0][:, :40]
    assert np.max(np.abs(U.T @ U - Uref @ Uref.T)) < 1e-9       # same dominant subspace (projector)
    # batched + rank deficient: rank 5 matrix -> keep == 5
    L = rng.normal(size=(3, 24, 5))
    Gb = torch.from_numpy(np.einsum("bik,bjk->bij", L, L)).cuda()
    ops.jacobi_rows(Gb)
    Ut, sigma, info, winfo = ops.jacobi_finalize(Gb, chi_max=24, rank_tol=1e-7, sqrt_mode=True)
    assert info[:, 0].tolist() == [5, 5, 5]


def test_misc_kernels():
    from syngular_b200 import ops
    from oracle import ref_numpy as R
    rng = np.random.default_rng(1)
    A = rng.normal(size=(3, 2, 4, 5)); B = rng.normal(size=(2, 2, 4, 6))
    for first, last, (a, b) in [(False, False, (A, B)), (True, False, (A[:1], B[:1])), (False, True, (A[..., :1], B[..., :1]))]:
        out = ops.add_site(torch.from_numpy(np.ascontiguousarray(a)).cuda(), torch.from_numpy(np.ascontiguousarray(b)).cuda(), first, last)
        assert np.array_equal(out.cpu().numpy(), R.site_add(a, b, first, last))
    out = ops.kron_site(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda())
    assert np.max(np.abs(out.cpu().numpy() - R.site_kron(A, B))) < 1e-15
    x = torch.from_numpy(rng.normal(size=(7, 3, 1))).cuda()
    ss = ops.sumsq(x)
    assert abs(ss.item() - float((x ** 2).sum().item())) < 1e-12
    y = x.clone()
    ops.scale_rsqrt_(y, ss)
    assert abs(float((y ** 2).sum().item()) - 1.0) < 1e-14
