"""GPU: the spectral-projection eigen-subspace solver (csrc/purify.cu) against numpy, and the density-matrix rounding sweep that
uses it against the textbook-SVD oracle and against the Jacobi path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _psd(rng, n, decay):
    """Symmetric PSD test matrix with a smooth, gapped-everywhere spectrum lam_i = exp(-decay i / n) (like a bond's squared
    singular values) and a Haar-random eigenbasis."""
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    lam = np.exp(-decay * np.arange(n) / n)
    return (q * lam) @ q.T, q, lam


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("n,ne,decay", [(64, 32, 6.0), (256, 128, 10.0), (512, 256, 12.0), (512, 100, 8.0), (384, 256, 9.0), (512, 192, 20.0),
                                        (1024, 512, 14.0), (128, 64, 3.0), (128, 32, 4.0), (160, 96, 5.0), (288, 32, 8.0)])
def test_dominant_subspace_matches_eigh(n, ne, decay, fused):
    from syngular_b200 import ops
    rng = np.random.default_rng(n + ne)
    A, q, lam = _psd(rng, n, decay)
    U, info = ops.dominant_subspace(torch.from_numpy(A).cuda(), ne, sp2_iters=56, ns_iters=30, fused=fused)
    h = info.cpu().numpy()
    Un = U.cpu().numpy()
    assert abs(h[0] - ne) < 1e-9 * ne and abs(h[1] - ne) < 1e-9 * ne, h
    assert h[4] < 1e-13 and np.max(np.abs(Un.T @ Un - np.eye(ne))) < 1e-13
    P = q[:, :ne] @ q[:, :ne].T
    assert np.max(np.abs(Un @ Un.T - P)) < 1e-11
    assert abs(h[2] - lam[:ne].sum()) < 1e-12 * lam.sum() and abs(h[5] - lam.sum()) < 1e-12 * lam.sum()
    assert abs(h[3] - np.linalg.norm(A)) < 1e-12 * np.linalg.norm(A)


def test_no_gap_is_reported_not_hidden():
    """Rank 40 < ne = 64: there is no projector of trace 64 to converge to; the info vector says so and the sweep falls back."""
    from syngular_b200 import ops
    rng = np.random.default_rng(1)
    B = rng.normal(size=(128, 40))
    A = B @ B.T
    for fused in (False, True):
        U, info = ops.dominant_subspace(torch.from_numpy(A).cuda(), 64, sp2_iters=44, ns_iters=22, fused=fused, sp2_max=60, ns_max=30)
        h = info.cpu().numpy()
        assert not (abs(h[0] - 64) < 1e-9 * 64 and abs(h[1] - 64) < 1e-9 * 64 and h[4] < 1e-12)


@pytest.mark.parametrize("n,chi,chiw", [(14, 16, 4), (16, 32, 8), (16, 64, 8), (16, 64, 2)])
def test_sweep_with_projection_solver_matches_oracle(n, chi, chiw):
    import bench
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    X, W = bench.make_chain(5, n=n, chi=chi, chiw=chiw)
    ref, spectra, discarded = S.apply_round_svd(X, W, chi)
    Xd, Wd = [sw.as_core(c) for c in X], [sw.as_core(c) for c in W]
    saved = sw.PURIFY_MIN_N
    try:
        sw.PURIFY_MIN_N = chi                      # every plateau bond (Gram matrices 2 chi x 2 chi) takes the projection solver
        sw.PURIFY_STATS.update(taken=0, fallback=0)
        out, trunc = sw.apply_round_dm(Xd, Wd, chi)
        taken = sw.PURIFY_STATS["taken"]
        sw.PURIFY_MIN_N = 0
        out_j, _ = sw.apply_round_dm(Xd, Wd, chi)
    finally:
        sw.PURIFY_MIN_N = saved
    # chi_W = 2 makes a steeply decaying spectrum (lambda_cut ~ 1e-7 .. 1e-11 lambda_0): the accuracy guard must send those bonds to
    # Cholesky + Jacobi; the other chains have C2-like spectra and take the projection solver
    assert taken >= 4 or chiw == 2
    got, got_j = [c.cpu().numpy() for c in out], [c.cpu().numpy() for c in out_j]
    dense_ref = R.to_dense(ref)
    scale = np.max(np.abs(dense_ref))
    assert np.max(np.abs(R.to_dense(got) - dense_ref)) < 2e-11 * scale           # north_star bound is 1e-10: keep a margin
    assert np.max(np.abs(R.to_dense(got) - R.to_dense(got_j))) < 2e-11 * scale
    sig, keep, disc = trunc.host()
    for k in range(n - 1):
        kk = ref[k].shape[-1]
        assert keep[k] == kk
        assert np.max(np.abs(sig[k][:kk] - spectra[k][:kk])) < 1e-10 * spectra[k][0]
        assert abs(disc[k] - discarded[k]) < 1e-10 * spectra[k][0] ** 2 * len(spectra[k])


@pytest.mark.parametrize("na,b", [(3, 16), (20, 64), (256, 256)])
def test_env_sandwich_matches_two_gemms(na, b):
    """csrc/env.cu against the two strided GEMMs it replaces and against numpy (einsum of the restated oracle)."""
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    rng = np.random.default_rng(na + b)
    l, i, o, r = 16, 2, 2, 16
    D = r * b
    P1 = torch.from_numpy(rng.normal(size=(na, i, r, D))).cuda()
    W = torch.from_numpy(rng.normal(size=(l, i, o, r))).cuda()
    assert ops.env_sandwich_fits(l, i, o, r, b)
    Z = torch.empty((na * l, l, i, b), dtype=torch.float64, device="cuda")
    ops.env_sandwich(P1, W, Z, na, b)
    Zg = sw._w_sandwich_gemms(P1, W, na, i, b, l, o, r, D)
    scale = Zg.abs().max().item()
    assert (Z - Zg).abs().max().item() < 1e-12 * scale
    if na <= 20:
        w, p1 = W.cpu().numpy(), P1.cpu().numpy().reshape(na, i, r, r, b)
        ref = np.einsum("aloqb,mjoq->almjb", np.einsum("lior,airqb->aloqb", w, p1), w).reshape(na * l, l, i, b)
        assert np.max(np.abs(Z.cpu().numpy() - ref)) < 1e-12 * scale


def test_symmetric_environment_build_matches_full_gemm():
    """Block-lower GEMMs + env_mirror against the full last GEMM of the environment update, on a chain whose plateau triggers it."""
    import bench
    from syngular.tensor import _sweeps as sw
    X, W = bench.make_chain(9, n=16, chi=128, chiw=4)
    Xd, Wd = [sw.as_core(c) for c in X], [sw.as_core(c) for c in W]
    saved = sw.ENV_SYMMETRIC_BLOCK
    try:
        sw.ENV_SYMMETRIC_BLOCK = 32
        E1 = sw.right_environments(Xd, Wd)
        sw.ENV_SYMMETRIC_BLOCK = 0
        E0 = sw.right_environments(Xd, Wd)
    finally:
        sw.ENV_SYMMETRIC_BLOCK = saved
    assert any(e.shape[0] >= 512 for e in E1[1:-1])
    for a, b in zip(E1[1:-1], E0[1:-1]):
        assert (a - b).abs().max().item() <= 1e-12 * b.abs().max().item()


@pytest.mark.parametrize("m,q,cols", [(512, 256, 4096), (256, 128, 300), (1024, 64, 64), (128, 64, 200), (160, 96, 128), (256, 32, 40)])
def test_polar_truncation_step_is_the_same_projection_as_householder(m, q, cols):
    """qrt_step through the fused Newton-Schulz kernel against numpy's QR of the leading columns (gauge-invariant: Q Q^T L)."""
    from syngular.tensor import _sweeps as sw
    from syngular_b200 import ops
    rng = np.random.default_rng(m + q)
    L = rng.normal(size=(m, cols))
    Ld = torch.from_numpy(L).cuda()
    assert ops.orthonormalize_columns_fits(m, q)
    Q, info = ops.orthonormalize_columns(Ld[:, :q])
    assert float(info[4].item()) < 1e-13
    Qs, S = sw.qrt_step(Ld, q)
    Qn, Sn = Qs.cpu().numpy(), S.cpu().numpy()
    assert np.max(np.abs(Qn.T @ Qn - np.eye(q))) < 1e-12
    qr, _ = np.linalg.qr(L[:, :q])
    assert np.max(np.abs(Qn @ Sn - qr @ (qr.T @ L))) < 1e-11 * np.max(np.abs(L))


def test_polar_truncation_step_falls_back_on_dependent_columns():
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(3)
    L = rng.normal(size=(256, 200))
    L[:, 5] = L[:, 2] * 0.5                                   # rank-deficient leading block: Newton-Schulz cannot converge
    Ld = torch.from_numpy(L).cuda()
    Q, S = sw.qrt_step(Ld, 128)
    Qn, Sn = Q.cpu().numpy(), S.cpu().numpy()
    assert np.max(np.abs(Qn.T @ Qn - np.eye(128))) < 1e-12
    assert np.max(np.abs((Qn @ Sn)[:, :128] - L[:, :128])) < 1e-11 * np.max(np.abs(L))


@pytest.mark.parametrize("n,chi,chiw,cutoff", [(16, 128, 8, 0.05), (18, 128, 16, 0.1)])
def test_relative_cutoff_stays_on_the_projection_solver(n, chi, chiw, cutoff):
    """north_star "fuse the singular-value cutoff": with cutoff > 0 the bonds still take the spectral-projection solver; the kept rank comes
    from the spectrum of U^T A U.  Against the textbook oracle with the same cutoff: ranks, dense tensor, kept spectra, discarded weights."""
    from syngular.tensor import _sweeps as sw
    from oracle import ref_numpy as R, svd_numpy as S
    import bench
    X, W = bench.make_chain(11, n=n, chi=chi, chiw=chiw)
    ref, spectra, discarded = S.apply_round_svd(X, W, chi, cutoff)
    Xd, Wd = [sw.as_core(c) for c in X], [sw.as_core(c) for c in W]
    sw.PURIFY_STATS.update(taken=0, fallback=0)
    out, trunc = sw.apply_round_dm(Xd, Wd, chi, cutoff=cutoff)
    assert sw.PURIFY_STATS["taken"] >= 2
    assert [tuple(c.shape) for c in out] == [tuple(c.shape) for c in ref]
    dense_ref = R.to_dense(ref)
    assert np.max(np.abs(R.to_dense([c.cpu().numpy() for c in out]) - dense_ref)) < 1e-10 * np.max(np.abs(dense_ref))
    sig, keep, disc = trunc.host()
    assert min(keep[6:-6]) < chi                            # the cutoff really cuts below chi_max on the plateau (ragged bonds)
    for k in range(n - 1):
        kk = ref[k].shape[-1]
        assert keep[k] == kk
        assert np.max(np.abs(np.sort(sig[k])[::-1][:kk] - spectra[k][:kk])) < 1e-10 * spectra[k][0]
        assert abs(disc[k] - discarded[k]) < 1e-10 * float(np.sum(spectra[k] ** 2))


@pytest.mark.parametrize("m,k", [(64, 32), (128, 64), (160, 96), (512, 256), (1024, 512)])
def test_complex_hermitian_projection_solver_matches_eigh(m, k):
    """syn_dominant_subspace_c128: planar complex Hermitian in, planar basis out; the fused kernel forms only the even rows of the embedded
    products.  Projector U U^H against LAPACK's, orthonormality, info doubles in the embedded convention (traces count twice), and the
    plain embedded route (the same kernel without the structure) for the same matrix."""
    import torch
    from syngular_b200 import ops, cplx
    rng = np.random.default_rng(m + k)
    Z = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
    Q, _ = np.linalg.qr(Z)
    lam = np.concatenate([np.exp(-rng.uniform(0.0, 4.0, size=k)), 1e-3 * np.exp(-rng.uniform(0.0, 6.0, size=m - k))])
    H = (Q * lam) @ Q.conj().T
    H = 0.5 * (H + H.conj().T)
    assert ops.dominant_subspace_c128_fits(m, k)
    Hre, Him = torch.from_numpy(np.ascontiguousarray(H.real)).cuda(), torch.from_numpy(np.ascontiguousarray(H.imag)).cuda()
    Ure, Uim, info = ops.dominant_subspace_c128(Hre, Him, k)
    U = Ure.cpu().numpy() + 1j * Uim.cpu().numpy()
    h = info.cpu().numpy()
    w, V = np.linalg.eigh(H)
    P = V[:, -k:] @ V[:, -k:].conj().T
    assert np.max(np.abs(U.conj().T @ U - np.eye(k))) < 1e-12
    assert np.max(np.abs(U @ U.conj().T - P)) < 1e-10
    assert abs(h[0] - 2 * k) < 2e-9 * k and abs(h[1] - 2 * k) < 2e-9 * k and h[4] < 1e-12
    assert abs(h[2] - 2 * w[-k:].sum()) < 1e-11 * w.sum() and abs(h[5] - 2 * w.sum()) < 1e-11 * w.sum()
    # the unstructured embedded route gives the same space
    Ve, info_e = ops.dominant_subspace(cplx.embed(cplx.Cx(Hre, Him)), 2 * k, sp2_max=90, ns_max=60)
    Ue = cplx.unembed_columns(Ve, m, k)
    Ue = Ue.re.cpu().numpy() + 1j * Ue.im.cpu().numpy()
    assert np.max(np.abs(Ue @ Ue.conj().T - U @ U.conj().T)) < 1e-10
    assert int(info_e.cpu()[7]) % 1000 == int(h[7]) % 1000           # the same number of SP2 steps: the iterates agree


def test_host_streaming_survives_a_rolled_back_site():
    """A steeply decaying spectrum (chi_W = 2) makes the accuracy guard reject projection results one site late: the sweep rolls back and
    re-solves those sites with Jacobi.  Cores are streamed to the host only once their verdict is in, so the host buffers must still hold
    exactly the cores the sweep returns."""
    import bench
    from syngular.tensor import _sweeps as sw
    X, W = bench.make_chain(5, n=16, chi=64, chiw=2)
    Xd, Wd = [sw.as_core(c) for c in X], [sw.as_core(c) for c in W]
    bufs = [torch.zeros(64 * 2 * 64, dtype=torch.float64).pin_memory() for _ in range(16)]
    saved = sw.PURIFY_MIN_N
    try:
        sw.PURIFY_MIN_N = 64
        sw.PURIFY_STATS.update(taken=0, fallback=0)
        out, _ = sw.apply_round_dm(Xd, Wd, 64, host_out=bufs)
        fallbacks = sw.PURIFY_STATS["fallback"]
    finally:
        sw.PURIFY_MIN_N = saved
    torch.cuda.synchronize()
    assert fallbacks >= 1
    for k, c in enumerate(out):
        assert torch.equal(bufs[k][: c.numel()], c.reshape(-1).cpu()), k
