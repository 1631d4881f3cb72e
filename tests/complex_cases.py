"""complex128 scenarios shared by the CPU host-logic test (kernels emulated, tests/test_complex_host.py) and the GPU test
(tests/test_gpu_complex.py): every case drives the PRODUCT code and compares with numpy / the oracle at 1e-10."""
import numpy as np
import torch

RTOL = 1e-10


def _dev():
    from syngular.tensor import _sweeps as sw
    return sw.device()


def _cx(a):
    from syngular_b200.cplx import Cx
    return Cx.from_torch(torch.from_numpy(np.ascontiguousarray(a)).to(_dev()))


def _np(c):
    return c.to_torch().cpu().numpy()


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def _crandn(rng, *shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


def haar(rng, n):
    q, r = np.linalg.qr(_crandn(rng, n, n))
    return q * (np.diag(r) / np.abs(np.diag(r)))


# ---- primitives ---------------------------------------------------------------------------------------------------
def case_matmul_and_views():
    from syngular_b200 import cplx
    rng = np.random.default_rng(1)
    A, B = _crandn(rng, 7, 5), _crandn(rng, 5, 9)
    assert _rel(_np(cplx.matmul(_cx(A), _cx(B))), A @ B) < 1e-13
    assert _rel(_np(cplx.matmul(_cx(A).h(), _cx(A))), A.conj().T @ A) < 1e-13
    C0 = _crandn(rng, 7, 9)
    out = _cx(C0)
    cplx.matmul(_cx(A), _cx(B), out=out, alpha=-0.5, beta=1.5)
    assert _rel(_np(out), -0.5 * A @ B + 1.5 * C0) < 1e-13


def case_qrt_projects_on_leading_columns():
    from syngular_b200 import cplx
    rng = np.random.default_rng(2)
    for m, n, q in ((12, 9, 4), (6, 10, 6), (8, 8, 8), (5, 9, 7)):
        A = _crandn(rng, m, n)
        Q, S = cplx.qrt(_cx(A), q)
        Qn, Sn = _np(Q), _np(S)
        k = min(q, m)
        assert Qn.shape == (m, k) and Sn.shape == (k, n)
        assert np.max(np.abs(Qn.conj().T @ Qn - np.eye(k))) < 1e-12
        Qr, _ = np.linalg.qr(A[:, :q])
        P = Qr[:, :k] @ Qr[:, :k].conj().T
        assert _rel(Qn @ Sn, P @ A) < RTOL                       # gauge-invariant: the projected matrix


def case_qrt_rank_deficient_columns():
    from syngular_b200 import cplx
    rng = np.random.default_rng(3)
    A = _crandn(rng, 8, 6)
    A[:, 1] = 0.0                                               # a null column: the completion is not a (q, Jq) pair a priori
    A[:, 3] = A[:, 0] * (0.3 - 0.4j)
    Q, S = cplx.qrt(_cx(A), 5)
    Qn, Sn = _np(Q), _np(S)
    assert np.max(np.abs(Qn.conj().T @ Qn - np.eye(5))) < 1e-12
    assert _rel((Qn @ Sn)[:, :5], A[:, :5]) < RTOL              # the leading columns are reproduced exactly


def _svd_projector(M, k):
    u, s, _ = np.linalg.svd(M, full_matrices=False)
    return u[:, :k] @ u[:, :k].conj().T, s


def case_svd_basis_generic():
    from syngular_b200 import cplx
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(4)
    for m, c, chi in ((6, 10, 3), (16, 16, 8), (40, 64, 12)):
        M = _crandn(rng, m, c)
        U, keep, sigma, disc = cplx.svd_basis(_cx(M), chi, 0.0, sw.eigh_gram)
        Un = _np(U)
        P, s = _svd_projector(M, chi)
        assert keep == chi and Un.shape == (m, chi)
        assert np.max(np.abs(Un.conj().T @ Un - np.eye(chi))) < 1e-12
        assert _rel(Un @ Un.conj().T, P) < 1e-9
        assert _rel(sigma.cpu().numpy()[:m], s) < 1e-9
        assert abs(float(disc) - float(np.sum(s[chi:] ** 2))) < 1e-9 * float(np.sum(s ** 2))


def case_svd_basis_degenerate_spectrum():
    """Bell / GHZ-like unfoldings: equal singular values, so the embedded eigenvectors of a cluster come out as an arbitrary real
    rotation of the (v, Jv) pairs and the even rows alone need not span the complex subspace."""
    from syngular_b200 import cplx
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(5)
    m, c = 8, 12
    u, v = haar(rng, m), haar(rng, c)
    s = np.array([1.0, 1.0, 1.0, 0.5, 0.5, 0.1, 0.0, 0.0])
    M = (u * s) @ v[:m]
    for chi in (3, 5, 6):
        U, keep, sigma, _ = cplx.svd_basis(_cx(M), chi, 0.0, sw.eigh_gram)
        Un = _np(U)
        assert keep == chi
        assert np.max(np.abs(Un.conj().T @ Un - np.eye(chi))) < 1e-12
        assert _rel(Un @ Un.conj().T, u[:, :chi] @ u[:, :chi].conj().T) < 1e-7        # Gram route: sqrt(eps)-level on a gap of 0.5
        assert np.max(np.abs(sigma.cpu().numpy()[:6] - s[:6])) < 1e-7
    # cutoff drops the numerically null directions
    U, keep, _, _ = cplx.svd_basis(_cx(M), 8, 1e-6, sw.eigh_gram)
    assert keep == 6


# ---- chain operators ----------------------------------------------------------------------------------------------
def _rand_chain(rng, bonds, d=2, mpo=False):
    b = [1] + list(bonds) + [1]
    if mpo:
        return [_crandn(rng, b[k], d, d, b[k + 1]) / np.sqrt(2 * b[k] * d) for k in range(len(b) - 1)]
    return [_crandn(rng, b[k], d, b[k + 1]) / np.sqrt(2 * b[k] * d) for k in range(len(b) - 1)]


def case_apply_gate_reference_mode_matches_oracle():
    from oracle import ref_numpy as R
    from syngular.tensor import MatrixProductState as MPS
    rng = np.random.default_rng(6)
    cores = _rand_chain(rng, (2, 4, 4, 2))
    X = MPS.from_sites(cores)
    for g, i in ((haar(rng, 2), 0), (haar(rng, 4).reshape(2, 2, 2, 2), 1), (haar(rng, 8).reshape((2,) * 6), 2), (haar(rng, 4).reshape(2, 2, 2, 2), 3)):
        Y = X.apply(g, i)
        ref = R.mps_apply(cores, g, i)
        assert [tuple(s.shape) for s in Y.sites] == [c.shape for c in ref]
        assert _rel(Y.to_tensor(), R.to_dense(ref)) < RTOL
    # a complex gate on a REAL chain makes only the touched cores complex
    Xr = MPS.from_sites([c.real.copy() for c in cores])
    Y = Xr.apply(np.array([[0, -1j], [1j, 0]]), 2)
    assert [bool(s.is_complex()) for s in Y.sites] == [False, False, True, False, False]
    assert _rel(Y.to_tensor(), R.to_dense(R.mps_apply([c.real.copy() for c in cores], np.array([[0, -1j], [1j, 0]]), 2))) < RTOL


def _dense_apply(psi, g, i, n):
    m = g.ndim // 2
    psi = psi.reshape((2,) * n)
    psi = np.tensordot(g, psi, axes=(list(range(m, 2 * m)), list(range(i, i + m))))
    return np.moveaxis(psi, list(range(m)), list(range(i, i + m))).reshape(-1)


def case_brickwall_circuit_of_haar_unitaries():
    """BASELINE configs[2] in miniature: |0..0>, brick-wall layers of Haar-random two-qubit unitaries, SVD split."""
    from syngular.quantum import Circuit
    rng = np.random.default_rng(7)
    n, depth = 6, 5
    structure, psi = [], np.zeros(2 ** n, dtype=complex)
    psi[0] = 1.0
    for layer in range(depth):
        for i in range(layer % 2, n - 1, 2):
            g = haar(rng, 4).reshape(2, 2, 2, 2)
            structure.append((g, i))
            psi = _dense_apply(psi, g, i, n)
    c = Circuit(n, structure=structure, chi_max=8)
    c.run()
    st = c.get().state
    assert [s.shape[2] for s in st.sites[:-1]] == [2, 4, 8, 4, 2]
    assert np.max(np.abs(c.get().to_tensor() - psi)) < RTOL
    assert abs((st.conj() | st) - 1.0) < RTOL                      # unitary evolution, exact bond: norm 1
    # truncated run against the same truncation done densely (sequential SVD split with numpy)
    c2 = Circuit(n, structure=structure, chi_max=4)
    c2.run()
    got = c2.get().to_tensor()
    assert np.linalg.norm(got) <= 1.0 + 1e-12
    assert _rel(got, _tebd_numpy(n, structure, 4)) < 1e-8


def _tebd_numpy(n, structure, chi):
    cores = [np.zeros((1, 2, 1), dtype=complex) for _ in range(n)]
    for c in cores:
        c[0, 0, 0] = 1.0
    for g, i in structure:
        a, b = cores[i], cores[i + 1]
        th = np.tensordot(a, b, axes=(2, 0))
        th = np.einsum("opij,lijr->lopr", g, th)
        l, r = th.shape[0], th.shape[3]
        u, s, vh = np.linalg.svd(th.reshape(l * 2, 2 * r), full_matrices=False)
        k = max(1, min(chi, int(np.sum(s > 3.2e-7 * s[0]))))
        cores[i] = u[:, :k].reshape(l, 2, k)
        cores[i + 1] = (u[:, :k].conj().T @ th.reshape(l * 2, 2 * r)).reshape(k, 2, r)
    T = cores[0]
    for c in cores[1:]:
        T = np.tensordot(T, c, axes=(T.ndim - 1, 0))
    return T.reshape(-1)


def case_bell_state_with_phase_gates():
    """Degenerate Schmidt values (1/sqrt2, 1/sqrt2) with complex amplitudes: H, S, CX, T, Y through the growing-bond split."""
    from syngular.quantum import Qbit, gate
    q = Qbit(3, chi_max=4)
    psi = np.zeros(8, dtype=complex)
    psi[0] = 1.0
    for op in ((gate.H, 0), (gate.S, 0), (gate.CX, 0, 1), (gate.T, 1), (gate.Y, 2), (gate.CX, 1, 2)):
        q = q @ op
        g = np.asarray(op[0])
        psi = _dense_apply(psi, g, op[1], 3)
    assert np.max(np.abs(q.to_tensor() - psi)) < RTOL
    # the reference's own mode (bond frozen at 2) reproduces the oracle's numbers with complex gates
    from oracle import ref_numpy as R
    q2, r2 = Qbit(3), R.Qbit(3)
    for op in ((gate.H, 0), (gate.S, 0), (gate.CX, 0, 1), (gate.T, 1), (gate.Y, 2)):
        q2 = q2 @ op
        r2 = r2 @ op
    assert _rel(q2.to_tensor(), r2.to_tensor()) < RTOL


def case_matmul_add_round_overlap_match_oracle():
    from oracle import ref_numpy as R
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    rng = np.random.default_rng(8)
    xs, ys, ws = _rand_chain(rng, (2, 4, 2)), _rand_chain(rng, (2, 3, 2)), _rand_chain(rng, (3, 3, 3), mpo=True)
    X, Y, W = MPS.from_sites(xs), MPS.from_sites(ys), MPO.from_sites(ws)
    Xr, Yr, Wr = R.MPS.from_sites(xs), R.MPS.from_sites(ys), R.MPO.from_sites(ws)
    assert abs((X | Y) - (Xr | Yr)) < RTOL * abs(Xr | Yr)                         # bilinear, no conjugation (MPS:116-129)
    # the reference's `+` assembles into a float64 np.zeros and silently drops imaginary parts (MPS:82-86); here: the direct sum
    assert _rel((X + Y).to_tensor(), R.to_dense(xs) + R.to_dense(ys)) < RTOL
    Z, Zr = W @ X, Wr @ Xr                                                       # site contraction + `>> min_bond` (QR truncation)
    assert [tuple(s.shape) for s in Z.sites] == [tuple(s.shape) for s in Zr.sites]
    assert _rel(Z.to_tensor(), R.to_dense(Zr.sites)) < RTOL
    sum_sites = [c.cpu().numpy() for c in (X + Y).sites]
    T, Tr = (X + Y) >> 3, R.MPS.from_sites(sum_sites) >> 3
    assert _rel(T.to_tensor(), R.to_dense(Tr.sites)) < RTOL
    X.left_orthonormalization()
    assert np.max(np.abs(X.left_orthogonality(1) - np.eye(X.sites[1].shape[2]))) < 1e-12
    assert _rel(X.to_tensor(), R.to_dense(xs)) < RTOL
    X.right_orthonormalization()
    assert np.max(np.abs(X.right_orthogonality(2) - np.eye(X.sites[2].shape[0]))) < 1e-12
    assert _rel(X.to_tensor(), R.to_dense(xs)) < RTOL
    assert abs(X[(1, 0, 1, 1)][0, 0] - R.to_dense(xs)[1, 0, 1, 1]) < RTOL
    nrm = MPS.from_sites(xs)
    nrm.left_orthonormalization()
    nrm.normalize()
    assert abs((nrm.conj() | nrm) - 1.0) < RTOL


def case_svd_rounding_of_a_complex_product():
    from oracle import svd_numpy as S, ref_numpy as R
    import syngular as syn
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    rng = np.random.default_rng(9)
    xs, ws = _rand_chain(rng, (2, 4, 4, 2)), _rand_chain(rng, (3, 3, 3, 3), mpo=True)
    Y = syn.mul(MPO.from_sites(ws), MPS.from_sites(xs), mode="optimized", bond=5)
    prod = [R.site_mpo_mps(x, w) for x, w in zip(xs, ws)]
    ref = _round_svd_numpy(prod, 5)
    assert [s.shape[2] for s in Y.sites[:-1]] == [c.shape[2] for c in ref[:-1]]
    assert _rel(Y.to_tensor(), R.to_dense(ref)) < 1e-8


def _round_svd_numpy(cores, chi):
    """Textbook rounding with conjugation (oracle/svd_numpy.round_svd is written for real cores)."""
    cores = [c.copy() for c in cores]
    for k in range(len(cores) - 1, 0, -1):
        c = cores[k]
        q, r = np.linalg.qr(c.reshape(c.shape[0], -1).T)
        cores[k] = q.T.reshape((-1,) + c.shape[1:])
        cores[k - 1] = np.tensordot(cores[k - 1], r.T, axes=(2, 0))
    for k in range(len(cores) - 1):
        c = cores[k]
        M = c.reshape(-1, c.shape[2])
        u, s, vh = np.linalg.svd(M, full_matrices=False)
        kk = max(1, min(chi, int(np.sum(s > 3.2e-7 * s[0]))))
        cores[k] = u[:, :kk].reshape(c.shape[0], c.shape[1], kk)
        cores[k + 1] = np.tensordot((u[:, :kk].conj().T @ M), cores[k + 1], axes=(1, 0))
    return cores




def case_large_complex_bond_takes_the_projection_solver():
    """A (128 x 160) complex unfolding cut to 32: the embedded problem is 256 x 256, so the fused spectral-projection kernel runs on
    the interleaved embedding and the complex basis is read off its even columns."""
    from syngular_b200 import cplx
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(12)
    m, c, chi = 128, 160, 32
    u, v = haar(rng, m), haar(rng, c)
    s = np.exp(-6.0 * np.arange(m) / m)
    M = (u * s) @ v[:m]
    U, keep, sigma, disc = cplx.svd_basis(_cx(M), chi, 0.0, sw.eigh_gram)
    assert sigma is None and keep == chi                          # the projection route was taken (no spectrum is formed)
    Un = _np(U)
    assert np.max(np.abs(Un.conj().T @ Un - np.eye(chi))) < 1e-12
    assert _rel(Un @ Un.conj().T, u[:, :chi] @ u[:, :chi].conj().T) < 1e-9
    assert abs(float(disc) - float(np.sum(s[chi:] ** 2))) < 1e-10 * float(np.sum(s ** 2))




def case_mpo_products_and_decompose_with_complex_cores():
    from oracle import ref_numpy as R
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS
    from syngular.variational import DMRG
    rng = np.random.default_rng(14)
    a, b = _rand_chain(rng, (2, 3, 2), mpo=True), _rand_chain(rng, (3, 2, 3), mpo=True)
    A, B = MPO.from_sites(a), MPO.from_sites(b)
    Ar, Br = R.MPO.from_sites(a), R.MPO.from_sites(b)
    C, Cr = A @ B, Ar @ Br                                   # site contraction + `>> min_bond` (MPO:278-289)
    assert [tuple(s.shape) for s in C.sites] == [tuple(s.shape) for s in Cr.sites]
    assert _rel(C.to_tensor(), R.to_dense(Cr.sites)) < RTOL
    assert abs(C[(1, 0, 1, 1), (0, 1, 1, 0)][0, 0] - Cr[(1, 0, 1, 1), (0, 1, 1, 0)][0, 0]) < RTOL * np.max(np.abs(R.to_dense(Cr.sites)))
    # TT decomposition of a dense complex tensor by the qrt step (MPS:298-319) reproduces it when the bonds are large enough
    t = _crandn(rng, 2, 3, 2, 2)
    X = MPS(t, bond_shape=(2, 6, 2)).decompose()
    assert _rel(X.to_tensor(), t) < RTOL
    assert _rel(X.to_tensor(), R.to_dense(R.MPS.dense(t, (2, 6, 2)).sites)) < RTOL
    # three-layer transfer blocks with complex cores (no conjugation, like the reference)
    xs = _rand_chain(rng, (2, 4, 2))
    got = DMRG.right_blocks(A, MPS.from_sites(xs))
    ref = R.dmrg_right_blocks(xs, a)
    for g, r in zip(got, ref):
        assert (g is None) == (r is None)
        if r is not None:
            assert _rel(g.cpu().numpy(), r) < RTOL


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}
