"""complex128 on the FP64 kernels: planar (split re / im) tensors and the complex versions of the library calls.

The sm_100a library is real FP64 (DMMA GEMM, Householder qrt, one-sided Jacobi).  Gates of the quantum callers
(reference: quantum/gate.py:9-13 Y, T, S; SURVEY 8f-1, BASELINE configs[2]) make cores complex, so this module maps the
complex operations onto the same kernels:

  * a complex tensor is held PLANAR -- two contiguous float64 tensors of the logical shape -- so every two-level strided
    index of the real GEMM descriptor applies unchanged to both parts; a complex GEMM is 4 real GEMM launches
    (C_re = A_re B_re - A_im B_im, C_im = A_re B_im + A_im B_re; alpha/beta accumulate in the kernel);
  * the truncation step qrt runs on the real embedding [[a, -b], [b, a]] of every entry (rows AND columns interleaved):
    Householder QR of that matrix is the embedding of the complex QR, so the even columns of its Q are the complex basis;
  * the bond SVD diagonalises the embedded Hermitian Gram matrix (2m x 2m real symmetric, every eigenvalue doubled) with the
    Jacobi kernel; one vector per pair is kept and the result is polished by Newton-Schulz steps (complex GEMMs), which also
    repairs the pairing when singular values are degenerate (Bell / GHZ states).  Large bonds (embedded size >= 256) with a gap at the
    cut use the fused spectral-projection kernel on the interleaved embedding instead: its projector is an embedding by construction,
    so the complex basis is read off directly and the Jacobi size limit does not apply.

No arithmetic happens on the host or in PyTorch: torch is used for allocation, views, concatenation / interleaving copies and
the host read-back of kept ranks.  Limit of the Jacobi route: embedded problems must fit the kernel (2 * rows <= 1024)."""
import torch

from . import ops
from ._lib import SynError

F64 = torch.float64
C128 = torch.complex128
JACOBI_MAX_N = 1024
PURIFY_MIN_N = 128             # embedded bond problems at least this large try the spectral-projection solver first (0 = off)
PURIFY_SP2_MAX = 90
PURIFY_NS_MAX = 60
PURIFY_MAX_LIFT = 14           # see syngular/tensor/_sweeps.py: cuts deeper in the spectrum go to the Jacobi route (when it fits)
PURIFY_MAX_LIFT_RANK_GAP = 24
IDENTITY_MIN_M = 128           # complex bonds of at least this many rows that keep their whole space skip the eigen-solve (0 = never)
STRUCTURED_PROJECTION = True   # use syn_dominant_subspace_c128 (even rows of the embedded products only) when the sizes fit


class Cx:
    """Planar complex tensor: `re` and `im` are float64 tensors with identical shape and strides (views allowed)."""
    __slots__ = ("re", "im")

    def __init__(self, re, im):
        assert re.shape == im.shape, (re.shape, im.shape)
        self.re, self.im = re, im

    # ---- construction / conversion ----------------------------------------------------------------------
    @staticmethod
    def from_torch(t):
        """torch tensor (complex128 or float64, already on the device) -> planar."""
        if t.dtype == C128:
            v = torch.view_as_real(t)
            return Cx(v[..., 0].contiguous(), v[..., 1].contiguous())
        t = t.to(F64).contiguous()
        return Cx(t, torch.zeros_like(t))

    def to_torch(self):
        return torch.complex(self.re.contiguous(), self.im.contiguous())

    @staticmethod
    def empty(shape, device):
        return Cx(torch.empty(tuple(shape), dtype=F64, device=device), torch.empty(tuple(shape), dtype=F64, device=device))

    @staticmethod
    def ones(shape, device):
        return Cx(torch.ones(tuple(shape), dtype=F64, device=device), torch.zeros(tuple(shape), dtype=F64, device=device))

    # ---- tensor-like view interface (what the sweeps use) -------------------------------------------------
    @property
    def shape(self):
        return self.re.shape

    @property
    def device(self):
        return self.re.device

    def dim(self):
        return self.re.dim()

    def stride(self, *a):
        return self.re.stride(*a)

    def is_contiguous(self):
        return self.re.is_contiguous() and self.im.is_contiguous()

    def _map(self, f):
        return Cx(f(self.re), f(self.im))

    def reshape(self, *shape):
        return self._map(lambda x: x.reshape(*shape))

    def t(self):
        return self._map(lambda x: x.t())

    def permute(self, *dims):
        return self._map(lambda x: x.permute(*dims))

    def contiguous(self):
        return self._map(lambda x: x.contiguous())

    def unsqueeze(self, d):
        return self._map(lambda x: x.unsqueeze(d))

    def clone(self):
        return self._map(lambda x: x.clone())

    def __getitem__(self, key):
        return Cx(self.re[key], self.im[key])

    def conj(self):
        return Cx(self.re, -self.im)

    def h(self):
        """Conjugate transpose of a 2-D view (the imaginary part is negated into a new buffer)."""
        return Cx(self.re.t(), (-self.im).t())


def is_cx(x):
    return isinstance(x, Cx)


# ---------------------------------------------------------------------------------------------------------
# complex library calls on planar operands
# ---------------------------------------------------------------------------------------------------------
def empty(shape, device):
    return Cx.empty(shape, device)


def gemm(A, B, C, alpha=1.0, beta=0.0, **kw):
    """Complex C = alpha A B + beta C (alpha, beta real) with the real kernel's strided descriptor `kw`: 4 launches."""
    ops.gemm(A.re, B.re, C.re, alpha=alpha, beta=beta, **kw)
    ops.gemm(A.im, B.im, C.re, alpha=-alpha, beta=1.0, **kw)
    ops.gemm(A.re, B.im, C.im, alpha=alpha, beta=beta, **kw)
    ops.gemm(A.im, B.re, C.im, alpha=alpha, beta=1.0, **kw)
    return C


def matmul(a, b, out=None, alpha=1.0, beta=0.0):
    """Complex out = alpha a @ b + beta out on 2-D strided planar views."""
    M, K = a.shape
    K2, N = b.shape
    assert K == K2, (a.shape, b.shape)
    if out is None:
        out = Cx.empty((M, N), a.device)
    return gemm(a, b, out, alpha=alpha, beta=beta, M=M, N=N, K=K, a_m=a.stride(0), a_k=a.stride(1), b_k=b.stride(0), b_n=b.stride(1),
                c_m=out.stride(0), c_n=out.stride(1))


def copy_strided(src):
    return Cx(ops.copy_strided(src.re), ops.copy_strided(src.im))


def add_site(A, B, first, last):
    """Block assembly of `A + B` is linear: one kernel call per part."""
    return Cx(ops.add_site(A.re, B.re, first, last), ops.add_site(A.im, B.im, first, last))


def embed(A):
    """(m x n) complex -> (2m x 2n) real with every entry a + ib replaced by [[a, -b], [b, a]] (rows and columns interleaved)."""
    m, n = A.shape
    E = torch.empty((m, 2, n, 2), dtype=F64, device=A.device)
    E[:, 0, :, 0] = A.re
    E[:, 1, :, 1] = A.re
    E[:, 1, :, 0] = A.im
    E[:, 0, :, 1] = -A.im
    return E.reshape(2 * m, 2 * n)


def unembed_columns(Qe, m, k):
    """Even columns of an embedded (2m x 2k) basis -> planar complex (m x k): column j = Qe[0::2, 2j] + i Qe[1::2, 2j]."""
    V = Qe.reshape(m, 2, k, 2)
    return Cx(V[:, 0, :, 0].contiguous(), V[:, 1, :, 0].contiguous())


def gram_deviation(Q):
    """max |Q^H Q - I| of a planar (m x k) matrix: device scalar (two Gram GEMM pairs + the identity-deviation kernel)."""
    k = Q.shape[1]
    G = matmul(Q.h(), Q)
    d_re = ops.identity_deviation(G.re)
    # imaginary part must vanish: measure it as the deviation of (I + G_im) from the identity
    eye = torch.eye(k, dtype=F64, device=Q.device)
    d_im = ops.identity_deviation(G.im + eye)
    return torch.maximum(d_re, d_im)


ORTHO_TOL = 1e-13


def polish_columns(Q, dev=None, max_steps=60):
    """Newton-Schulz orthonormalisation of the columns of a planar (m x k) matrix, Q <- Q (1.5 I - 0.5 Q^H Q), until
    max|Q^H Q - I| < ORTHO_TOL.  Quadratic once the deviation is below ~0.5 (2 steps from 1e-4); preserves the column space."""
    d = float((gram_deviation(Q) if dev is None else dev).item())
    steps = 0
    while d > ORTHO_TOL:
        if steps >= max_steps:
            raise SynError("complex basis did not orthonormalise (deviation %.2e after %d Newton-Schulz steps)" % (d, steps))
        G = matmul(Q.h(), Q)
        Qn = copy_strided(Q)
        matmul(Q, G, out=Qn, alpha=-0.5, beta=1.5)
        Q = Qn
        d = float(gram_deviation(Q).item())
        steps += 1
    return Q


DEPENDENT_TOL = 1e-12


def _embedded_basis(Aq, qk, want_R):
    """Real Householder factorisation of the embedding of Aq (m x k): (planar basis (m x qk) from the even columns, |R_jj| of the
    complex columns as a device vector or None)."""
    m, k = Aq.shape
    Qe, Se = ops.qrt(embed(Aq), 2 * qk, want_S=want_R)
    Q = unembed_columns(Qe, m, qk)
    if not want_R:
        return Q, None
    kk = min(qk, k)
    return Q, Se[: 2 * kk, : 2 * kk].diagonal()[0::2].abs()


def qrt(A, q, want_S=True):
    """Complex truncation step (reference MPS:443-446 on complex cores): Q (m x qk) = orthonormal basis of span(A[:, :q]),
    S = Q^H A, qk = min(q, m).  Runs the real Householder kernel on the embedded first q columns: with rows and columns
    interleaved the real factorisation is the embedding of the complex one, so its even columns are the complex basis.
    When the leading columns are numerically dependent the reflectors built from rounding noise break that pairing (detected
    by the orthonormality check): dependent columns are then removed one at a time (the first flagged column is always
    reliable), the basis is completed with fixed pseudo-random columns and factored again.  q > n (the reference's bond inflation by
    the complete QR) is completed the same way.  The completion directions are as arbitrary as LAPACK's in the reference;
    span(A[:, :q]) is reproduced exactly."""
    m, n = A.shape
    q = int(q)
    qk, qc = min(q, m), min(q, n)
    Aq = A[:, :qc]
    if not Aq.is_contiguous():
        Aq = copy_strided(Aq)

    def completed(cols):
        """[A[:, cols] | fixed pseudo-random columns] with qk columns in all: q > n inflates the bond like the reference's complete QR
        (the extra directions are arbitrary there too), and dropped dependent columns are replaced the same way."""
        extra = qk - len(cols)
        if extra <= 0 and len(cols) == qc:
            return Aq
        g = torch.Generator(device="cpu").manual_seed(1000 * m + q)
        fill = torch.randn((2, m, max(extra, 0)), dtype=F64, generator=g).to(A.device)
        return Cx(torch.cat([Aq.re[:, cols], fill[0]], dim=1).contiguous(), torch.cat([Aq.im[:, cols], fill[1]], dim=1).contiguous())

    Q, _ = _embedded_basis(completed(list(range(min(qc, qk)))), qk, want_R=False)
    if float(gram_deviation(Q).item()) > ORTHO_TOL:
        cols = list(range(qc))
        while cols:
            sub = Cx(Aq.re[:, cols].contiguous(), Aq.im[:, cols].contiguous())
            _, diag = _embedded_basis(sub, min(len(cols), m), want_R=True)
            d = diag.cpu()
            bad = (d <= DEPENDENT_TOL * float(d.max())).nonzero()
            if bad.numel() == 0 and len(cols) <= m:
                break
            cols.pop(int(bad[0]) if bad.numel() else len(cols) - 1)
        Q, _ = _embedded_basis(completed(cols), qk, want_R=False)
        Q = polish_columns(Q)
    S = matmul(Q.h(), A) if want_S else None
    return Q, S


POLAR_MIN_ROWS = 128           # tall complex unfoldings with at least this many rows are reduced by the Newton-Schulz kernel (0 = never)


def polar_basis(A):
    """Orthonormal basis Q (m x c planar) of the column space of a tall planar complex matrix A (m > c) as the polar factor
    A (A^H A)^(-1/2), by the fused Newton-Schulz kernel on the interleaved real embedding (a polynomial in E^T E applied to E is an
    embedding again, so the even columns are the complex basis).  A 2048 x 1024 embedding takes ~4 ms against 51 ms for the Householder
    kernel.  Returns None when the sizes do not fit or the iteration did not reach orthonormality (rank-deficient A): the caller then
    takes qrt."""
    m, c = A.shape
    if not (POLAR_MIN_ROWS and m >= POLAR_MIN_ROWS and m > c and ops.orthonormalize_columns_fits(2 * m, 2 * c)):
        return None
    Qe, info = ops.orthonormalize_columns(embed(A))
    h = info.cpu()
    if not (bool(torch.isfinite(h).all()) and float(h[4]) < 1e-12):
        return None
    return polish_columns(unembed_columns(Qe, m, c))


def _hermitian_embedding(H):
    """(m x m) Hermitian planar -> (2m x 2m) real symmetric [[Hr, -Hi], [Hi, Hr]] (block layout; symmetrised)."""
    top = torch.cat([H.re, -H.im], dim=1)
    bot = torch.cat([H.im, H.re], dim=1)
    S = torch.cat([top, bot], dim=0)
    return (0.5 * (S + S.t())).contiguous()


def svd_basis(M, chi_max, cutoff, eigh, rank_tol=3.2e-7):
    """Left singular basis of a planar complex unfolding M (m x c), m <= c or reduced by the caller:
    returns (U (m x keep) planar with orthonormal columns, keep, sigma (device, m values, descending), discarded weight).
    `eigh(S, chi, cutoff, rank_tol)` is the real symmetric eigen-solver (syngular.tensor._sweeps.eigh_gram)."""
    m, c = M.shape
    target = min(int(chi_max), m, c)
    if IDENTITY_MIN_M and cutoff == 0.0 and target == m and m >= IDENTITY_MIN_M:
        # nothing is truncated at this bond: every unitary basis of the whole row space is a valid gauge -- the split U (U^H M) is exact for
        # any unitary U, and later truncations see the same state (their local SVDs differ by that unitary on the bond index only) -- so
        # the identity is taken and no eigen-problem is solved (the growth phase of a circuit: one-sided Jacobi on the 2m x 2m embedding
        # took 21.7 ms at m = 512).  Same rule as _sweeps.IDENTITY_WHEN_FULL.  Small bonds keep the rank-revealing SVD (structured
        # circuits -- CX ladders, GHZ -- stay at their true ranks); a large bond that is rank-deficient here is pruned by the next
        # truncating split, whose solvers drop singular values below rank_tol.
        eye = torch.eye(m, dtype=F64, device=M.device)
        return Cx(eye, torch.zeros_like(eye)), m, None, torch.zeros((), dtype=F64, device=M.device)
    H = matmul(M, M.h())                                         # Hermitian Gram matrix (squared singular values)
    if PURIFY_MIN_N and 2 * m >= PURIFY_MIN_N and cutoff == 0.0 and target < m and ops.dominant_subspace_fused_fits(2 * m, 2 * target):
        # Spectral projection on the INTERLEAVED embedding: the projector is a polynomial of the embedded matrix, hence itself an
        # embedding, and its first 2*target columns are the pairs (P e_j, J P e_j) -- Newton-Schulz keeps that structure, so the even
        # columns of the orthonormalised basis are the complex basis (no pairing problem, no size limit from the Jacobi kernel).
        if STRUCTURED_PROJECTION and ops.dominant_subspace_c128_fits(m, target):
            # planar in / planar out; the kernel computes only the even rows of every embedded product (half the DMMA work)
            Ure, Uim, info = ops.dominant_subspace_c128(H.re.contiguous(), H.im.contiguous(), target, sp2_max=PURIFY_SP2_MAX, ns_max=PURIFY_NS_MAX)
            V = None
        else:
            V, info = ops.dominant_subspace(embed(H), 2 * target, sp2_max=PURIFY_SP2_MAX, ns_max=PURIFY_NS_MAX)
        h = info.cpu()
        tr, f2, kept_w, dev, tr_a, idem = (float(h[k]) for k in (0, 1, 2, 4, 5, 6))
        if (abs(tr - 2 * target) < 2e-9 * target and abs(f2 - 2 * target) < 2e-9 * target and abs(idem) < 2e-11 * target and dev < 1e-12
                and bool(torch.isfinite(h).all())
                and (int(h[7]) // 1000000 <= (PURIFY_MAX_LIFT_RANK_GAP if target >= min(m, c) else PURIFY_MAX_LIFT) or 2 * m > JACOBI_MAX_N)):
            U = polish_columns(Cx(Ure, Uim) if V is None else unembed_columns(V, m, target))
            return U, target, None, torch.tensor(max(0.5 * (tr_a - kept_w), 0.0), dtype=F64)
    if 2 * m > JACOBI_MAX_N:
        raise NotImplementedError("complex bond SVD without a spectral gap at the cut needs 2 * rows <= %d (got rows = %d)" % (JACOBI_MAX_N, m))
    S = _hermitian_embedding(H)
    Ut, sigma2, info, winfo = eigh(S, 2 * int(chi_max), cutoff, rank_tol)
    keep = max(1, int(info[0].item()) // 2)
    sigma = sigma2[0::2]
    discarded = 0.5 * winfo[0]
    # rows of Ut are the embedded eigenvectors [x ; y], pairs (v, Jv) adjacent when the spectrum is simple: keep one per pair
    V = Ut[0:2 * keep:2]
    U = Cx(ops.copy_strided(V[:, :m].t()), ops.copy_strided(V[:, m:].t()))          # (m, keep): column j = x_j + i y_j
    dev = gram_deviation(U)
    d = float(dev.item())
    if d > 1e-3:
        # degenerate singular values: the Jacobi basis of a cluster is an arbitrary real rotation of the (v, Jv) pairs, so the
        # even rows need not be complex-independent.  Mix ALL 2 keep rows with a fixed well-conditioned real matrix (the span is
        # the same complex subspace) and orthonormalise.
        W = Ut[0:2 * keep]
        C = Cx(ops.copy_strided(W[:, :m].t()), ops.copy_strided(W[:, m:].t()))      # (m, 2 keep)
        g = torch.Generator(device="cpu").manual_seed(keep)
        mix = torch.randn((2 * keep, keep), dtype=F64, generator=g).to(M.device) / (2.0 * float(2 * keep) ** 0.5)
        U = matmul(C, Cx(mix, torch.zeros_like(mix)))
        U = polish_columns(U)
    else:
        U = polish_columns(U, dev=dev)
    return U, keep, sigma, discarded
