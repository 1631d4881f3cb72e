"""Batches of independent matrix-product states on one GPU (BASELINE configs[3]: 8192 random MPS, N=32, d=2, chi=64).

The reference has no batching construct on this path: a batch is a Python loop over `A | B` (MPS:116-129) or `W @ X`
(MPO:181-192).  Here core k of all B states is ONE (B, l, d, r) CUDA tensor and every step of the transfer-matrix / apply+round
sweeps is a single batched launch of the same kernels the single-chain path uses.  Across GPUs the batch is sharded in
contiguous blocks (syngular_b200.parallel) and only the B result scalars are gathered.
"""
import numpy as np
import torch

from . import ops

F64 = torch.float64
SMALL_CORE = True                              # shared-MPO-core contractions by the streaming kernel csrc/smallcore.cu (False: GEMMs)
PROJECTION_BATCHED = True                      # bond problems that fit csrc/purify_batched.cu use it (False: Jacobi only)
PROJECTION_STATS = {"taken": 0, "fallback": 0}
RANGE_FINDER = True                            # chain-end bonds (kept rank = dimension of the right part): QR of a Gaussian mix of M E
RANGE_FINDER_STATS = {"bonds": 0}
_GAUSS = {}


def _gauss(D, k, dev):
    """A fixed (D x k) standard normal matrix per shape and device (seeded: results are reproducible)."""
    key = (int(D), int(k), str(dev))
    if key not in _GAUSS:
        g = torch.Generator(device="cpu").manual_seed(1000003 * int(D) + int(k))
        _GAUSS[key] = (torch.randn((int(D), int(k)), dtype=F64, generator=g) / float(D) ** 0.5).to(dev)
    return _GAUSS[key]


def _rejected(h, ne, rank_gap):
    """Indices of the batch members whose projection result is not accepted: the per-member form of _sweeps._projection_verdict on the
    (B, 8) host copy of the info doubles."""
    from syngular.tensor import _sweeps as sw
    tr, f2, dev, idem = h[:, 0], h[:, 1], h[:, 4], h[:, 6]
    lift = (h[:, 7] // 1000000).astype(np.int64)
    with np.errstate(invalid="ignore"):
        ok = ((np.abs(tr - ne) < 1e-9 * ne) & (np.abs(f2 - ne) < 1e-9 * ne) & (np.abs(idem) < 1e-11 * ne) & (dev < 1e-12)
              & np.isfinite(h).all(axis=1) & (lift <= (sw.PURIFY_MAX_LIFT_RANK_GAP if rank_gap else sw.PURIFY_MAX_LIFT)))
    return np.nonzero(~ok)[0]


class BatchedMatrixProductState:
    def __init__(self, cores):
        self.sites = [c.contiguous() for c in cores]
        self.batch = int(self.sites[0].shape[0])
        self.sites_number = len(self.sites)
        self.shape = [tuple(int(x) for x in c.shape[1:]) for c in self.sites]
        self.bond_shape = tuple(s[-1] for s in self.shape[:-1])
        self.input_shape = tuple(s[1] for s in self.shape)

    # ---- constructors ---------------------------------------------------------------------------------------
    @staticmethod
    def from_states(states):
        """Stack single MatrixProductState objects (or lists of cores) of identical shapes."""
        lists = [s.sites if hasattr(s, "sites") else s for s in states]
        n = len(lists[0])
        cores = []
        for k in range(n):
            cores.append(torch.stack([torch.as_tensor(l[k]).to(F64) if not isinstance(l[k], torch.Tensor) else l[k] for l in lists]).cuda())
        return BatchedMatrixProductState(cores)

    @staticmethod
    def random(batch, input_shape, bond_shape, seed=0, device=None):
        """Cores ~ N(0, 1/(l d)) drawn on the device (large batches: 8192 x 1.35 MB); use from_states for bit-identical host data."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=device).manual_seed(int(seed))
        b = (1,) + tuple(int(x) for x in bond_shape) + (1,)
        cores = []
        for k, d in enumerate(input_shape):
            c = torch.randn((batch, b[k], d, b[k + 1]), dtype=F64, device=device, generator=g)
            cores.append(c / float(np.sqrt(b[k] * d)))
        return BatchedMatrixProductState(cores)

    def state(self, b):
        from syngular.tensor import MatrixProductState
        return MatrixProductState.from_sites([c[b] for c in self.sites])

    # ---- `|` for the whole batch ----------------------------------------------------------------------------------
    def overlap(self, other):
        """(B,) tensor of <self_b | other_b> (bilinear, like MPS:116-129): one launch of the fused transfer-matrix kernel
        (csrc/overlap.cu, one CTA per pair of chains) when every bond is <= 64, else two batched strided GEMMs per site."""
        assert other.batch == self.batch and other.sites_number == self.sites_number
        B = self.batch
        if self.sites[0].shape[1] == 1 and other.sites[0].shape[1] == 1 and ops.overlap_fits(self.sites, other.sites):
            return ops.overlap_batched(self.sites, other.sites).reshape(B)
        E = torch.ones((B, 1, 1), dtype=F64, device=self.sites[0].device)
        for a, b in zip(self.sites, other.sites):
            la, ra, lb, rb = a.shape[1], a.shape[-1], b.shape[1], b.shape[-1]
            T = ops.matmul(E.transpose(1, 2), a.reshape(B, la, -1))                  # (B, lb, d*ra)
            E = ops.matmul(T.reshape(B, -1, ra).transpose(1, 2), b.reshape(B, -1, rb))  # (B, ra, rb)
        return E.reshape(B)

    def norms2(self):
        return self.overlap(self)

    # ---- MPO application + rounding for the whole batch (shared operator) ---------------------------------------------
    def apply_round(self, mpo, dim):
        """`W @ X_b` followed by `>> dim` (reference QR-truncation semantics) for every state of the batch with ONE shared MPO;
        fused like the single-chain path (product cores never formed), every step batched over B."""
        W = mpo.sites if hasattr(mpo, "sites") else mpo
        B, n = self.batch, self.sites_number
        dev = self.sites[0].device
        T = torch.ones((B, 1, 1, 1), dtype=F64, device=dev)            # carry (B, s, l, a)
        out = []
        for k in range(n):
            X, Wk = self.sites[k], W[k]
            _, a, i, b = X.shape
            l, _, o, r = Wk.shape
            s = T.shape[1]
            T1 = torch.empty((B, s, l, i, b), dtype=F64, device=dev)
            ops.gemm(T, X, T1, M=s, N=i * b, K=a, a_m=l * a, a_k=1, b_k=i * b, b_n=1, c_m=l * i * b, c_n=1,
                     batch=B * l, a_b=(s * l * a, a, l), b_b=(a * i * b, 0, l), c_b=(s * l * i * b, i * b, l))
            M = torch.empty((B, s, o, b * r), dtype=F64, device=dev)
            if SMALL_CORE and ops.small_core_fits(l * i, o * r):         # shared small core: streaming kernel (see _apply_round_svd_chunk)
                ops.apply_small_core(T1, Wk.reshape(l * i, o * r).t().contiguous(), M, Q=B * s, L=b,
                                     x_q=l * i * b, x_r=b, x_l=1, y_q=o * b * r, y_ro=(b * r, 1, r), y_l=r)
            else:
                ops.gemm(T1, Wk, M, M=s * b, N=o * r, K=l * i, a_m=(l * i * b, 1, b), a_k=b, b_k=o * r, b_n=1,
                         c_m=(o * b * r, r, b), c_n=(b * r, 1, r), batch=B, a_b=s * l * i * b, b_b=0, c_b=s * o * b * r)
            if k == n - 1:
                out.append(M)
                break
            L = M.reshape(B, s * o, b * r)
            Q, _ = ops.qrt(L, dim, want_S=False)
            kept = Q.shape[2]
            out.append(Q.reshape(B, s, o, kept))
            Tn = torch.empty((B, kept, r, b), dtype=F64, device=dev)
            ops.gemm(Q, L, Tn, M=kept, N=b * r, K=s * o, a_m=1, a_k=kept, b_k=b * r, b_n=1, c_m=r * b, c_n=(1, b, r),
                     batch=B, a_b=s * o * kept, b_b=s * o * b * r, c_b=kept * r * b)
            T = Tn
        return BatchedMatrixProductState(out)

    # ---- MPO application + optimal (SVD) rounding for the whole batch -------------------------------------------------
    def apply_round_svd(self, mpo, chi, chunk=512):
        """`W @ X_b` + SVD rounding to bond `chi` for every state with ONE shared MPO: the density-matrix algorithm of the
        single-chain path (syngular.tensor._sweeps.apply_round_dm), every GEMM / Jacobi launch batched over the states.
        Fixed-rank variant: every member keeps exactly min(chi, rows, D, right space) vectors per bond (no per-member cutoff), so the
        results stack.  Environments cost B * D^2 doubles per site, hence the chunking over the batch."""
        W = mpo.sites if hasattr(mpo, "sites") else mpo
        outs = None
        for lo in range(0, self.batch, chunk):
            part = BatchedMatrixProductState([c[lo:lo + chunk] for c in self.sites])._apply_round_svd_chunk(W, chi)
            outs = [[c] for c in part] if outs is None else [o + [c] for o, c in zip(outs, part)]
        return BatchedMatrixProductState([torch.cat(o, dim=0) if len(o) > 1 else o[0] for o in outs])

    @staticmethod
    def _jacobi_basis(A, chi, keep):
        """(B, rows, keep) orthonormal bases of the dominant eigenspaces of the Gram matrices A by one-sided Jacobi on the rows of the
        shifted Cholesky factor (fewer sweeps than on A itself, see csrc/chol.cu)."""
        Bf, shift = ops.chol_upper(A)
        Ut, sigma, info, winfo = ops.jacobi_solve(Bf, chi, rank_tol=0.0, sqrt_mode=2, shift=shift, null_rel=0.0)
        return ops.copy_strided(Ut[:, :keep, :].transpose(1, 2))

    def _apply_round_svd_chunk(self, W, chi):
        B, n = self.batch, self.sites_number
        dev = self.sites[0].device

        def empty(*shape):
            return torch.empty(shape, dtype=F64, device=dev)

        # right environments, rows (b,r), columns (r',b')   (see _sweeps.right_environments)
        E = [None] * (n + 1)
        E[n] = torch.ones((B, 1, 1), dtype=F64, device=dev)
        for k in range(n - 1, 0, -1):
            X, Wk, En = self.sites[k], W[k], E[k + 1]
            _, a, i, b = X.shape
            l, _, o, r = Wk.shape
            D = b * r
            P1 = empty(B, a, i, r, D)
            ops.gemm(X, En, P1, M=a * i, N=r * D, K=b, a_m=b, a_k=1, b_k=r * D, b_n=1, c_m=r * D, c_n=1,
                     batch=B, a_b=a * i * b, b_b=D * D, c_b=a * i * r * D)
            P2 = empty(B, a, l, o, D)
            # the shared MPO core against ALL (state, a) slices: an (l o) x (i r) matrix applied to the middle index of P1 -- pure streaming
            # (csrc/smallcore.cu); other core sizes: one (l o) x (B a D) x (i r) GEMM with the slice index folded into the column index
            if SMALL_CORE and ops.small_core_fits(i * r, l * o):
                ops.apply_small_core(P1, Wk.permute(0, 2, 1, 3).reshape(l * o, i * r).contiguous(), P2, Q=B * a, L=D,
                                     x_q=i * r * D, x_r=D, x_l=1, y_q=l * o * D, y_ro=(0, D, l * o), y_l=1)
            else:
                ops.gemm(Wk, P1, P2, M=l * o, N=B * a * D, K=i * r, a_m=(i * o * r, r, o), a_k=(o * r, 1, r), b_k=D, b_n=(i * r * D, 1, D),
                         c_m=D, c_n=(l * o * D, 1, D))
            Z = empty(B, a * l, l, i, b)
            if SMALL_CORE and ops.small_core_fits(o * r, l * i):
                # Z[q = (B, a, l')][(l, i)][b'] = sum_(o, r) W[(l, i), (o, r)] P2[q][(o, r)][b']: the core is read in place (row-major l i x o r)
                ops.apply_small_core(P2, Wk.contiguous().reshape(l * i, o * r), Z, Q=B * a * l, L=b,
                                     x_q=o * D, x_r=b, x_l=1, y_q=l * i * b, y_ro=(0, b, l * i), y_l=1)
            else:
                ops.gemm(P2, Wk, Z, M=B * a * l * b, N=l * i, K=o * r, a_m=(o * D, 1, b), a_k=(D, b, r), b_k=1, b_n=(i * o * r, o * r, i),
                         c_m=(l * i * b, 1, b), c_n=b)                    # likewise: rows (state, a, l', b')
            Ek = empty(B, a * l, l * a)
            ops.gemm(Z, X, Ek, M=a * l, N=a, K=i * b, a_m=l * i * b, a_k=1, b_k=1, b_n=i * b, c_m=a * l, c_n=1,
                     batch=B * l, a_b=(a * l * l * i * b, i * b, l), b_b=(a * i * b, 0, l), c_b=(a * l * a * l, a, l))
            E[k] = Ek
        T = torch.ones((B, 1, 1, 1), dtype=F64, device=dev)            # carry (B, s, l, a)
        out = []
        for k in range(n):
            X, Wk = self.sites[k], W[k]
            _, a, i, b = X.shape
            l, _, o, r = Wk.shape
            s, D = T.shape[1], b * r
            T1 = empty(B, s, l, i, b)
            ops.gemm(T, X, T1, M=s, N=i * b, K=a, a_m=l * a, a_k=1, b_k=i * b, b_n=1, c_m=l * i * b, c_n=1,
                     batch=B * l, a_b=(s * l * a, a, l), b_b=(a * i * b, 0, l), c_b=(s * l * i * b, i * b, l))
            M = empty(B, s * o, D)
            if SMALL_CORE and ops.small_core_fits(l * i, o * r):
                # M[(B, s)][o][b'][r] = sum_(l, i) W[(l, i), (o, r)] T1[(B, s)][(l, i)][b']: the transposed core (o r x l i) as the small matrix,
                # the output's (o, r) index split around b' by the kernel's two-level index
                ops.apply_small_core(T1, Wk.reshape(l * i, o * r).t().contiguous(), M, Q=B * s, L=b,
                                     x_q=l * i * b, x_r=b, x_l=1, y_q=o * D, y_ro=(b * r, 1, r), y_l=r)
            else:
                ops.gemm(T1, Wk, M, M=s * b, N=o * r, K=l * i, a_m=(l * i * b, 1, b), a_k=b, b_k=o * r, b_n=1,
                         c_m=(o * b * r, r, b), c_n=(b * r, 1, r), batch=B, a_b=s * l * i * b, b_b=0, c_b=s * o * D)
            if k == n - 1:
                out.append(M.reshape(B, s, o, D))
                break
            rows = s * o
            right_dim = 1                                               # dimension of the space to the right of this bond
            for wj in W[k + 1:]:
                right_dim = min(right_dim * int(wj.shape[2]), 1 << 30)
            keep = min(int(chi), rows, D, right_dim)
            if keep == rows:
                # nothing is truncated at this bond: every orthonormal basis of the whole space is a valid gauge -- identity cores,
                # the unfolding itself is carried, no eigen-solve (ramp-up bonds; same rule as _sweeps.apply_round_dm)
                eye = torch.eye(rows, dtype=F64, device=dev)
                out.append(eye.expand(B, rows, rows).contiguous().reshape(B, s, o, rows))
                Tn = empty(B, rows, r, b)
                ops.gemm(eye, M, Tn, M=rows, N=D, K=rows, a_m=rows, a_k=1, b_k=D, b_n=1, c_m=r * b, c_n=(1, b, r),
                         batch=B, a_b=0, b_b=rows * D, c_b=rows * r * b)
                T = Tn
                continue
            ME = empty(B, rows, D)                                     # columns permuted back to (b', r') on the fly
            ops.gemm(M, E[k + 1], ME, M=rows, N=D, K=D, a_m=D, a_k=1, b_k=D, b_n=1, c_m=D, c_n=(1, r, b),
                     batch=B, a_b=rows * D, b_b=D * D, c_b=rows * D)
            if RANGE_FINDER and keep == right_dim and keep < rows and keep < D:
                # the cut is the structural rank of the right part (chain end): the kept space is the whole range of A = M E M^T, which is the
                # column space of M E (rank = keep exactly).  A fixed Gaussian mix of its columns has the same range, and Householder QR of
                # that (rows x keep) block is an exact basis of it -- no eigen-problem and no squared conditioning (these bonds sit 25-29
                # doublings deep in the spectrum of A: the projection solver's accuracy guard rejected every member and Jacobi took over).
                Y = ops.matmul(ME, _gauss(D, keep, dev).expand(B, D, keep))
                U, _ = ops.qrt(Y, keep, want_S=False)
                RANGE_FINDER_STATS["bonds"] += 1
                out.append(U.reshape(B, s, o, keep))
                Tn = empty(B, keep, r, b)
                ops.gemm(U, M, Tn, M=keep, N=D, K=rows, a_m=1, a_k=keep, b_k=D, b_n=1, c_m=r * b, c_n=(1, b, r),
                         batch=B, a_b=rows * keep, b_b=rows * D, c_b=keep * r * b)
                T = Tn
                continue
            A = ops.matmul(ME, M.transpose(1, 2))
            U = None
            if PROJECTION_BATCHED and ops.dominant_subspace_batched_fits(rows, keep):
                # spectral projection, one CTA per state (csrc/purify_batched.cu); the verdict of every member is read once per bond and
                # the members without a usable gap at the cut (or with the cut too deep in the spectrum, see _sweeps.PURIFY_MAX_LIFT) are
                # redone by the Jacobi route
                U, info = ops.dominant_subspace_batched(A, keep)
                bad = _rejected(info.cpu().numpy(), keep, keep >= min(D, right_dim))
                PROJECTION_STATS["taken"] += B - len(bad)
                PROJECTION_STATS["fallback"] += len(bad)
                if len(bad):
                    idx = torch.as_tensor(bad, device=dev)
                    U[idx] = self._jacobi_basis(A[idx].contiguous(), chi, keep)
            else:
                U = self._jacobi_basis(A, chi, keep)
            out.append(U.reshape(B, s, o, keep))
            Tn = empty(B, keep, r, b)
            ops.gemm(U, M, Tn, M=keep, N=D, K=rows, a_m=1, a_k=keep, b_k=D, b_n=1, c_m=r * b, c_n=(1, b, r),
                     batch=B, a_b=rows * keep, b_b=rows * D, c_b=keep * r * b)
            T = Tn
        return out
