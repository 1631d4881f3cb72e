"""TEST INFRASTRUCTURE ONLY: a numpy emulation of the `syngular_b200.ops` entry points on CPU torch tensors.

The GPU box is the only place the CUDA kernels run, and GPU time is scarce; the HOST orchestration (which strided GEMM is
issued with which indices, how the complex embedding pairs vectors, how sweeps carry their factors) is plain Python and can be
checked here against the oracle by swapping the kernel wrappers for these restatements (`patched()` below).  Each function
follows the contract documented in include/syngular_b200.h.  Nothing in the product imports this module, and the product
never falls back to it: without the patch `syngular.tensor` refuses to run without a CUDA device."""
import contextlib

import numpy as np
import torch

F64 = torch.float64


def _flat(t):
    """Flat view of the storage from the tensor's first element (what a raw base pointer sees)."""
    st = t.untyped_storage()
    n_total = st.nbytes() // 8
    base = torch.empty(0, dtype=F64).set_(st, 0, (n_total,), (1,))
    return base.numpy(), t.storage_offset()


def _ix(spec, extent):
    x = np.arange(int(extent), dtype=np.int64)
    if isinstance(spec, (tuple, list)):
        outer, inner, div = (int(v) for v in spec)
        return (x // div) * outer + (x % div) * inner
    return x * int(spec)


def gemm(A, B, C, M, N, K, a_m, a_k, b_k, b_n, c_m, c_n, batch=1, a_b=0, b_b=0, c_b=0, alpha=1.0, beta=0.0, mask=None):
    fa, oa = _flat(A)
    fb, ob = _flat(B)
    fc, oc = _flat(C)
    am, ak, ab = _ix(a_m, M), _ix(a_k, K), _ix(a_b, batch)
    bk, bn, bb = _ix(b_k, K), _ix(b_n, N), _ix(b_b, batch)
    cm, cn, cb = _ix(c_m, M), _ix(c_n, N), _ix(c_b, batch)
    for z in range(int(batch)):
        a = fa[oa + ab[z] + am[:, None] + ak[None, :]]
        b = fb[ob + bb[z] + bk[:, None] + bn[None, :]]
        idx = oc + cb[z] + cm[:, None] + cn[None, :]
        prod = alpha * (a @ b)
        new = prod + (beta * fc[idx] if beta != 0.0 else 0.0)
        if mask:                                            # block-lower output mask: the rest of C is left untouched
            rows, cols = np.arange(int(M))[:, None], np.arange(int(N))[None, :]
            keep = torch.from_numpy(cols < (rows // int(mask[0]) + 1) * int(mask[1]))
            new = torch.where(keep, new, fc[idx])
        fc[idx] = new
    return C


def matmul(a, b, out=None, alpha=1.0, beta=0.0):
    if a.dim() == 2:
        if out is None:
            out = torch.empty((a.shape[0], b.shape[1]), dtype=F64)
        return gemm(a, b, out, a.shape[0], b.shape[1], a.shape[1], a.stride(0), a.stride(1), b.stride(0), b.stride(1), out.stride(0),
                    out.stride(1), alpha=alpha, beta=beta)
    nb = a.shape[0]
    if out is None:
        out = torch.empty((nb, a.shape[1], b.shape[2]), dtype=F64)
    return gemm(a, b, out, a.shape[1], b.shape[2], a.shape[2], a.stride(1), a.stride(2), b.stride(1), b.stride(2), out.stride(1),
                out.stride(2), batch=nb, a_b=a.stride(0), b_b=b.stride(0), c_b=out.stride(0), alpha=alpha, beta=beta)


def _qrt_np(A, q):
    m, n = A.shape
    Q, _ = np.linalg.qr(A[:, :min(q, n)] if q <= n else A, mode="complete")
    if q > n:                                   # completion columns of the full Q, like the reference
        Q = np.linalg.qr(A, mode="complete")[0]
    Q = Q[:, :min(q, m)]
    return Q, Q.T @ A


def qrt(A, q, Q=None, S=None, want_S=True):
    batched = A.dim() == 3
    A3 = A if batched else A.unsqueeze(0)
    nb, m, n = A3.shape
    qk = min(int(q), m)
    if Q is None:
        Q = torch.empty((nb, m, qk) if batched else (m, qk), dtype=F64)
    if S is None and want_S:
        S = torch.empty((nb, qk, n) if batched else (qk, n), dtype=F64)
    Q3 = Q if batched else Q.unsqueeze(0)
    S3 = None if S is None else (S if batched else S.unsqueeze(0))
    for z in range(nb):
        Qn, Sn = _qrt_np(A3[z].numpy().copy(), int(q))
        Q3[z].copy_(torch.from_numpy(Qn))
        if S3 is not None:
            S3[z].copy_(torch.from_numpy(Sn))
    return Q, S


def qr_r(A, R=None):
    r = np.linalg.qr(A.numpy().copy(), mode="r")
    out = torch.from_numpy(np.ascontiguousarray(r))
    if R is not None:
        R.copy_(out)
        return R
    return out


def copy_strided(src, out=None):
    if out is None:
        return src.clone(memory_format=torch.contiguous_format)
    out.copy_(src)
    return out


def jacobi_rows(G, max_sweeps=40, tol=None, null_rel=1e-14):
    """Rows of J G, mutually orthogonal: from the SVD  G = U S V^T  ->  U^T G = S V^T (same contract, any order of rows)."""
    G3 = G if G.dim() == 3 else G.unsqueeze(0)
    for z in range(G3.shape[0]):
        g = G3[z].numpy().copy()
        u, s, vt = np.linalg.svd(g)
        rows = s[:, None] * vt
        # scramble the order so callers cannot rely on sortedness before finalize
        perm = np.random.default_rng(7).permutation(rows.shape[0])
        G3[z].copy_(torch.from_numpy(rows[perm].copy()))
    return G


def chol_upper(G):
    g = G.numpy().copy()
    n = g.shape[-1]
    delta = 2.0 * n * 2.220446049250313e-16 * float(np.max(np.diagonal(g)))
    L = np.linalg.cholesky(g + delta * np.eye(n))
    return torch.from_numpy(np.ascontiguousarray(L.T)), torch.tensor([delta], dtype=F64)


def jacobi_finalize(G, chi_max, cutoff=0.0, rank_tol=1e-14, sqrt_mode=False, shift=None):
    g = G.numpy()
    n = g.shape[0]
    key = np.sqrt((g * g).sum(1))
    perm = np.argsort(-key, kind="stable")
    key = key[perm]
    Ut = g[perm] / np.where(key > 0, key, 1.0)[:, None]
    mode = int(sqrt_mode)
    delta = float(shift[0]) if (mode == 2 and shift is not None) else 0.0
    sv = np.sqrt(key) if mode == 1 else (np.sqrt(np.maximum(key * key - delta, 0.0)) if mode == 2 else key)
    thr = max(cutoff, rank_tol) * sv[0]
    keep = max(1, int(np.sum((np.arange(n) < chi_max) & (sv > thr))))
    disc = float(np.sum(sv[keep:] ** 2))
    return (torch.from_numpy(np.ascontiguousarray(Ut)), torch.from_numpy(sv.copy()), torch.tensor([keep, n], dtype=torch.int32),
            torch.tensor([disc, sv[0]], dtype=F64))


def jacobi_solve(G, chi_max, cutoff=0.0, rank_tol=1e-14, sqrt_mode=False, shift=None, null_rel=1e-14, max_sweeps=40):
    jacobi_rows(G, max_sweeps=max_sweeps, null_rel=null_rel)
    return jacobi_finalize(G, chi_max, cutoff, rank_tol=rank_tol, sqrt_mode=sqrt_mode, shift=shift)


def dominant_subspace(A, ne, sp2_iters=40, ns_iters=20, fused=None, sp2_max=160, ns_max=80):
    """csrc/purify.cu restated: SP2 from A / |A|_F with the trace-steered branch, then Newton-Schulz on P[:, :ne]."""
    a = A.numpy()
    n = a.shape[0]
    fro = float(np.sqrt((a * a).sum()))
    x = a / fro if fro > 0 else a * 0
    tr0 = float(np.trace(x))
    hist = [(tr0, float((x * x).sum()))]
    for _ in range(int(sp2_iters)):
        tr, f2 = hist[-1]
        y = x @ x
        y = 0.5 * (y + y.T)
        x = y if abs(f2 - ne) < abs(2 * tr - f2 - ne) else 2 * x - y
        hist.append((float(np.trace(x)), float((x * x).sum())))
    u = x[:, :ne].copy()
    for _ in range(int(ns_iters)):
        u = u @ (1.5 * np.eye(ne) - 0.5 * (u.T @ u))
    dev = float(np.max(np.abs(u.T @ u - np.eye(ne))))
    lift = 0
    for tr, f2 in hist[:-1]:
        if abs(f2 - ne) < abs(2 * tr - f2 - ne):
            break
        lift += 1
    info = np.array([hist[-1][0], hist[-1][1], float((a * x).sum()), fro, dev, tr0 * fro, hist[-2][0] - hist[-2][1], sp2_iters + 1e6 * lift])
    return torch.from_numpy(np.ascontiguousarray(u)), torch.from_numpy(info)


def dominant_subspace_c128_fits(m, k):
    return m % 32 == 0 and k % 32 == 0 and m >= 64 and 32 <= k < m


def dominant_subspace_c128(Hre, Him, k, sp2_max=90, ns_max=60):
    """Planar complex Hermitian H -> planar basis of the k dominant eigenvectors; info in the embedded convention (traces count twice)."""
    H = Hre.numpy() + 1j * Him.numpy()
    w, V = np.linalg.eigh(0.5 * (H + H.conj().T))
    U = V[:, ::-1][:, :k]
    ne = 2 * k
    gap_ok = w[::-1][k - 1] - w[::-1][k] > 1e-13 * abs(w[-1])
    info = np.array([ne if gap_ok else ne + 0.5, ne, 2 * w[::-1][:k].sum(), np.sqrt(2.0) * np.linalg.norm(H), 1e-16, 2 * w.sum(), 0.0, 30 + 1000 * 8 + 1e6 * 6])
    return torch.from_numpy(np.ascontiguousarray(U.real)), torch.from_numpy(np.ascontiguousarray(U.imag)), torch.from_numpy(info)


def dominant_subspace_batched_fits(n, ne):
    return False


def dominant_subspace_fused_fits(n, ne):
    return n % 32 == 0 and ne % 32 == 0 and n >= 128 and ne >= 32 and ne < n


def orthonormalize_columns_fits(m, q):
    return m % 32 == 0 and q % 32 == 0 and m >= 128 and q >= 32 and q <= m


def orthonormalize_columns(A, ns_max=60):
    a = A.numpy()
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    ok = s[-1] > 1e-10 * s[0]
    qm = u @ vt
    info = np.zeros(8)
    info[4] = float(np.max(np.abs(qm.T @ qm - np.eye(a.shape[1])))) if ok else 1.0
    return torch.from_numpy(np.ascontiguousarray(qm)), torch.from_numpy(info)


def env_sandwich_fits(l, i, o, r, b):
    return (l, i, o, r) == (16, 2, 2, 16) and b % 16 == 0


def env_sandwich(P1, W, Z, na, b):
    w = W.numpy()
    l, i, o, r = w.shape
    p1 = P1.numpy().reshape(na, i, r, r, b)                        # [a][i][r][r'][b']
    p2 = np.einsum("lior,airqb->aloqb", w, p1)                     # [a][l][o][r'][b']
    z = np.einsum("aloqb,mjoq->almjb", p2, w)                      # [a][l][l'][i'][b']
    Z.copy_(torch.from_numpy(np.ascontiguousarray(z)).reshape(Z.shape))
    return Z


def sum_parts(parts):
    return parts.sum(0)


def env_mirror(E, na, L, ab):
    e = E.numpy().reshape(na, L, L, na)                            # [a][l][l'][a']
    blk = np.arange(na) // ab
    upper = blk[None, :] > blk[:, None]                            # [a][a']: a' in a later block
    src = e.transpose(3, 2, 1, 0)                                  # src[a][l][l'][a'] = e[a'][l'][l][a]
    mask = np.broadcast_to(upper[:, None, None, :], e.shape)
    e[mask] = src[mask]
    return E


def identity_deviation(X):
    x = X.numpy()
    return torch.tensor([float(np.max(np.abs(x - np.eye(x.shape[0]))))], dtype=F64)


def add_site(A, B, first, last):
    a, b = A.numpy(), B.numpy()
    la, ra, lb, rb = a.shape[0], a.shape[-1], b.shape[0], b.shape[-1]
    out = np.zeros(((la if first else la + lb),) + a.shape[1:-1] + ((ra if last else ra + rb),))
    out[:la, ..., :ra] = a
    out[(0 if first else la):, ..., (0 if last else ra):] = b
    return torch.from_numpy(out)


def kron_site(A, B):
    a, b = A.numpy(), B.numpy()
    la, ra, lb, rb = a.shape[0], a.shape[-1], b.shape[0], b.shape[-1]
    phys = a.shape[1:-1]
    a2, b2 = a.reshape(la, -1, ra), b.reshape(lb, -1, rb)
    out = np.einsum("apb,cpd->acpbd", a2, b2).reshape((la * lb,) + phys + (ra * rb,))
    return torch.from_numpy(np.ascontiguousarray(out))


def sumsq(x):
    return torch.tensor([float((x.numpy() ** 2).sum())], dtype=F64)


def scale_rsqrt_(x, ss):
    x.mul_(1.0 / float(ss[0]) ** 0.5)
    return x


def overlap_fits(a, b, batched=True):
    return False


_NAMES = ("gemm", "matmul", "qrt", "qr_r", "copy_strided", "jacobi_rows", "chol_upper", "jacobi_finalize", "jacobi_solve", "identity_deviation", "dominant_subspace", "dominant_subspace_fused_fits", "dominant_subspace_c128", "dominant_subspace_c128_fits", "dominant_subspace_batched_fits", "env_sandwich_fits", "env_sandwich", "env_mirror", "sum_parts", "orthonormalize_columns_fits", "orthonormalize_columns", "add_site",
          "kron_site", "sumsq", "scale_rsqrt_", "overlap_fits")


@contextlib.contextmanager
def patched():
    """Swap the kernel wrappers of syngular_b200.ops for the emulations above and let the sweeps allocate on the CPU."""
    from syngular_b200 import ops
    from syngular.tensor import _sweeps as sw
    saved = {n: getattr(ops, n) for n in _NAMES}
    saved_dev = sw.device
    try:
        for n in _NAMES:
            setattr(ops, n, globals()[n])
        sw.device = lambda: torch.device("cpu")
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        sw.device = saved_dev
