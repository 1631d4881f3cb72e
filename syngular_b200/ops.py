"""Thin Python wrappers over the C ABI (one function per exported entry point).  No arithmetic happens here."""
import ctypes
import os

import torch

from . import _lib
from ._lib import GemmDesc, check, ix, lib, ptr, require_cuda_f64, stream_ptr


# When set to a list, every Python-level gemm() call is bracketed by CUDA events on the launching stream and
# (start, end, flops, bytes, (M, N, K, batch)) is appended: bench.py uses it to time the dominant kernel inside a real sweep.
GEMM_PROFILE = None


def gemm(A, B, C, M, N, K, a_m, a_k, b_k, b_n, c_m, c_n, batch=1, a_b=0, b_b=0, c_b=0, alpha=1.0, beta=0.0, mask=None):
    """C[m,n] = alpha * sum_k A[m,k] B[k,n] + beta * C[m,n] with two-level strided indices (see syngular_b200.h).
    A, B, C are CUDA float64 tensors used as base pointers (their own strides are ignored).
    mask = (rows, cols): block-lower output, only n < (m // rows + 1) * cols is computed (the rest of C is left untouched)."""
    require_cuda_f64(A, B, C)
    mr, mc = (int(mask[0]), int(mask[1])) if mask else (0, 0)
    d = GemmDesc(int(M), int(N), int(K), int(batch), ix(a_m), ix(a_k), ix(a_b), ix(b_k), ix(b_n), ix(b_b),
                 ix(c_m), ix(c_n), ix(c_b), float(alpha), float(beta), mr, mc)
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.syn_gemm_f64(ctypes.byref(d), ptr(A), ptr(B), ptr(C), stream_ptr()), "syn_gemm_f64")
    if prof is not None:
        e1.record()
        M_, N_, K_, b_ = int(M), int(N), int(K), int(batch)
        frac = 1.0
        if mr:                     # block-lower mask: the computed fraction of the M x N outputs
            frac = sum(min(N_, (m0 // mr + 1) * mc) for m0 in range(0, M_, mr)) * float(min(mr, M_)) / (float(M_) * N_)
        prof.append((e0, e1, 2.0 * M_ * N_ * K_ * b_ * frac, 8.0 * b_ * (M_ * K_ + K_ * N_ + M_ * N_ * frac), (M_, N_, K_, b_) + ((mr, mc) if mr else ())))
    return C


def matmul(a, b, out=None, alpha=1.0, beta=0.0):
    """out = alpha * a @ b + beta * out for 2-D (or batched 3-D) strided views; no copies are made."""
    require_cuda_f64(a, b)
    if a.dim() == 2 and b.dim() == 2:
        M, K = a.shape
        K2, N = b.shape
        assert K == K2, (a.shape, b.shape)
        if out is None:
            out = torch.empty((M, N), dtype=torch.float64, device=a.device)
        return gemm(a, b, out, M, N, K, a.stride(0), a.stride(1), b.stride(0), b.stride(1), out.stride(0), out.stride(1),
                    alpha=alpha, beta=beta)
    assert a.dim() == 3 and b.dim() == 3 and a.shape[0] == b.shape[0]
    nb, M, K = a.shape
    _, K2, N = b.shape
    assert K == K2
    if out is None:
        out = torch.empty((nb, M, N), dtype=torch.float64, device=a.device)
    return gemm(a, b, out, M, N, K, a.stride(1), a.stride(2), b.stride(1), b.stride(2), out.stride(1), out.stride(2),
                batch=nb, a_b=a.stride(0), b_b=b.stride(0), c_b=out.stride(0), alpha=alpha, beta=beta)


# ---------------------------------------------------------------------------------------------------------
_i64, _i32, _sz, _dbl = ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t, ctypes.c_double
lib.syn_qrt_workspace_f64.restype = ctypes.c_size_t
lib.syn_qrt_workspace_f64.argtypes = [_i32, _i32, _i32, _i32]
lib.syn_jacobi_ctrl_bytes.restype = ctypes.c_size_t
lib.syn_jacobi_ctrl_bytes.argtypes = [_i32, _i32]

_workspaces = {}


def workspace(nbytes, device, tag="ws"):
    """A cached, 256-byte aligned scratch buffer owned by PyTorch (the library never allocates)."""
    key = (tag, device)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        buf = torch.empty((int(nbytes) + 7) // 8 + 32, dtype=torch.float64, device=device)
        _workspaces[key] = buf
    return buf


def qrt(A, q, Q=None, S=None, want_S=True):
    """The reference truncation step on a 2-D (or batched 3-D) strided view A (m x n):
    Q (m x qk) = orthonormal basis of span(A[:, :q]) (completed when q > n), S = Q^T A (qk x n), qk = min(q, m).
    Replaces np.linalg.qr(L, mode="complete") + slicing (MPS:443-446, MPO:555-558) and the reduced QRs (MPS:560,574)."""
    require_cuda_f64(A)
    batched = A.dim() == 3
    A3 = A if batched else A.unsqueeze(0)
    nb, m, n = A3.shape
    qk = min(int(q), m)
    if Q is None:
        Q = torch.empty((nb, m, qk) if batched else (m, qk), dtype=torch.float64, device=A.device)
    if S is None and want_S:
        S = torch.empty((nb, qk, n) if batched else (qk, n), dtype=torch.float64, device=A.device)
    Q3 = Q if batched else Q.unsqueeze(0)
    S3 = None if S is None else (S if batched else S.unsqueeze(0))
    wsb = lib.syn_qrt_workspace_f64(m, n, int(q), nb)
    ws = workspace(wsb, A.device)
    qk_out = ctypes.c_int(0)
    rc = lib.syn_qrt_f64(ptr(A3), _i64(A3.stride(1)), _i64(A3.stride(2)), _i64(A3.stride(0)), _i32(m), _i32(n), _i32(int(q)), _i32(nb),
                         ptr(Q3), _i64(Q3.stride(1)), _i64(Q3.stride(2)), _i64(Q3.stride(0)),
                         ptr(S3) if S3 is not None else None,
                         _i64(S3.stride(1) if S3 is not None else 0), _i64(S3.stride(2) if S3 is not None else 0),
                         _i64(S3.stride(0) if S3 is not None else 0),
                         ptr(ws), _sz(ws.numel() * 8), ctypes.byref(qk_out), stream_ptr())
    check(rc, "syn_qrt_f64")
    return Q, S


def qr_r(A, R=None):
    """Upper-triangular R factor (n x n) of a tall strided view A (m x n, m >= n)."""
    require_cuda_f64(A)
    batched = A.dim() == 3
    A3 = A if batched else A.unsqueeze(0)
    nb, m, n = A3.shape
    if R is None:
        R = torch.empty((nb, n, n) if batched else (n, n), dtype=torch.float64, device=A.device)
    R3 = R if batched else R.unsqueeze(0)
    ws = workspace(lib.syn_qrt_workspace_f64(m, n, n, nb), A.device)
    rc = lib.syn_qr_r_f64(ptr(A3), _i64(A3.stride(1)), _i64(A3.stride(2)), _i64(A3.stride(0)), _i32(m), _i32(n), _i32(nb),
                          ptr(R3), _i64(R3.stride(1)), _i64(R3.stride(2)), _i64(R3.stride(0)), ptr(ws), _sz(ws.numel() * 8), stream_ptr())
    check(rc, "syn_qr_r_f64")
    return R


def copy_strided(src, out=None):
    """out (contiguous) = src (any 2-D / batched 3-D strided view); a tiled transpose when src is a .T view."""
    require_cuda_f64(src)
    batched = src.dim() == 3
    s3 = src if batched else src.unsqueeze(0)
    nb, m, n = s3.shape
    if out is None:
        out = torch.empty(tuple(src.shape), dtype=torch.float64, device=src.device)
    assert out.is_contiguous()
    rc = lib.syn_copy_strided_f64(ptr(s3), _i64(s3.stride(1)), _i64(s3.stride(2)), _i64(s3.stride(0)), ptr(out), _i64(m * n),
                                  _i32(m), _i32(n), _i32(nb), stream_ptr())
    check(rc, "syn_copy_strided_f64")
    return out


def jacobi_tol(n):
    return 4.0 * (float(n) ** 0.5) * 1.1102230246251565e-16


def jacobi_rows(G, max_sweeps=40, tol=None, null_rel=1e-14):
    """In place: G (n x n contiguous, or batched) <- J G with mutually orthogonal rows (one-sided Jacobi)."""
    require_cuda_f64(G)
    batched = G.dim() == 3
    G3 = G if batched else G.unsqueeze(0)
    nb, n, n2 = G3.shape
    assert n == n2 and G3.stride(2) == 1
    cb = lib.syn_jacobi_ctrl_bytes(nb, max_sweeps)
    ctrl = workspace(cb, G.device, tag="jacobi_ctrl")
    rc = lib.syn_jacobi_rows_f64(ptr(G3), _i64(G3.stride(1)), _i64(G3.stride(0)), _i32(n), _i32(nb), ptr(ctrl), _sz(ctrl.numel() * 8),
                                 _i32(max_sweeps), _dbl(tol if tol is not None else jacobi_tol(n)), _dbl(null_rel), stream_ptr())
    check(rc, "syn_jacobi_rows_f64")
    global _last_jacobi
    _last_jacobi = (ctrl, max_sweeps, nb, G3.data_ptr())
    return G


_last_jacobi = None


def jacobi_sweeps_used():
    """Sweeps the last jacobi_rows call needed, per batch member (diagnostic; synchronises)."""
    ctrl, max_sweeps, nb, _ = _last_jacobi
    stride = lib.syn_jacobi_ctrl_stride(int(max_sweeps))
    words = ctrl.view(torch.int32)[: nb * stride].reshape(nb, stride)
    return words[:, max_sweeps + 1].tolist()


def chol_upper(G):
    """Shifted Cholesky factor of symmetric PSD matrices G (n x n or batch x n x n, contiguous rows; OVERWRITTEN when n > 128): returns
    (B, shift) with B upper triangular, G + shift I = B^T B and shift a device array (one per problem).  jacobi_rows(B) +
    jacobi_finalize(B, ..., sqrt_mode=2, shift=shift) is then the eigen-decomposition of G in fewer sweeps than jacobi_rows(G)
    (csrc/chol.cu)."""
    require_cuda_f64(G)
    batched = G.dim() == 3
    G3 = G if batched else G.unsqueeze(0)
    nb, n, n2 = G3.shape
    assert n == n2 and G3.stride(2) == 1
    B = torch.empty((nb, n, n), dtype=torch.float64, device=G.device)
    shift = torch.empty((nb,), dtype=torch.float64, device=G.device)
    check(lib.syn_chol_upper_f64(ptr(G3), _i64(G3.stride(1)), _i64(G3.stride(0)), _i32(n), _i32(nb), ptr(B), _i64(n), _i64(n * n), ptr(shift),
                                 stream_ptr()), "syn_chol_upper_f64")
    return (B, shift) if batched else (B[0], shift)


def jacobi_finalize(G, chi_max, cutoff=0.0, rank_tol=1e-14, sqrt_mode=False, shift=None):
    """Sort / normalise / cut the rows produced by jacobi_rows.  Returns (Ut, sigma, info[int32: keep, +-n], winfo[f64: discarded, s0])
    -- all DEVICE tensors; the caller decides when to synchronise on `info`.  When G is the matrix the last jacobi_rows call worked
    on, info[1] is NEGATIVE if that call ran out of sweeps before its convergence vote passed (see jacobi_solve)."""
    require_cuda_f64(G)
    batched = G.dim() == 3
    G3 = G if batched else G.unsqueeze(0)
    nb, n, _ = G3.shape
    Ut = torch.empty((nb, n, n), dtype=torch.float64, device=G.device)
    sigma = torch.empty((nb, n), dtype=torch.float64, device=G.device)
    info = torch.empty((nb, 2), dtype=torch.int32, device=G.device)
    winfo = torch.empty((nb, 2), dtype=torch.float64, device=G.device)
    mine = _last_jacobi is not None and _last_jacobi[2] == nb and _last_jacobi[3] == G3.data_ptr()
    rc = lib.syn_jacobi_finalize_f64(ptr(G3), _i64(G3.stride(1)), _i64(G3.stride(0)), _i32(n), _i32(nb), ptr(Ut), _i64(n), _i64(n * n),
                                     ptr(sigma), _i64(n), ptr(info), ptr(winfo), _i32(int(chi_max)), _dbl(cutoff), _dbl(rank_tol),
                                     _i32(int(sqrt_mode)), ptr(shift) if shift is not None else None,
                                     ptr(_last_jacobi[0]) if mine else None, _i32(_last_jacobi[1] if mine else 0), stream_ptr())
    check(rc, "syn_jacobi_finalize_f64")
    if not batched:
        return Ut[0], sigma[0], info[0], winfo[0]
    return Ut, sigma, info, winfo


JACOBI_MAX_ROUNDS = 4


def jacobi_solve(G, chi_max, cutoff=0.0, rank_tol=1e-14, sqrt_mode=False, shift=None, null_rel=1e-14, max_sweeps=40):
    """jacobi_rows + jacobi_finalize + ONE host read of `info` (kept rank and the convergence verdict of the rows kernel).
    A problem that ran out of sweeps is not truncated on: the rows kernel is called again on the same G (it resumes -- the rows are
    J G for an orthogonal J at any point) up to JACOBI_MAX_ROUNDS times, then SynError.  Returns (Ut, sigma, info ON THE HOST, winfo)."""
    for _ in range(JACOBI_MAX_ROUNDS):
        jacobi_rows(G, max_sweeps=max_sweeps, null_rel=null_rel)
        Ut, sigma, info, winfo = jacobi_finalize(G, chi_max, cutoff, rank_tol=rank_tol, sqrt_mode=sqrt_mode, shift=shift)
        h = info.cpu()
        if bool((h[..., 1] > 0).all()):
            return Ut, sigma, h, winfo
    raise SynError("one-sided Jacobi did not converge in %d x %d sweeps (n = %d)" % (JACOBI_MAX_ROUNDS, max_sweeps, G.shape[-1]))


lib.syn_dominant_subspace_workspace_f64.restype = ctypes.c_size_t
lib.syn_dominant_subspace_workspace_f64.argtypes = [_i32, _i32, _i32]


lib.syn_dominant_subspace_fused_workspace_f64.restype = ctypes.c_size_t
lib.syn_dominant_subspace_fused_workspace_f64.argtypes = [_i32, _i32, _i32, _i32]
PURIFY_FUSED = os.environ.get("SYN_PURIFY_FUSED", "1") != "0"        # experiment knob: 0 = the multi-launch sequence only


def dominant_subspace_fused_fits(n, ne):
    return PURIFY_FUSED and bool(lib.syn_dominant_subspace_fused_fits(_i32(int(n)), _i32(int(ne))))


def dominant_subspace(A, ne, sp2_iters=40, ns_iters=20, fused=None, sp2_max=160, ns_max=80):
    """Orthonormal basis U (n x ne) of the span of the `ne` dominant eigenvectors of the symmetric PSD matrix A (n x n contiguous,
    not modified) by SP2 spectral projection + Newton-Schulz (csrc/purify.cu): GEMM-bound, no host round trip.
    fused (default when n, ne are multiples of 64): one persistent cooperative kernel, iteration counts adapt on the device (at most
    sp2_max / ns_max); otherwise a fixed sequence of sp2_iters / ns_iters launches.
    Returns (U, info) with info a DEVICE vector of 8 doubles (see syngular_b200.h); nothing is synchronised here."""
    require_cuda_f64(A)
    n = A.shape[0]
    assert A.dim() == 2 and A.shape[1] == n and A.is_contiguous()
    U = torch.empty((n, int(ne)), dtype=torch.float64, device=A.device)
    info = torch.zeros((8,), dtype=torch.float64, device=A.device)
    if fused is None:
        fused = PURIFY_FUSED
    if fused and lib.syn_dominant_subspace_fused_fits(_i32(n), _i32(int(ne))):
        ws = workspace(lib.syn_dominant_subspace_fused_workspace_f64(n, int(ne), int(sp2_max), int(ns_max)), A.device, tag="purify_fused")
        check(lib.syn_dominant_subspace_fused_f64(ptr(A), _i32(n), _i32(int(ne)), _i32(int(sp2_max)), _i32(int(ns_max)), ptr(U), ptr(ws),
                                                  _sz(ws.numel() * 8), ptr(info), stream_ptr()), "syn_dominant_subspace_fused_f64")
        return U, info
    ws = workspace(lib.syn_dominant_subspace_workspace_f64(n, int(ne), int(sp2_iters)), A.device, tag="purify")
    check(lib.syn_dominant_subspace_f64(ptr(A), _i32(n), _i32(int(ne)), _i32(int(sp2_iters)), _i32(int(ns_iters)), ptr(U), ptr(ws),
                                        _sz(ws.numel() * 8), ptr(info), stream_ptr()), "syn_dominant_subspace_f64")
    return U, info


lib.syn_dominant_subspace_c128_workspace.restype = ctypes.c_size_t
lib.syn_dominant_subspace_c128_workspace.argtypes = [_i32, _i32, _i32, _i32]


def dominant_subspace_c128_fits(m, k):
    return PURIFY_FUSED and bool(lib.syn_dominant_subspace_c128_fits(_i32(int(m)), _i32(int(k))))


def dominant_subspace_c128(Hre, Him, k, sp2_max=90, ns_max=60):
    """Planar complex Hermitian PSD H (m x m) -> planar orthonormal basis (Ure, Uim) (m x k) of its k dominant eigenvectors and the 8 info
    doubles (embedded convention: traces count twice), all on the device; one fused kernel on the real embedding that computes only the
    even rows of every product (csrc/purify.cu, cx mode)."""
    require_cuda_f64(Hre); require_cuda_f64(Him)
    m = int(Hre.shape[0])
    assert Hre.shape == (m, m) and Him.shape == (m, m) and Hre.is_contiguous() and Him.is_contiguous()
    Ure = torch.empty((m, int(k)), dtype=torch.float64, device=Hre.device)
    Uim = torch.empty_like(Ure)
    info = torch.zeros((8,), dtype=torch.float64, device=Hre.device)
    ws = workspace(lib.syn_dominant_subspace_c128_workspace(m, int(k), int(sp2_max), int(ns_max)), Hre.device, tag="purify_c128")
    check(lib.syn_dominant_subspace_c128(ptr(Hre), ptr(Him), _i32(m), _i32(int(k)), _i32(int(sp2_max)), _i32(int(ns_max)), ptr(Ure), ptr(Uim),
                                         ptr(ws), _sz(ws.numel() * 8), ptr(info), stream_ptr()), "syn_dominant_subspace_c128")
    return Ure, Uim, info


def small_core_fits(rin, rout):
    return bool(lib.syn_apply_small_core_fits(_i32(int(rin)), _i32(int(rout))))


def apply_small_core(X, Wm, Y, Q, L, x_q, x_r, x_l, y_q, y_ro, y_l, w=None):
    """Y[q][ro][x] = sum_ri Wm[ro][ri] X[q][ri][x] (csrc/smallcore.cu): Wm a small contiguous (rout x rin) matrix -- or, with
    w = (rout, rin, stride_ro, stride_ri), any tensor read in place as that strided matrix -- X / Y raw strided views given by element
    strides; y_ro = (outer, inner, div) two-level index of ro.  Streaming kernel for the shared-MPO-core contractions."""
    require_cuda_f64(X, Wm, Y)
    if w is None:
        rout, rin = int(Wm.shape[0]), int(Wm.shape[1])
        assert Wm.is_contiguous()
        w_ro, w_ri = rin, 1
    else:
        rout, rin, w_ro, w_ri = (int(v) for v in w)
    check(lib.syn_apply_small_core_strided_f64(ptr(X), ptr(Wm), _i64(w_ro), _i64(w_ri), ptr(Y), _i64(int(Q)), _i32(rin), _i32(rout), _i32(int(L)),
                                               _i64(int(x_q)), _i64(int(x_r)), _i64(int(x_l)), _i64(int(y_q)), _i64(int(y_ro[0])), _i64(int(y_ro[1])),
                                               _i32(int(y_ro[2])), _i64(int(y_l)), stream_ptr()), "syn_apply_small_core_f64")
    return Y


def dominant_subspace_batched_fits(n, ne):
    return bool(lib.syn_dominant_subspace_batched_fits(_i32(int(n)), _i32(int(ne))))


def dominant_subspace_batched(A, ne, sp2_max=90, ns_max=60):
    """The projection solver for a batch of small problems: A (B x n x n contiguous, symmetric PSD) -> (U (B x n x ne), info (B x 8)),
    both on the device, one launch with one CTA per problem (csrc/purify_batched.cu).  n, ne multiples of 32, ne < n <= 128."""
    require_cuda_f64(A)
    assert A.dim() == 3 and A.shape[1] == A.shape[2] and A.is_contiguous()
    B, n = int(A.shape[0]), int(A.shape[1])
    U = torch.empty((B, n, int(ne)), dtype=torch.float64, device=A.device)
    info = torch.empty((B, 8), dtype=torch.float64, device=A.device)
    check(lib.syn_dominant_subspace_batched_f64(ptr(A), _i32(B), _i32(n), _i32(int(ne)), _i32(int(sp2_max)), _i32(int(ns_max)), ptr(U), ptr(info),
                                                stream_ptr()), "syn_dominant_subspace_batched_f64")
    return U, info


lib.syn_orthonormalize_columns_workspace_f64.restype = ctypes.c_size_t
lib.syn_orthonormalize_columns_workspace_f64.argtypes = [_i32, _i32, _i32]


def orthonormalize_columns_fits(m, q):
    return PURIFY_FUSED and bool(lib.syn_orthonormalize_columns_fits(_i32(int(m)), _i32(int(q))))


def orthonormalize_columns(A, ns_max=60):
    """Q (m x q) = A (A^T A)^(-1/2): orthonormal basis of the column space of the 2-D view A (unit column stride) by the fused
    Newton-Schulz kernel (csrc/purify.cu).  Returns (Q, info) with info on the DEVICE ([4] = max |Q^T Q - I|)."""
    require_cuda_f64(A)
    m, q = A.shape
    assert A.stride(1) == 1
    Q = torch.empty((m, q), dtype=torch.float64, device=A.device)
    info = torch.zeros((8,), dtype=torch.float64, device=A.device)
    ws = workspace(lib.syn_orthonormalize_columns_workspace_f64(m, q, int(ns_max)), A.device, tag="purify_fused")
    check(lib.syn_orthonormalize_columns_f64(ptr(A), _i64(A.stride(0)), _i32(m), _i32(q), _i32(int(ns_max)), ptr(Q), ptr(ws),
                                             _sz(ws.numel() * 8), ptr(info), stream_ptr()), "syn_orthonormalize_columns_f64")
    return Q, info


ENV_FUSED = os.environ.get("SYN_ENV_FUSED", "1") != "0"              # experiment knob: 0 = two GEMM launches


def env_sandwich_fits(l, i, o, r, b):
    return ENV_FUSED and bool(lib.syn_env_sandwich_fits(_i32(l), _i32(i), _i32(o), _i32(r), _i32(b)))


def env_sandwich(P1, W, Z, na, b):
    """Z[a,l,(l',i'),b'] = sum W[l',i',o,r'] W[l,i,o,r] P1[a,(i,r),(r',b')] in one kernel (csrc/env.cu); all operands contiguous."""
    require_cuda_f64(P1, W, Z)
    assert P1.is_contiguous() and W.is_contiguous() and Z.is_contiguous()
    l, i, o, r = W.shape
    check(lib.syn_env_sandwich_f64(ptr(P1), ptr(W), ptr(Z), _i32(int(na)), _i32(l), _i32(i), _i32(o), _i32(r), _i32(int(b)), stream_ptr()),
          "syn_env_sandwich_f64")
    return Z


def sum_parts(parts):
    """out = parts.sum(0) for a contiguous (nparts, ...) stack of split-K partial products (csrc/purify.cu: sum_parts_kernel)."""
    require_cuda_f64(parts)
    assert parts.is_contiguous() and parts.dim() >= 2
    out = torch.empty(tuple(parts.shape[1:]), dtype=torch.float64, device=parts.device)
    count = out.numel()
    check(lib.syn_sum_parts_f64(ptr(parts), _i64(count), _i32(parts.shape[0]), ptr(out), _i64(count), stream_ptr()), "syn_sum_parts_f64")
    return out


def env_mirror(E, na, L, ab):
    """Fill the strictly upper a-blocks of the symmetric environment E[(a,l),(l',a')] from the lower ones (csrc/env.cu), in place."""
    require_cuda_f64(E)
    assert E.is_contiguous()
    check(lib.syn_env_mirror_f64(ptr(E), _i32(int(na)), _i32(int(L)), _i32(int(ab)), stream_ptr()), "syn_env_mirror_f64")
    return E


def add_site(A, B, first, last):
    """Block assembly of `A + B` for one site (MPS:82-96, MPO:90-106).  Cores contiguous, physical legs flattened by the kernel."""
    require_cuda_f64(A, B)
    A, B = A.contiguous(), B.contiguous()
    la, ra, lb, rb = A.shape[0], A.shape[-1], B.shape[0], B.shape[-1]
    phys = 1
    for d in A.shape[1:-1]:
        phys *= d
    out = torch.empty(((la if first else la + lb),) + tuple(A.shape[1:-1]) + ((ra if last else ra + rb),), dtype=torch.float64, device=A.device)
    check(lib.syn_add_site_f64(ptr(A), ptr(B), ptr(out), _i32(la), _i32(ra), _i32(lb), _i32(rb), _i32(phys), _i32(int(first)), _i32(int(last)),
                               stream_ptr()), "syn_add_site_f64")
    return out


def kron_site(A, B):
    """Per-(in,out) Kronecker product of the bond matrices (MPO:140-152)."""
    require_cuda_f64(A, B)
    A, B = A.contiguous(), B.contiguous()
    la, ra, lb, rb = A.shape[0], A.shape[-1], B.shape[0], B.shape[-1]
    phys = 1
    for d in A.shape[1:-1]:
        phys *= d
    out = torch.empty((la * lb,) + tuple(A.shape[1:-1]) + (ra * rb,), dtype=torch.float64, device=A.device)
    check(lib.syn_kron_site_f64(ptr(A), ptr(B), ptr(out), _i32(la), _i32(ra), _i32(lb), _i32(rb), _i32(phys), stream_ptr()), "syn_kron_site_f64")
    return out


def sumsq(x):
    require_cuda_f64(x)
    x = x.contiguous()
    out = torch.empty((1,), dtype=torch.float64, device=x.device)
    check(lib.syn_sumsq_f64(ptr(x), _i64(x.numel()), ptr(out), stream_ptr()), "syn_sumsq_f64")
    return out


def scale_rsqrt_(x, ss):
    require_cuda_f64(x, ss)
    assert x.is_contiguous()
    check(lib.syn_scale_rsqrt_f64(ptr(x), _i64(x.numel()), ptr(ss), stream_ptr()), "syn_scale_rsqrt_f64")
    return x


def bias_act_(y, bias, act="relu"):
    """In place: y[r, c] = act(y[r, c] + bias[c]) (TensorDense epilogue)."""
    require_cuda_f64(y)
    assert y.is_contiguous() and y.dim() == 2
    code = {"relu": 1, None: 0, "linear": 0, "identity": 0}[act]
    b = ptr(bias) if bias is not None else None
    check(lib.syn_bias_act_f64(ptr(y), b, _i64(y.shape[0]), _i32(y.shape[1]), _i32(code), stream_ptr()), "syn_bias_act_f64")
    return y


lib.syn_apply_round_chain_workspace_f64.restype = ctypes.c_size_t
lib.syn_round_chain_workspace_f64.restype = ctypes.c_size_t


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def _int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


def apply_round_chain(X, W, dim):
    """`W @ X` + strict `>> dim` (reference semantics) over the whole chain in ONE library call (csrc/chain.cu).  X: list of contiguous
    CUDA float64 MPS cores (a,i,b); W: MPO cores (l,i,o,r).  Returns the list of rounded cores."""
    n = len(X)
    require_cuda_f64(*X)
    require_cuda_f64(*W)
    X = [x.contiguous() for x in X]
    W = [w.contiguous() for w in W]
    xs = _int_array([d for x in X for d in x.shape])
    wsh = _int_array([d for w in W for d in w.shape])
    oshape = (ctypes.c_int * (3 * n))()
    check(lib.syn_apply_round_chain_shapes(_i32(n), xs, wsh, _i32(int(dim)), oshape), "syn_apply_round_chain_shapes")
    dev = X[0].device
    out = [torch.empty((oshape[3 * k], oshape[3 * k + 1], oshape[3 * k + 2]), dtype=torch.float64, device=dev) for k in range(n)]
    ws = workspace(lib.syn_apply_round_chain_workspace_f64(_i32(n), xs, wsh, _i32(int(dim))), dev, tag="chain")
    check(lib.syn_apply_round_chain_f64(_i32(n), _ptr_array(X), xs, _ptr_array(W), wsh, _i32(int(dim)), _ptr_array(out), ptr(ws),
                                        _sz(ws.numel() * 8), stream_ptr()), "syn_apply_round_chain_f64")
    return out


def round_chain(sites, dim):
    """Strict `>> dim` sweep (reference semantics) over the whole chain in ONE library call; cores (l, phys..., r) contiguous CUDA float64."""
    n = len(sites)
    require_cuda_f64(*sites)
    cores = [c.contiguous() for c in sites]
    flat = []
    for c in cores:
        d = 1
        for e in c.shape[1:-1]:
            d *= int(e)
        flat += [int(c.shape[0]), d, int(c.shape[-1])]
    sh = _int_array(flat)
    oshape = (ctypes.c_int * (3 * n))()
    check(lib.syn_round_chain_shapes(_i32(n), sh, _i32(int(dim)), oshape), "syn_round_chain_shapes")
    dev = cores[0].device
    out = [torch.empty((oshape[3 * k],) + tuple(cores[k].shape[1:-1]) + (oshape[3 * k + 2],), dtype=torch.float64, device=dev) for k in range(n)]
    ws = workspace(lib.syn_round_chain_workspace_f64(_i32(n), sh, _i32(int(dim))), dev, tag="chain")
    check(lib.syn_round_chain_f64(_i32(n), _ptr_array(cores), sh, _i32(int(dim)), _ptr_array(out), ptr(ws), _sz(ws.numel() * 8), stream_ptr()),
          "syn_round_chain_f64")
    return out


lib.syn_tt_dense3_packed_floats.restype = ctypes.c_size_t


def tt_dense3_fits(tt_input_shape, tt_output_shape, tt_bond_shape):
    """Shapes covered by the fused tcgen05 TF32 kernel (csrc/ttdense.cu): three cores, every mode and bond equal to 16."""
    return tuple(tt_input_shape) == (16, 16, 16) and tuple(tt_output_shape) == (16, 16, 16) and tuple(tt_bond_shape) == (16, 16)


def tt_dense3_pack(G1, G2, G3):
    """Cores in the reference's layouts (float32 CUDA: (16,16,16), (16,16,16,16), (16,16,16)) -> the pre-swizzled operand images."""
    for g, shape in ((G1, (16, 16, 16)), (G2, (16, 16, 16, 16)), (G3, (16, 16, 16))):
        if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and tuple(g.shape) == shape):
            raise SynError("tt_dense3_pack: expected contiguous CUDA float32 cores of extent 16, got %r %r" % (tuple(g.shape), g.dtype))
    packed = torch.empty(int(lib.syn_tt_dense3_packed_floats()), dtype=torch.float32, device=G1.device)
    check(lib.syn_tt_dense3_pack_tf32(ptr(G1), ptr(G2), ptr(G3), ptr(packed), stream_ptr()), "syn_tt_dense3_pack_tf32")
    return packed


def tt_dense3_tf32(x, packed, bias=None, relu=True, out=None, pair=False):
    """y = act(TT-matvec(x) + bias) for a (batch, 4096) float32 CUDA input, on tcgen05.mma.kind::tf32 (one fused launch)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == 4096):
        raise SynError("tt_dense3_tf32: expected a contiguous CUDA float32 (batch, 4096) input")
    if out is None:
        out = torch.empty_like(x)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape == x.shape
    if bias is not None:
        assert bias.is_cuda and bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == 4096
    fn = lib.syn_tt_dense3_tf32_pair if pair else lib.syn_tt_dense3_tf32            # pair: the cta_group::2 variant (checked alternative)
    check(fn(ptr(x), ptr(packed), ptr(bias) if bias is not None else None, ptr(out), _i32(x.shape[0]), _i32(1 if relu else 0), stream_ptr()),
          "syn_tt_dense3_tf32")
    return out


def identity_deviation(X):
    require_cuda_f64(X)
    out = torch.empty((1,), dtype=torch.float64, device=X.device)
    check(lib.syn_identity_deviation_f64(ptr(X.contiguous()), _i32(X.shape[0]), ptr(out), stream_ptr()), "syn_identity_deviation_f64")
    return out


# ---------------------------------------------------------------------------------------------------------
# fused transfer-matrix inner product (csrc/overlap.cu)
class OverlapSite(ctypes.Structure):
    _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("a_stride", _i64), ("b_stride", _i64),
                ("la", _i32), ("ra", _i32), ("lb", _i32), ("rb", _i32), ("d", _i32), ("_pad", _i32)]


OVERLAP_MAX_SITES = 64
OVERLAP_FUSED = os.environ.get("SYN_OVERLAP_FUSED", "1") != "0"      # experiment knob: 0 = GEMM-per-site route only
lib.syn_overlap_batched_fits.argtypes = [ctypes.POINTER(OverlapSite), _i32]
lib.syn_overlap_batched_f64.argtypes = [ctypes.POINTER(OverlapSite), _i32, _i32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]


def _overlap_sites(a_cores, b_cores, batched):
    """Site table of syn_overlap_batched_f64 for cores (B, l, d, r) (batched; B = 1 on one side means a shared chain) or (l, d, r)."""
    n = len(a_cores)
    tab = (OverlapSite * n)()
    for k, (a, b) in enumerate(zip(a_cores, b_cores)):
        require_cuda_f64(a, b)
        if not (a.is_contiguous() and b.is_contiguous()):
            raise _lib.SynError("overlap: cores must be contiguous")
        sa, sb = (a.shape[1:], b.shape[1:]) if batched else (a.shape, b.shape)
        if len(sa) != 3 or len(sb) != 3 or sa[1] != sb[1]:
            raise _lib.SynError("overlap: cores must be (l, d, r) with equal physical dimensions, got %r and %r" % (tuple(sa), tuple(sb)))
        ast = sa[0] * sa[1] * sa[2] if (batched and a.shape[0] > 1) else 0
        bst = sb[0] * sb[1] * sb[2] if (batched and b.shape[0] > 1) else 0
        tab[k] = OverlapSite(a.data_ptr(), b.data_ptr(), ast, bst, sa[0], sa[2], sb[0], sb[2], sa[1], 0)
    return tab


def overlap_fits(a_cores, b_cores, batched=True):
    """True when the fused kernel takes these chains (bonds <= 64, round8(l) * d <= 128, d * r <= 128)."""
    if not OVERLAP_FUSED or len(a_cores) != len(b_cores) or len(a_cores) == 0:
        return False
    tab = _overlap_sites(a_cores, b_cores, batched)
    n = len(a_cores)
    return all(lib.syn_overlap_batched_fits(ctypes.cast(ctypes.byref(tab, k0 * ctypes.sizeof(OverlapSite)), ctypes.POINTER(OverlapSite)),
                                            min(OVERLAP_MAX_SITES, n - k0)) != 0 for k0 in range(0, n, OVERLAP_MAX_SITES))


def overlap_batched(a_cores, b_cores, batched=True):
    """E (batch, r_a, r_b) of the transfer-matrix chain  E <- sum A[a,i,b] E[a,a'] B[a',i,b']  over all sites (MPS:116-129) in ONE
    launch per 64 sites; cores are lists of contiguous (B, l, d, r) tensors (or (l, d, r) with batched=False)."""
    n = len(a_cores)
    tab = _overlap_sites(a_cores, b_cores, batched)
    batch = max(int(a_cores[0].shape[0]), int(b_cores[0].shape[0])) if batched else 1
    dev = a_cores[0].device
    if tab[0].la == 1 and tab[0].lb == 1:
        E = None
    else:
        raise _lib.SynError("overlap: the first bonds must be 1")
    for k0 in range(0, n, OVERLAP_MAX_SITES):
        m = min(OVERLAP_MAX_SITES, n - k0)
        out = torch.empty((batch, tab[k0 + m - 1].ra, tab[k0 + m - 1].rb), dtype=torch.float64, device=dev)
        seg = ctypes.cast(ctypes.byref(tab, k0 * ctypes.sizeof(OverlapSite)), ctypes.POINTER(OverlapSite))
        check(lib.syn_overlap_batched_f64(seg, m, batch, ptr(E) if E is not None else None, ptr(out), stream_ptr()),
              "syn_overlap_batched_f64")
        E = out
    return E
