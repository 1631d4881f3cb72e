"""Host-side sweep algorithms of the matrix-product hot path, on lists of CUDA float64 cores.

Pure orchestration: every arithmetic step is a call into libsyngular_b200.so through `syngular_b200.ops`
(strided DMMA GEMM, Householder qrt, one-sided Jacobi, block assembly).  PyTorch only provides device buffers
and views.  Citations are to the reference (MPS = tensor/matrix_product_state.py, MPO = tensor/matrix_product_operator.py).
"""
import numpy as np
import torch

from syngular_b200 import cplx, ops
from syngular_b200.cplx import Cx

F64 = torch.float64
C128 = torch.complex128


def device():
    if not torch.cuda.is_available():
        raise ops._lib.SynError("syngular_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def as_core(x):
    """numpy array / torch tensor -> contiguous CUDA float64 tensor (host arrays are uploaded)."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if np.iscomplexobj(a):
            if a.size and np.max(np.abs(a.imag)) > 0:
                t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))
                return t.to(device()).contiguous()
            a = a.real
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if t.dtype not in (F64, C128):
        t = t.to(C128 if t.is_complex() else F64)
    if not t.is_cuda:
        # pinned host tensors upload asynchronously on the current stream (the kernels that read them are queued behind the copy; as with
        # any non_blocking copy the caller must not overwrite the pinned buffer before the stream has passed it); pageable memory blocks
        t = t.to(device(), non_blocking=t.is_pinned())
    return t.contiguous()


def empty(*shape):
    return torch.empty(shape, dtype=F64, device=device())


# ---- complex128 cores: planar operands on the same kernels, see syngular_b200/cplx.py -------------------------
def any_complex(*chains):
    """True when any core / tensor of the given lists (or single tensors) is complex128."""
    for ch in chains:
        for t in (ch if isinstance(ch, (list, tuple)) else (ch,)):
            if t is not None and (isinstance(t, Cx) or t.dtype == C128):
                return True
    return False


def lift(x):
    """torch core(s) -> planar complex operand(s)."""
    if isinstance(x, (list, tuple)):
        return [lift(t) for t in x]
    return x if isinstance(x, Cx) else Cx.from_torch(x)


def lower(x):
    """planar complex operand(s) -> torch complex128 core(s)."""
    if isinstance(x, (list, tuple)):
        return [lower(t) for t in x]
    return x.to_torch() if isinstance(x, Cx) else x


def _O(x):
    """Kernel namespace for an operand: the real wrappers, or their 4-launch complex compositions."""
    return cplx if isinstance(x, Cx) else ops


def empty_for(x, *shape):
    return Cx.empty(shape, x.device) if isinstance(x, Cx) else torch.empty(shape, dtype=F64, device=x.device)


def ones_for(x, *shape):
    return Cx.ones(shape, x.device) if isinstance(x, Cx) else torch.ones(shape, dtype=F64, device=x.device)


def mirror(sites):
    """Reverse the chain and swap the bond legs of every core (physical legs flattened): right sweeps become left sweeps."""
    return [s.reshape(s.shape[0], -1, s.shape[-1]).permute(2, 1, 0).contiguous() for s in reversed(sites)]


def unmirror(swept, like):
    out = []
    for s, old in zip(reversed(swept), like):
        t = s.permute(2, 1, 0).contiguous()
        out.append(t.reshape((t.shape[0],) + tuple(old.shape[1:-1]) + (t.shape[2],)))
    return out


def _is_operand(a):
    return isinstance(a, (torch.Tensor, Cx)) or (isinstance(a, (list, tuple)) and len(a) > 0 and isinstance(a[0], (torch.Tensor, Cx)))


def _lower_result(r):
    if isinstance(r, Cx):
        return r.to_torch()
    if isinstance(r, list):
        return [_lower_result(x) for x in r]
    if isinstance(r, tuple):
        return tuple(_lower_result(x) for x in r)
    return r


def complex_aware(fn):
    """Entry points called with torch cores: when any operand is complex128, every tensor operand is converted to planar form
    (real ones get a zero imaginary part), the generic body runs on the complex compositions, and results come back complex128."""
    def wrapper(*args, **kw):
        operands = [a for a in args if _is_operand(a)]
        if not any_complex(*operands) or any(isinstance(a, Cx) or (isinstance(a, (list, tuple)) and isinstance(a[0], Cx)) for a in operands):
            return fn(*args, **kw)
        return _lower_result(fn(*[lift(a) if _is_operand(a) else a for a in args], **kw))
    wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
    return wrapper


# ---------------------------------------------------------------------------------------------------------
# site contractions (K1 / K2 of SURVEY section 2.4) as single strided GEMMs
# ---------------------------------------------------------------------------------------------------------
@complex_aware
def site_mpo_mps(X, W):
    """C[(a,l), o, (b,r)] = sum_i X[a,i,b] W[l,i,o,r]   (MPO:184-190; MPS bond major, MPO bond minor)."""
    a, i, b = X.shape
    l, i2, o, r = W.shape
    assert i == i2, (X.shape, W.shape)
    out = empty_for(X, a * l, o, b * r)
    _O(X).gemm(X, W, out, M=b, N=o * r, K=i,
             a_m=1, a_k=b, b_k=o * r, b_n=1,
             c_m=r, c_n=(b * r, 1, r),
             batch=a * l, a_b=(i * b, 0, l), b_b=(0, i * o * r, l), c_b=o * b * r)
    return out


@complex_aware
def site_mpo_mpo(A, B):
    """C[(lA,lB), iA, oB, (rA,rB)] = sum_x A[lA,iA,x,rA] B[lB,x,oB,rB]   (MPO:280-287; A acts first)."""
    la, ia, xa, ra = A.shape
    lb, xb, ob, rb = B.shape
    assert xa == xb, (A.shape, B.shape)
    out = empty_for(A, la * lb, ia, ob, ra * rb)
    _O(A).gemm(A, B, out, M=ia * ra, N=ob * rb, K=xa,
             a_m=(xa * ra, 1, ra), a_k=ra, b_k=ob * rb, b_n=1,
             c_m=(ob * ra * rb, rb, ra), c_n=(ra * rb, 1, rb),
             batch=la * lb, a_b=(ia * xa * ra, 0, lb), b_b=(0, xb * ob * rb, lb), c_b=ia * ob * ra * rb)
    return out


# ---------------------------------------------------------------------------------------------------------
# reference-semantic sweeps
# ---------------------------------------------------------------------------------------------------------
ORTHO_MIN_ROWS = 256           # truncation steps with at least this many rows try the fused Newton-Schulz kernel first (0 = Householder only)


def qrt_step(L, q, want_S=True):
    """The reference truncation step `Q, R = np.linalg.qr(L, "complete"); Q[:, :q]; R[:q, :]` (MPS:443-446, MPO:555-558) = projection
    of L on the span of its first q columns.  Any orthonormal basis of that span gives the same projection, so large blocks take the
    GEMM-bound polar orthonormalisation Q = Lq (Lq^T Lq)^(-1/2) (fused Newton-Schulz kernel, csrc/purify.cu) and fall back to the
    Householder kernel when it does not converge (rank-deficient leading columns) or the shape is not covered."""
    if isinstance(L, Cx):
        return cplx.qrt(L, q, want_S=want_S)
    m, n = L.shape
    q = int(q)
    if (ORTHO_MIN_ROWS and m >= ORTHO_MIN_ROWS and q <= n and L.dim() == 2 and L.stride(1) == 1 and L.data_ptr() % 16 == 0
            and ops.orthonormalize_columns_fits(m, q)):
        Q, info = ops.orthonormalize_columns(L[:, :q])
        h = info.cpu()
        if bool(torch.isfinite(h).all()) and float(h[4]) < 1e-12:
            return Q, (ops.matmul(Q.t(), L) if want_S else None)
    return ops.qrt(L, q, want_S=want_S)


CHAIN_CALLS = True             # real float64 chains run their `>>` / `@` + `>>` sweeps as ONE library call (csrc/chain.cu)


def _real_cuda_chain(*chains):
    return CHAIN_CALLS and all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == F64 for ch in chains for t in ch)


@complex_aware
def round_qr(sites, dim):
    """Strict `>>` sweep (MPS:432-468, MPO:544-580): QR-truncation left to right, no canonicalisation.
    Natural clamp: kept width = min(dim, rows).  Real chains: one call of syn_round_chain_f64; the loop below is the same sweep
    step by step for planar complex operands (and the CPU-emulated host tests)."""
    if len(sites) > 1 and _real_cuda_chain(sites):
        return ops.round_chain(sites, int(dim))
    out = list(sites)
    for k in range(len(out) - 1):
        cur, nxt = out[k], out[k + 1]
        L = cur.reshape(-1, cur.shape[-1])
        Q, S = qrt_step(L, dim)
        kept = Q.shape[1]
        Wn = _O(S).matmul(S, nxt.reshape(nxt.shape[0], -1))
        out[k] = Q.reshape(tuple(cur.shape[:-1]) + (kept,))
        out[k + 1] = Wn.reshape((kept,) + tuple(nxt.shape[1:]))
    return out


@complex_aware
def left_orthonormalize(sites):
    """MPS:554-566 / MPO:673-694 (reduced QR, left to right)."""
    out = list(sites)
    for k in range(len(out) - 1):
        cur, nxt = out[k], out[k + 1]
        L = cur.reshape(-1, cur.shape[-1])
        Q, S = _O(L).qrt(L, min(L.shape))
        kept = Q.shape[1]
        Wn = _O(S).matmul(S, nxt.reshape(nxt.shape[0], -1))
        out[k] = Q.reshape(tuple(cur.shape[:-1]) + (kept,))
        out[k + 1] = Wn.reshape((kept,) + tuple(nxt.shape[1:]))
    return out


@complex_aware
def right_orthonormalize(sites):
    """MPS:568-580 / MPO:696-719: QR of R^T right to left; the transposed factor is written straight into the core."""
    if any_complex(sites):                                 # planar operands: the mirrored left sweep (same factorisation)
        return unmirror(left_orthonormalize(mirror(sites)), sites)
    out = list(sites)
    for k in range(len(out) - 1, 0, -1):
        cur, prv = out[k], out[k - 1]
        R = cur.reshape(cur.shape[0], -1)                  # (l, d*r)
        kept = min(R.shape)
        core = empty(kept, R.shape[1])
        _, S = ops.qrt(R.t(), kept, Q=core.t())            # Q = core^T ; S = Q^T R^T  (kept x l)
        Lp = prv.reshape(-1, prv.shape[-1])
        Wp = ops.matmul(Lp, S.t())
        out[k] = core.reshape((kept,) + tuple(cur.shape[1:]))
        out[k - 1] = Wp.reshape(tuple(prv.shape[:-1]) + (kept,))
    return out


@complex_aware
def overlap(A, B):
    """`A | B` (MPS:116-129): bilinear transfer-matrix contraction.  Chains with bonds <= 64 run as ONE launch of the fused kernel
    (csrc/overlap.cu: the transfer matrix never leaves shared memory); larger bonds take two GEMMs per site.  Returns a (1,1)
    device tensor."""
    cx = isinstance(A[0], Cx)
    if (not cx and A[-1].shape[-1] == 1 and B[-1].shape[-1] == 1 and A[0].shape[0] == 1 and B[0].shape[0] == 1
            and ops.overlap_fits(A, B, batched=False)):
        return ops.overlap_batched(A, B, batched=False).reshape(1, 1)
    O = _O(A[0])
    E = ones_for(A[0], 1, 1)
    for a, b in zip(A, B):
        la, lb = a.shape[0], b.shape[0]
        T = O.matmul(E.t(), a.reshape(la, -1))             # (lb, d*ra')
        ra, rb = a.shape[-1], b.shape[-1]
        E = O.matmul(T.reshape(-1, ra).t(), b.reshape(-1, rb))     # (ra', rb')
    return E


@complex_aware
def to_dense(sites):
    """Dense tensor by one left-to-right chain of GEMMs (the reference loops over every index: MPS:265-273, MPO:405-416)."""
    T = sites[0].reshape(-1, sites[0].shape[-1])
    dims = list(sites[0].shape[1:-1])
    for c in sites[1:]:
        T = _O(T).matmul(T, c.reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
        dims += list(c.shape[1:-1])
    T = T.reshape(dims)
    if sites[0].dim() == 4:
        n = len(sites)
        T = T.permute(list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))
    return T


@complex_aware
def retrieve(sites, idx_in, idx_out=None):
    """One amplitude = product of the sliced bond matrices (MPS:546-551, MPO:663-670); (1,1) device tensor."""
    v = None
    for k, c in enumerate(sites):
        m = c[:, int(idx_in[k]), :] if idx_out is None else c[:, int(idx_in[k]), int(idx_out[k]), :]
        v = m if v is None else _O(v).matmul(v, m)
    return v


@complex_aware
def add_site(A, B, first, last):
    """Block assembly of `A + B` for one site (MPS:82-96, MPO:90-106)."""
    return _O(A).add_site(A, B, first, last)


@complex_aware
def gram(M, side):
    """Gram matrix of an unfolding: side="left" -> M^H M (columns), "right" -> M M^H (rows)  (MPS:624-630, MPO:771-777)."""
    Mh = M.h() if isinstance(M, Cx) else M.t()
    return _O(M).matmul(Mh, M) if side == "left" else _O(M).matmul(M, Mh)


def normalize_last(core):
    """New last core divided by its Frobenius norm (MPS:252-256); complex cores through their float64 (..., 2) view."""
    last = core.clone()
    flat = torch.view_as_real(last) if last.dtype == C128 else last
    ops.scale_rsqrt_(flat, ops.sumsq(flat))
    return last


def apply_gate(sites, gate, index, mode="compress", chi_max=None, cutoff=0.0):
    """`MatrixProductState.apply` (MPS:487-534): contract a dense m-site gate (legs out..., in...; column-vector convention)
    into cores index..index+m-1, then re-split.
      mode="compress" (reference): qrt split keeping the EXISTING bonds -- the bond never grows;
      mode="svd" (extension): local SVD split, bonds grow up to chi_max (relative cutoff `cutoff`).
    Complex gates or cores (quantum/gate.py:9-13) run planar on the same kernels (syngular_b200/cplx.py); only the m cores
    touched are converted, and the result cores are complex128."""
    out = list(sites)
    m = gate.dim() // 2
    touched = out[index:index + m]
    cx = any_complex(touched, gate)
    if cx:
        touched, gate = lift(touched), lift(gate)
    T = touched[0]
    O = _O(T)
    for c in touched[1:]:                                               # merge the m cores: (l, in_0..in_j, r)
        T = O.matmul(T.reshape(-1, T.shape[-1]), c.reshape(c.shape[0], -1)).reshape(tuple(T.shape[:-1]) + tuple(c.shape[1:]))
    l, r = T.shape[0], T.shape[-1]
    dims_out = tuple(gate.shape[:m])
    dout, din = int(np.prod(dims_out)), int(np.prod(gate.shape[m:]))
    G = gate.reshape(dout, din)
    T3 = T.reshape(l, din, r)
    Tn = empty_for(T3, l, dout, r)
    # Tn[l] (dout x r) = G (dout x din) @ T3[l] (din x r), batched over l with the gate shared
    O.gemm(G, T3, Tn, M=dout, N=r, K=din, a_m=din, a_k=1, b_k=r, b_n=1, c_m=r, c_n=1, batch=l, a_b=0, b_b=din * r, c_b=dout * r)
    T = Tn.reshape((l,) + dims_out + (r,))
    trunc = Truncation()
    new = []
    for k in range(index, index + m - 1):
        lr, d = T.shape[0], T.shape[1]
        L = T.reshape(lr * d, -1)
        if mode == "svd":
            core2d, kept = _svd_basis(L, chi_max if chi_max is not None else min(L.shape), cutoff, trunc)
            S = O.matmul(core2d.h() if cx else core2d.t(), L)
        else:
            core2d, S = O.qrt(L, int(sites[k].shape[2]))
            kept = core2d.shape[1]
        new.append(core2d.reshape(lr, d, kept))
        T = S.reshape((kept,) + tuple(T.shape[2:]))
    new.append(T.reshape(T.shape[0], -1, T.shape[-1]))
    out[index:index + m] = lower(new) if cx else new
    return out


def mpo_apply_range(sites, op_sites, indices):
    """`MatrixProductOperator.apply(operator, indices)` with an MPO operator (MPO:582-626): one strided GEMM per site forms the
    product cores, a chain of GEMMs contracts the range, and the block is re-split by the qrt step keeping the target's right
    bonds.  The reference re-splits the block with its legs ordered (l, in..., out..., r) as if they were interleaved (MPO:609):
    that literal behaviour is what is computed (golden: tests/golden/mpo_apply.npz).  Returns the new list of cores."""
    out = list(sites)
    m = len(indices)
    if op_sites[0].shape[0] != 1 or op_sites[-1].shape[-1] != 1:
        raise ValueError("cannot select an axis to squeeze out which has size not equal to one")      # np.squeeze, MPO:602
    T, dims = None, []
    for idx, jdx in enumerate(indices):
        c = site_mpo_mpo(out[jdx], op_sites[idx])                       # ((l,l'), in, out', (r,r'))
        T = c.reshape(-1, c.shape[-1]) if T is None else ops.matmul(T, c.reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
        dims += [c.shape[1], c.shape[2]]
    l0, r_last = out[indices[0]].shape[0], out[indices[-1]].shape[-1]
    T = T.reshape([l0] + dims + [r_last])
    T = T.permute([0] + [1 + 2 * k for k in range(m)] + [2 + 2 * k for k in range(m)] + [2 * m + 1]).contiguous()
    for idx in range(m - 1):
        jdx = indices[idx]
        l, i, o = out[jdx].shape[0], out[jdx].shape[1], op_sites[idx].shape[2]
        Q, S = ops.qrt(T.reshape(l * i * o, -1), int(out[jdx].shape[3]))
        out[jdx] = Q.reshape(l, i, o, Q.shape[1])
        T = S
    jdx = indices[-1]
    out[jdx] = T.reshape(out[jdx].shape[0], out[jdx].shape[1], op_sites[m - 1].shape[2], out[jdx].shape[3])
    return out


@complex_aware
def crumble(left, right):
    """Merged operator core (l, i, o, i', o', r') of two neighbouring sites (differential_matrix_product_operator.py:13-27)."""
    l, i, o, r = left.shape
    _, i2, o2, r2 = right.shape
    return _O(left).matmul(left.reshape(l * i * o, r), right.reshape(r, i2 * o2 * r2)).reshape(l, i, o, i2, o2, r2)


@complex_aware
def project_wings(state, operator, index):
    """Wings of W @ X around the operator cores index, index + 1 (tensor/differential_matrix_product_operator.py:104-163):
    left (1, prod(out_0..out_{index-1}), a, w) and right (a, w, prod(out_{index+2}..), 1), MPS bond before MPO bond.  Each site is the
    `@` contraction (one strided GEMM, product bonds flattened MPS-bond major) and the chain is multiplied up with one GEMM per site."""
    n = len(state)
    left = None                                            # (prod(out), D) with D = (a, w) flattened
    for k in range(index):
        C = site_mpo_mps(state[k], operator[k])            # ((a,w), o, (b,v))
        D0, o, D1 = C.shape
        left = C.reshape(D0 * o, D1) if left is None else _O(left).matmul(left, C.reshape(D0, o * D1)).reshape(-1, D1)
    right = None                                           # (D, prod(out))
    for k in range(n - 1, index + 1, -1):
        C = site_mpo_mps(state[k], operator[k])
        D0, o, D1 = C.shape
        right = C.reshape(D0, o * D1) if right is None else _O(right).matmul(C.reshape(D0 * o, D1), right).reshape(D0, -1)
    a, w = state[index].shape[0], operator[index].shape[0]
    a2, w2 = state[index + 1].shape[2], operator[index + 1].shape[3]
    return left.reshape(1, -1, a, w), right.reshape(a2, w2, -1, 1)


# ---------------------------------------------------------------------------------------------------------
# three-layer transfer contractions: the DMRG environment blocks (variational/dmrg.py:65-112)
# ---------------------------------------------------------------------------------------------------------
@complex_aware
def dmrg_right_blocks(state, operator):
    """R_k[a, w, a'] = sum S_k[a,o,b] W_k[w,i,o,v] S_k[a',i,b'] R_{k+1}[b,v,b'] for k = n-1 .. 2 (dmrg.py:65-87; no conjugation, the
    first state copy meets the operator's OUT leg): three GEMMs per site -- the `|` transfer contraction with an MPO in the middle."""
    n = len(state)
    blocks = [None] * n
    R = ones_for(state[0], 1, 1, 1)
    for k in range(n - 1, 1, -1):
        S, W = state[k], operator[k]
        O = _O(S)
        a, d, b = S.shape
        w, i, o, v = W.shape
        # T1[(a,o), (v,b')] = sum_b S[(a,o), b] R[b, (v,b')]
        T1 = O.matmul(S.reshape(a * d, b), R.reshape(b, -1))
        bp = R.shape[2]
        # T2[a, (w,i), b'] = sum_{(o,v)} W[(w,i), (o,v)] T1[a, (o,v), b']            batched over a
        T2 = empty_for(S, a, w * i, bp)
        O.gemm(W, T1, T2, M=w * i, N=bp, K=o * v, a_m=o * v, a_k=1, b_k=bp, b_n=1, c_m=bp, c_n=1,
               batch=a, a_b=0, b_b=o * v * bp, c_b=w * i * bp)
        # R_k[(a,w), a'] = sum_{(i,b')} T2[(a,w), (i,b')] S[a', (i,b')]
        Sm = S.reshape(a, d * b)
        R = O.matmul(T2.reshape(a * w, i * bp), Sm.t()).reshape(a, w, a)
        blocks[k] = R
    return blocks


@complex_aware
def dmrg_left_blocks(state, operator):
    """L_k[b, v, b'] = sum L_{k-1}[a,w,a'] S_k[a,o,b] W_k[w,i,o,v] S_k[a',i,b'] for k = 0 .. n-3 (dmrg.py:90-112)."""
    n = len(state)
    blocks = [None] * n
    L = ones_for(state[0], 1, 1, 1)
    for k in range(n - 2):
        S, W = state[k], operator[k]
        O = _O(S)
        a, d, b = S.shape
        w, i, o, v = W.shape
        ap = L.shape[2]
        # T1[(w,a'), (o,b)] = sum_a L[a, (w,a')]^T S[a, (o,b)]
        T1 = O.matmul(L.reshape(a, w * ap).t(), S.reshape(a, d * b))
        # T2[a', (i,v), b] = sum_{(w,o)} W[w,i,o,v] T1[w, a', o, b]                       batched over a'
        T2 = empty_for(S, ap, i * v, b)
        O.gemm(W, T1, T2, M=i * v, N=b, K=w * o, a_m=(o * v, 1, v), a_k=(i * o * v, v, o), b_k=(ap * o * b, b, o), b_n=1, c_m=b, c_n=1,
               batch=ap, a_b=0, b_b=o * b, c_b=i * v * b)
        # G[(v,b), b'] = sum_{(a',i)} T2[a', i, v, b] S[a', i, b'] ;  L_k[b, v, b'] = G[v, b, b']
        G = empty_for(S, v * b, b)
        O.gemm(T2, S, G, M=v * b, N=b, K=ap * i, a_m=1, a_k=(i * v * b, v * b, i), b_k=b, b_n=1, c_m=b, c_n=1)
        L = G.reshape(v, b, b).permute(1, 0, 2).contiguous()
        blocks[k] = L
    return blocks


@complex_aware
def decompose_left(T, shapes):
    """TT decomposition by the qrt step (MPS:298-319, MPO:430-450)."""
    cores, l, n = [], 1, len(shapes)
    T = T.reshape(-1)
    for k in range(n - 1):
        phys = int(np.prod(shapes[k][1:-1]))
        L = T.reshape(l * phys, -1)
        Q, S = _O(L).qrt(L, int(shapes[k][-1]))
        kept = Q.shape[1]
        cores.append(Q.reshape((l,) + tuple(shapes[k][1:-1]) + (kept,)))
        T, l = S, kept
    cores.append(T.reshape((l,) + tuple(shapes[n - 1][1:-1]) + (1,)))
    return cores


def decompose_right(T, shapes):
    """MPS-only right-to-left variant (MPS:324-347)."""
    n = len(shapes)
    cores, r = [None] * n, 1
    T = T.reshape(-1)
    for k in range(n - 1, 0, -1):
        phys = int(np.prod(shapes[k][1:-1]))
        Rm = T.reshape(-1, phys * r)                        # (rest, d*r)
        rest = Rm.shape[0]
        kept = min(int(shapes[k][0]), phys * r)
        core = empty(kept, phys * r)
        Snext = empty(rest, kept)                           # S^T, so the remainder stays (rest, kept) row-major
        ops.qrt(Rm.t(), kept, Q=core.t(), S=Snext.t())
        cores[k] = core.reshape((kept,) + tuple(shapes[k][1:-1]) + (r,))
        T, r = Snext, kept
    cores[0] = T.reshape((1,) + tuple(shapes[0][1:-1]) + (r,))
    return cores


# ---------------------------------------------------------------------------------------------------------
# fused MPO x MPS application: the product core (a*l, o, b*r) is never materialised
# ---------------------------------------------------------------------------------------------------------
SMALL_CORE = True      # the carry x MPO-core contraction on the streaming kernel when the core is at most 16 x 16 (False: GEMM)


def contract_carry(T, X, W):
    """M[s, o, (b,r)] = sum_{a,l,i} T[s,l,a] X[a,i,b] W[l,i,o,r].
    `T` is the carry of the sweep, stored (s, l, a) -- MPO bond major -- so that both GEMMs read unit-stride operands;
    the output columns are in the reference's product order (b major, r minor), which the QR truncation depends on."""
    s, l, a = T.shape
    a2, i, b = X.shape
    l2, i2, o, r = W.shape
    assert a == a2 and l == l2 and i == i2, (T.shape, X.shape, W.shape)
    T1 = empty(s, l, i, b)
    # T1[s,l,(i,b)] = sum_a T[s,l,a] X[a,(i,b)]      batch over l
    ops.gemm(T, X, T1, M=s, N=i * b, K=a, a_m=l * a, a_k=1, b_k=i * b, b_n=1, c_m=l * i * b, c_n=1,
             batch=l, a_b=a, b_b=0, c_b=i * b)
    M = empty(s, o, b * r)
    # M[(s,b),(o,r)] = sum_{(l,i)} T1[s,(l,i),b] W[(l,i),(o,r)]
    if SMALL_CORE and ops.small_core_fits(l * i, o * r) and W.is_contiguous() and s * b >= 4096:
        # an MPO core of at most 16 x 16 against s b columns: streaming kernel, the core read in place as its transpose (csrc/smallcore.cu;
        # at 32 x 32 -- the C2 chain -- the DMMA GEMM is faster: 22 vs 40 us)
        ops.apply_small_core(T1, W, M, Q=s, L=b, x_q=l * i * b, x_r=b, x_l=1, y_q=o * b * r, y_ro=(b * r, 1, r), y_l=r,
                             w=(o * r, l * i, 1, o * r))
    else:
        ops.gemm(T1, W, M, M=s * b, N=o * r, K=l * i, a_m=(l * i * b, 1, b), a_k=b, b_k=o * r, b_n=1,
                 c_m=(o * b * r, r, b), c_n=(b * r, 1, r))
    return M


def _carry_from(Ut_or_Q, L, kept, b, r, transposed_basis):
    """T_next[kept, r, b] = basis^T L with the columns of L (b major) re-ordered to (r, b) on the fly."""
    rows = L.shape[0]
    Tn = empty(kept, r, b)
    if transposed_basis:     # basis given as (kept, rows) rows
        ops.gemm(Ut_or_Q, L, Tn, M=kept, N=b * r, K=rows, a_m=Ut_or_Q.stride(0), a_k=Ut_or_Q.stride(1), b_k=L.stride(0), b_n=L.stride(1),
                 c_m=r * b, c_n=(1, b, r))
    else:                    # basis given as (rows, kept) columns
        ops.gemm(Ut_or_Q, L, Tn, M=kept, N=b * r, K=rows, a_m=Ut_or_Q.stride(1), a_k=Ut_or_Q.stride(0), b_k=L.stride(0), b_n=L.stride(1),
                 c_m=r * b, c_n=(1, b, r))
    return Tn


def apply_round_qr(X, W, dim):
    """`W @ X` followed by `>> dim` with the reference's semantics (MPO:181-192 + MPS:432-468), fused: identical numbers
    to site_mpo_mps + round_qr (same projections, same column order), without the D = chi*chi_W product cores.
    One call of syn_apply_round_chain_f64 (csrc/chain.cu holds the sweep); apply_round_qr_steps is the same sweep step by step."""
    if _real_cuda_chain(X, W):
        return ops.apply_round_chain(X, W, int(dim))
    return apply_round_qr_steps(X, W, dim)


def apply_round_qr_steps(X, W, dim):
    """The fused reference-semantic sweep driven site by site from the host (what csrc/chain.cu does in one call)."""
    n = len(X)
    out = []
    T = torch.ones((1, 1, 1), dtype=F64, device=X[0].device)
    for k in range(n - 1):
        M = contract_carry(T, X[k], W[k])
        s, o, _ = M.shape
        b, r = X[k].shape[2], W[k].shape[3]
        L = M.reshape(s * o, b * r)
        Q, _ = qrt_step(L, dim, want_S=False)
        kept = Q.shape[1]
        out.append(Q.reshape(s, o, kept))
        T = _carry_from(Q, L, kept, b, r, transposed_basis=False)
    M = contract_carry(T, X[-1], W[-1])
    out.append(M)
    return out


# ---------------------------------------------------------------------------------------------------------
# SVD rounding (north_star; absent from the reference -- oracle/svd_numpy.py)
# ---------------------------------------------------------------------------------------------------------
class LazySpectrum:
    """Singular values of a bond whose kept subspace came from the spectral-projection solver (no eigenvalues are formed there):
    computed on demand as sqrt(eig(U^T A U)) with the Jacobi kernel -- only the kept values exist."""

    def __init__(self, A, U):
        self.A, self.U, self._sigma = A, U, None

    def get(self):
        if self._sigma is None:
            T = ops.matmul(self.U.t(), ops.matmul(self.A, self.U))
            T = ops.copy_strided(T)
            _, sigma, _, _ = ops.jacobi_solve(T, T.shape[0], 0.0, rank_tol=0.0, sqrt_mode=1, null_rel=0.0)
            self._sigma, self.A, self.U = sigma, None, None
        return self._sigma


class Truncation:
    """Per-bond record of an SVD rounding sweep (device tensors; .host() synchronises)."""

    def __init__(self):
        self.sigma, self.keep, self.discarded = [], [], []

    def host(self):
        sig = [s.get() if isinstance(s, LazySpectrum) else s for s in self.sigma]
        # None: the bond was cut by the spectral-projection solver on a complex unfolding (no spectrum is formed there)
        return ([np.zeros(0) if s is None else s.detach().cpu().numpy() for s in sig], list(self.keep), [float(d) for d in self.discarded])


def _svd_basis(M, chi_max, cutoff, trunc):
    """Left singular basis of the unfolding M (m x c): returns (core2d (m x keep) contiguous, keep)."""
    m, c = M.shape
    if isinstance(M, Cx):
        # complex unfolding: eigen-decomposition of the embedded Hermitian Gram matrix (cplx.svd_basis); a tall unfolding is
        # first reduced to its square triangular factor by the complex qrt
        if m > c:
            Q1 = cplx.polar_basis(M)                        # Newton-Schulz polar factor when it applies (large, full column rank)
            if Q1 is None:
                Q1, R1 = cplx.qrt(M, c)
            else:
                R1 = cplx.matmul(Q1.h(), M)
            U, keep, sigma, disc = cplx.svd_basis(R1, chi_max, cutoff, eigh_gram)
            core2d = cplx.matmul(Q1, U)
        else:
            core2d, keep, sigma, disc = cplx.svd_basis(M, chi_max, cutoff, eigh_gram)
        trunc.sigma.append(sigma)
        trunc.keep.append(keep)
        trunc.discarded.append(float(disc.item()))
        return core2d, keep
    if m <= c:
        G = ops.qr_r(M.t())                                 # R factor of M^T (m x m); rows rotate to sigma_i u_i^T
        Ut, sigma, info, winfo = ops.jacobi_solve(G, chi_max, cutoff, rank_tol=1e-14)
        keep = int(info[0].item())
        core2d = ops.copy_strided(Ut[:keep].t())
    else:
        Gt = empty(c, c)
        Q1, _ = ops.qrt(M, c, S=Gt.t())                     # Gt = R1^T ; rows rotate to sigma_i u1_i^T
        Ut, sigma, info, winfo = ops.jacobi_solve(Gt, chi_max, cutoff, rank_tol=1e-14)
        keep = int(info[0].item())
        core2d = ops.matmul(Q1, Ut[:keep].t())
    trunc.sigma.append(sigma)
    trunc.keep.append(keep)
    trunc.discarded.append(winfo[0].item())
    return core2d, keep


@complex_aware
def round_svd(sites, chi_max, cutoff=0.0, canonicalize=True):
    """Textbook rounding: right-to-left QR canonicalisation, then left-to-right truncated SVD with S V^T absorbed into the
    next core (oracle/svd_numpy.round_svd).  The SVD is Householder reduction + one-sided Jacobi on the device; the cutoff and
    chi_max selection happen in the finalize kernel, the host only reads the kept rank back."""
    out = right_orthonormalize(sites) if canonicalize else list(sites)
    trunc = Truncation()
    for k in range(len(out) - 1):
        cur, nxt = out[k], out[k + 1]
        M = cur.reshape(-1, cur.shape[-1])
        core2d, keep = _svd_basis(M, chi_max, cutoff, trunc)
        O = _O(M)
        carry = O.matmul(core2d.h() if isinstance(M, Cx) else core2d.t(), M)       # U^H M = S V^H  (keep x c)
        Wn = O.matmul(carry, nxt.reshape(nxt.shape[0], -1))
        out[k] = core2d.reshape(tuple(cur.shape[:-1]) + (keep,))
        out[k + 1] = Wn.reshape((keep,) + tuple(nxt.shape[1:]))
    return out, trunc


def right_environments(X, W):
    """E[k] = Gram matrix of the product chain to the right of bond k (D_k x D_k), k = 1..n-1, without forming product cores:
    two large strided GEMMs around the fused W-sandwich kernel per site (finished form of MPO:193-260).
    Storage: rows indexed (b, r) MPS-bond major, COLUMNS indexed (r', b') MPO-bond major -- so that the last GEMM of every
    site writes unit-stride rows (its batch index l' becomes the outer column index) and every consumer reads unit strides."""
    n = len(X)
    E = [None] * (n + 1)
    E[n] = torch.ones((1, 1), dtype=F64, device=X[0].device)
    for k in range(n - 1, 0, -1):
        Xk, Wk, En = X[k], W[k], E[k + 1]
        a, i, b = Xk.shape
        l, _, o, r = Wk.shape
        D = b * r
        # P1[(a,i),(r,y)] = sum_b X[(a,i),b] E[b,(r,y)]                       y = (r', b')
        P1 = empty(a, i, r, D)
        ops.gemm(Xk, En, P1, M=a * i, N=r * D, K=b, a_m=b, a_k=1, b_k=r * D, b_n=1, c_m=r * D, c_n=1)
        if ops.env_sandwich_fits(l, i, o, r, b) and Wk.is_contiguous():
            # both W contractions in one kernel, the (a, l, o, D) intermediate never reaches HBM (csrc/env.cu)
            Z = empty(a * l, l, i, b)
            ops.env_sandwich(P1, Wk, Z, a, b)
        else:
            Z = _w_sandwich_gemms(P1, Wk, a, i, b, l, o, r, D)
        # E[(a,l),(l',a')] = sum_{(i',b')} Z[(a,l), l', (i',b')] X[a',(i',b')]        batch over l'
        Ek = empty(a * l, l * a)
        ab = ENV_SYMMETRIC_BLOCK
        if ab and a % ab == 0 and a >= 2 * ab and l * l <= 65535 and (ab * l) % 128 == 0 and ab % 64 == 0:      # env_mirror launches gridDim.z = l * l
            # E is symmetric under (a,l) <-> (a',l'): form only the a-blocks on and below the diagonal -- ONE launch with a block-lower
            # output mask (rows in steps of ab * l, columns in steps of ab: 62.5 % of the flops at four blocks) -- then mirror the rest
            ops.gemm(Z, Xk, Ek, M=a * l, N=a, K=i * b, a_m=l * i * b, a_k=1, b_k=1, b_n=i * b, c_m=a * l, c_n=1,
                     batch=l, a_b=i * b, b_b=0, c_b=a, mask=(ab * l, ab))
            ops.env_mirror(Ek, a, l, ab)
        else:
            ops.gemm(Z, Xk, Ek, M=a * l, N=a, K=i * b, a_m=l * i * b, a_k=1, b_k=1, b_n=i * b, c_m=a * l, c_n=1,
                     batch=l, a_b=i * b, b_b=0, c_b=a)
        E[k] = Ek
    return E


ENV_SYMMETRIC_BLOCK = 64       # a-block of the block-lower environment build (0 = full GEMM): one masked launch of 128 x 64 tiles


def _w_sandwich_gemms(P1, Wk, a, i, b, l, o, r, D):
    """The two W contractions of the environment update as two strided GEMMs (shapes the fused kernel does not cover)."""
    # P2[a,(l,o),y] = sum_{(i,r)} W[l,i,o,r] P1[a,(i,r),y]                batch over a
    P2 = empty(a, l, o, D)
    ops.gemm(Wk, P1, P2, M=l * o, N=D, K=i * r, a_m=(i * o * r, r, o), a_k=(o * r, 1, r), b_k=D, b_n=1, c_m=D, c_n=1,
             batch=a, a_b=0, b_b=i * r * D, c_b=l * o * D)
    # Z[(a,l), l', i', b'] = sum_{(o,r')} P2[(a,l), o, (r',b')] W[l', i', o, r']  batch over (a,l)
    Z = empty(a * l, l, i, b)
    ops.gemm(P2, Wk, Z, M=b, N=l * i, K=o * r, a_m=1, a_k=(D, b, r), b_k=1, b_n=(i * o * r, o * r, i), c_m=1, c_n=b,
             batch=a * l, a_b=o * D, b_b=0, c_b=l * i * b)
    return Z




def gram_with_environment(M2, E, b, r):
    """A = M2 E M2^T for an unfolding M2 (rows x (b,r), MPS-bond major) and an environment stored as in right_environments
    (rows (b,r), columns (r',b')).  M2 E is written with its columns permuted back to (b',r') on the fly (two-level C index),
    and the final rows x rows product, whose K = D is long and whose output is small, is split 8 ways along K so that it
    fills the machine; the partial sums are added by a small reduction kernel."""
    rows, D = M2.shape
    ME = empty(rows, D)
    ops.gemm(M2, E, ME, M=rows, N=D, K=D, a_m=M2.stride(0), a_k=1, b_k=E.stride(0), b_n=1, c_m=D, c_n=(1, r, b))
    split = 8 if (D % (8 * 32) == 0 and D >= 2048) else 1
    if split == 1:
        return ops.matmul(ME, M2.t())
    kc = D // split
    part = empty(split, rows, rows)
    ops.gemm(ME, M2, part, M=rows, N=rows, K=kc, a_m=D, a_k=1, b_k=1, b_n=M2.stride(0), c_m=rows, c_n=1,
             batch=split, a_b=kc, b_b=kc, c_b=rows * rows)
    return ops.sum_parts(part)


CHOLESKY_MIN_N = 129          # multi-CTA Jacobi problems run on the shifted Cholesky factor of the Gram matrix (see eigh_gram); 0 = off


def eigh_gram(A, chi_max, cutoff, rank_tol):
    """Eigen-decomposition of a symmetric PSD matrix A (n x n, overwritten) for the density-matrix rounding:
    returns (Ut, sigma, info, winfo) like ops.jacobi_solve(sqrt_mode=True) -- rows of Ut are the eigenvectors, sorted; `info` is on
    the host and the solve has been checked for convergence (a non-converged Jacobi is resumed, never truncated on).

    (An FP32-preconditioned variant -- FP32 Jacobi for an approximate basis, two Newton-Schulz steps, three FP64 sweeps -- was measured
    in round 1 at 7.7 ms against 8.2 ms for the plain sweeps at n = 512 and removed in round 2: not worth its moving parts.)"""
    n = A.shape[0]
    null_rel = rank_tol * rank_tol
    if CHOLESKY_MIN_N and n >= CHOLESKY_MIN_N and A.is_contiguous():
        # Jacobi on the rows of B = L^T, G + delta I = L L^T: works on a matrix similar to G instead of G^2 -- 10 sweeps instead of
        # 13 on the C2 plateau, 15 instead of 27 on its rank-deficient sites (tools/chol_experiment.py); sigma comes out directly.
        B, shift = ops.chol_upper(A)
        return ops.jacobi_solve(B, chi_max, cutoff, rank_tol=rank_tol, sqrt_mode=2, shift=shift, null_rel=0.0)
    return ops.jacobi_solve(A, chi_max, cutoff, rank_tol=rank_tol, sqrt_mode=True, null_rel=null_rel)


# Bonds whose Gram matrix is at least this large take the spectral-projection solver (csrc/purify.cu) instead of Cholesky + Jacobi:
# GEMM-bound (~1 ms at n = 512 against 6.3 ms), same kept subspace.  0 = off.  Needs a spectral gap at the cut, no cutoff, and
# chi_max < n; otherwise (or when its own checks fail) the Jacobi path runs.
PURIFY_MIN_N = 128
PURIFY_SP2_ITERS = 52          # multi-launch variant: fixed counts
PURIFY_NS_ITERS = 26
PURIFY_SP2_MAX = 90            # fused kernel: upper limits (it stops by itself); 90 steps resolve gaps down to ~1e-12 |A|
PURIFY_NS_MAX = 60
# Accuracy guard.  The projector is accurate normwise: the kept space tilts by ~3e-15 |A| / gap, i.e. the state moves by
# ~3e-15 (lambda_0 / gap) sqrt(lambda_cut / lambda_0) (measured on a chain with a steeply decaying spectrum, tools/purify_diag.py),
# whereas Jacobi on the Cholesky factor is accurate relative to each singular value.  The kernel reports lift = number of leading
# 2X - X^2 steps ~ log2(|A|_F / lambda_cut); with gaps ~ lambda_cut / 100 (C2 plateau: lift 6-10) the 1e-10 parity bound needs
# lift <= ~14; when the cut is the edge of the null space (target = structural rank, gap = lambda_cut) lift <= 24 is enough.
PURIFY_MAX_LIFT = 14
PURIFY_MAX_LIFT_RANK_GAP = 24
PURIFY_STATS = {"taken": 0, "fallback": 0}
IDENTITY_WHEN_FULL = True      # bonds that keep their whole space skip the eigen-solve (gauge: identity core)


def _projection_verdict(h, ne, rank_gap):
    """Accept / reject the result of ops.dominant_subspace from its 8 info doubles (host copy): (ok, discarded weight)."""
    tr, f2, kept_w, dev, tr_a, idem = h[0], h[1], h[2], h[4], h[5], h[6]
    ok = (abs(tr - ne) < 1e-9 * ne and abs(f2 - ne) < 1e-9 * ne and abs(idem) < 1e-11 * ne and dev < 1e-12
          and np.isfinite(h).all() and int(h[7]) // 1000000 <= (PURIFY_MAX_LIFT_RANK_GAP if rank_gap else PURIFY_MAX_LIFT))
    return bool(ok), max(float(tr_a - kept_w), 0.0)


def dominant_subspace(A, chi_max, rank_gap=False):
    """(U (n x chi_max) orthonormal, discarded weight) by spectral projection, or None when the iteration did not reach a projector
    of trace chi_max that is orthonormalised to 1e-12 (no gap at the cut: rank-deficient bonds) or the cut lies too deep in the
    spectrum for its normwise accuracy (see PURIFY_MAX_LIFT) -- one host read of 8 doubles."""
    U, info = ops.dominant_subspace(A, chi_max, PURIFY_SP2_ITERS, PURIFY_NS_ITERS, sp2_max=PURIFY_SP2_MAX, ns_max=PURIFY_NS_MAX)
    ok, disc = _projection_verdict(info.cpu().numpy(), chi_max, rank_gap)
    PURIFY_STATS["taken" if ok else "fallback"] += 1
    return (U, disc) if ok else None


class _AsyncVerdict:
    """The 8 info doubles of a projection solve, copied to pinned host memory on a SIDE stream that waits only for the solver:
    reading it does not wait for whatever was queued on the main stream afterwards (a `.cpu()` on the main stream would)."""
    _side, _ring, _next = {}, {}, {}

    def __init__(self, info):
        dev = info.device
        key = dev.index
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device=dev)
            self._ring[key] = [torch.empty(8, dtype=F64).pin_memory() for _ in range(4)]
            self._next[key] = 0
        side = self._side[key]
        self.host = self._ring[key][self._next[key] % 4]
        self._next[key] += 1
        ready = torch.cuda.Event()
        ready.record()                                   # main stream: right after the solver
        side.wait_event(ready)
        info.record_stream(side)
        with torch.cuda.stream(side):
            self.host.copy_(info, non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(side)

    def get(self):
        self.done.synchronize()
        return self.host.numpy().copy()


class HostStreamer:
    """Result cores leave the device while the sweep is still running: a core is copied to its pinned host buffer on a side stream as soon
    as it is final (its projection verdict has been accepted), the caller's stream waits for the last copy at the end.  `bufs`: one pinned
    1-D float64 host tensor per site, at least as long as the core."""

    def __init__(self, bufs):
        self.bufs, self.copied = list(bufs), 0
        self.stream = torch.cuda.Stream(device())

    def push(self, out, upto):
        if upto <= self.copied:
            return
        ev = torch.cuda.Event()
        ev.record()                                   # everything that produces out[:upto] is queued before this point
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            for idx in range(self.copied, upto):
                c = out[idx]
                self.bufs[idx][: c.numel()].copy_(c.reshape(-1), non_blocking=True)
        self.copied = upto

    def finish(self, out):
        self.push(out, len(out))
        torch.cuda.current_stream().wait_stream(self.stream)


def copy_to_host(sites, bufs):
    """Plain (un-overlapped) form of HostStreamer for the routes that do not stream."""
    h = HostStreamer(bufs)
    h.finish(list(sites))


def apply_round_dm(X, W, chi_max, cutoff=0.0, rank_tol=3.2e-7, capture=None, host_out=None):
    """`W @ X` + optimal (SVD) rounding by the density-matrix algorithm: right environments, then a left-to-right sweep that finds the
    dominant eigenspace of M E M^T at every bond -- mathematically the truncation of round_svd(site_mpo_mps(...)) but every matrix
    stays <= (chi d) x D.  Eigenvalues are squared singular values, so values below ~1e-8 sigma_0 are noise: rank_tol.

    The projection solver's verdict (8 doubles) is read ONE SITE LATE: its basis is used at once, the next site's kernels are queued, and
    only then does the host look at the verdict, which travels to pinned host memory on a side stream (_AsyncVerdict) -- the host
    waits for the solver of site k only, with site k+1 already queued, so the GPU never waits for the host.  A rejected bond (no gap at the cut, or the
    accuracy guard) rolls the sweep back to that site, which is then solved by Cholesky + Jacobi.

    A relative cutoff BELOW rank_tol asks for singular values the Gram matrix cannot resolve (its eigenvalues carry ~1e-16 sigma_0^2 of
    noise): such calls take the textbook route -- product cores, right-QR, left-to-right Jacobi SVD of the triangular factors
    (round_svd), which is accurate relative to every singular value -- so the result follows the cutoff exactly as the oracle's does.

    `capture` (tests only): a dict {site: None}; the Gram matrix of that bond and the basis the solver kept are stored in it."""
    if 0.0 < cutoff < rank_tol:
        res = round_svd([site_mpo_mps(x, w) for x, w in zip(X, W)], chi_max, cutoff)
        if host_out is not None:
            copy_to_host(res[0], host_out)
        return res
    streamer = HostStreamer(host_out) if host_out is not None else None
    n = len(X)
    E = right_environments(X, W)
    trunc = Truncation()
    out = []
    T = torch.ones((1, 1, 1), dtype=F64, device=X[0].device)
    right_dim = [1] * (n + 1)                 # dimension of the physical space to the right of bond k: bounds the rank of that bond
    for k in range(n - 1, -1, -1):
        right_dim[k] = min(right_dim[k + 1] * int(W[k].shape[2]), 1 << 40)

    def jacobi_site(A, M2, s, o, b, r):
        nA = s * o
        if nA > 1024:
            raise NotImplementedError("density-matrix rounding needs chi*d <= 1024 (got %d)" % nA)
        Ut, sigma, info, winfo = eigh_gram(A, chi_max, cutoff, rank_tol)
        keep = int(info[0].item())
        trunc.sigma.append(sigma); trunc.keep.append(keep); trunc.discarded.append(winfo[0].item())
        out.append(ops.copy_strided(Ut[:keep].t()).reshape(s, o, keep))
        return _carry_from(Ut[:keep], M2, keep, b, r, transposed_basis=True)

    k = 0
    pending = None                # (site, info on the device, target rank, rank_gap, carry before the site, slot in out / trunc)
    while k < n - 1 or pending is not None:
        queued = None
        if k < n - 1:
            M = contract_carry(T, X[k], W[k])
            s, o, D = M.shape
            b, r = X[k].shape[2], W[k].shape[3]
            M2 = M.reshape(s * o, D)
            A = gram_with_environment(M2, E[k + 1], b, r)
            nA = s * o
            # structural rank of the bond: when it is below chi_max nothing is truncated and the kept space is the range of A -- the
            # projection solver then finds it at the (huge) gap between the last non-zero eigenvalue and the null space
            ne = min(chi_max, D, right_dim[k + 1])
            if IDENTITY_WHEN_FULL and cutoff == 0.0 and ne >= nA:
                # nothing to truncate and (structurally) nothing rank-deficient: every orthonormal basis of the whole space is a valid
                # gauge, so the core is the identity and the unfolding itself is carried -- no eigen-solve (ramp-up sites of a chain)
                eye = torch.eye(nA, dtype=F64, device=A.device)
                trunc.sigma.append(LazySpectrum(A, eye)); trunc.keep.append(nA); trunc.discarded.append(0.0)
                out.append(eye.reshape(s, o, nA))
                T = _carry_from(eye, M2, nA, b, r, transposed_basis=False)    # = M2 with its columns re-ordered to (r, b)
            elif PURIFY_MIN_N and nA >= PURIFY_MIN_N and cutoff > 0.0 and ne < nA and A.is_contiguous():
                # relative cutoff on the fast path: the projection solver finds the chi_max-dimensional dominant space, the kept
                # singular values are the spectrum of the SMALL matrix U^T A U (Jacobi on ne x ne instead of nA x nA), and the cutoff
                # picks a prefix of them; the verdict is read at once (the kept rank is data dependent)
                U, info = ops.dominant_subspace(A, ne, PURIFY_SP2_ITERS, PURIFY_NS_ITERS, sp2_max=PURIFY_SP2_MAX, ns_max=PURIFY_NS_MAX)
                ok, disc = _projection_verdict(info.cpu().numpy(), ne, ne >= min(D, right_dim[k + 1]))
                PURIFY_STATS["taken" if ok else "fallback"] += 1
                if ok:
                    Tm = ops.copy_strided(ops.matmul(U.t(), ops.matmul(A, U)))
                    Vt, sigma, hinfo, winfo = ops.jacobi_solve(Tm, ne, cutoff, rank_tol=rank_tol, sqrt_mode=1, null_rel=0.0)
                    keep = int(hinfo[0])
                    if keep < ne:
                        U = ops.matmul(U, Vt[:keep].t())           # the kept eigenvectors of U^T A U, back in the full space
                    trunc.sigma.append(sigma); trunc.keep.append(keep); trunc.discarded.append(disc + float(winfo[0].item()))
                    out.append(U.reshape(s, o, keep))
                    T = _carry_from(U, M2, keep, b, r, transposed_basis=False)
                else:
                    T = jacobi_site(A, M2, s, o, b, r)
            elif PURIFY_MIN_N and nA >= PURIFY_MIN_N and ne < nA and A.is_contiguous():
                U, info = ops.dominant_subspace(A, ne, PURIFY_SP2_ITERS, PURIFY_NS_ITERS, sp2_max=PURIFY_SP2_MAX, ns_max=PURIFY_NS_MAX)
                queued = (k, _AsyncVerdict(info) if info.is_cuda else info, ne, ne >= min(D, right_dim[k + 1]), T, len(out))
                if capture is not None and k in capture:
                    capture[k] = (A.clone(), U.clone())
                trunc.sigma.append(LazySpectrum(A, U)); trunc.keep.append(ne); trunc.discarded.append(None)
                out.append(U.reshape(s, o, ne))
                T = _carry_from(U, M2, ne, b, r, transposed_basis=False)
            else:
                T = jacobi_site(A, M2, s, o, b, r)
            k += 1
        if pending is not None:
            pk, pinfo, pne, pgap, pT, slot = pending
            ok, disc = _projection_verdict(pinfo.get() if isinstance(pinfo, _AsyncVerdict) else pinfo.cpu().numpy(), pne, pgap)
            PURIFY_STATS["taken" if ok else "fallback"] += 1
            if ok:
                trunc.discarded[slot] = disc
            else:
                # roll back to site pk: drop its speculative result and whatever was queued after it, solve it with Jacobi
                del out[slot:], trunc.sigma[slot:], trunc.keep[slot:], trunc.discarded[slot:]
                M = contract_carry(pT, X[pk], W[pk])
                s, o, D = M.shape
                b, r = X[pk].shape[2], W[pk].shape[3]
                M2 = M.reshape(s * o, D)
                T = jacobi_site(gram_with_environment(M2, E[pk + 1], b, r), M2, s, o, b, r)
                k = pk + 1
                queued = None
        pending = queued
        if streamer is not None:                       # cores before the one still awaiting its verdict are final
            streamer.push(out, queued[5] if queued is not None else len(out))
    out.append(contract_carry(T, X[-1], W[-1]))
    if streamer is not None:
        streamer.finish(out)
    return out, trunc
