"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's matrix-product hot path.

This is the parity ORACLE for the CUDA path.  It restates, on the CPU, what
antoine311200/Syngular's `syngular.tensor` does (citations `file:line` are relative to the
reference root), with exactly ONE generalisation: the *natural clamp* -- the width kept by
the QR-truncation step is min(q, rows(L)) (= Q.shape[1]); the reference crashes in `reshape`
when that is < q.  Wherever the reference runs, this agrees with it:
pinned by tests/golden/*.npz, which oracle/gen_golden.py produced from the unmodified reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file.

Conventions (tensor/matrix_product_state.py:54-56, tensor/matrix_product_operator.py:49-60):
  MPS core k : (l_k, d_k, r_k)            C-contiguous, l_0 = r_{N-1} = 1
  MPO core k : (l_k, in_k, out_k, r_k)
  dense MPO tensor is given as (in_0..in_{N-1}, out_0..out_{N-1}) and interleaved internally.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# elementary steps
# --------------------------------------------------------------------------------------
def qrt(L, q):
    """The reference truncation step (matrix_product_state.py:443-446, matrix_product_operator.py:555-558,
    and decompose :307-311 / :436-440): complete QR, keep the first q columns of Q and rows of R.
    Equivalent to projecting L on the span of its own first q columns.  Natural clamp: the kept
    width is min(q, L.shape[0])."""
    Q, R = np.linalg.qr(L, mode="complete")
    return Q[:, :q], R[:q, :]


def site_mpo_mps(X, W):
    """K1, matrix_product_operator.py:184-190:  C[(a,l), o, (b,r)] = sum_i X[a,i,b] W[l,i,o,r]."""
    a, i, b = X.shape
    l, i2, o, r = W.shape
    assert i == i2
    C = np.einsum("aib,lior->alobr", X, W)
    return C.reshape(a * l, o, b * r)


def site_mpo_mpo(A, B):
    """K2, matrix_product_operator.py:280-287: C[(lA,lB), iA, oB, (rA,rB)] = sum_x A[lA,iA,x,rA] B[lB,x,oB,rB]."""
    la, ia, xa, ra = A.shape
    lb, xb, ob, rb = B.shape
    assert xa == xb
    C = np.einsum("aixb,cxod->aciobd", A, B)
    return C.reshape(la * lb, ia, ob, ra * rb)


def site_add(A, B, first, last):
    """K9, matrix_product_state.py:82-96 / matrix_product_operator.py:90-106: direct sum of the bond spaces
    (self's block first); first core concatenates on r, last core on l.  Always float64 (np.zeros)."""
    la, ra = A.shape[0], A.shape[-1]
    lb, rb = B.shape[0], B.shape[-1]
    phys = A.shape[1:-1]
    out = np.zeros(((la if first else la + lb),) + phys + ((ra if last else ra + rb),))
    if first and last:
        # single-site chain: the reference's `wing_left` branch wins (np.block([A, B]) cannot fit) -> crash there;
        # here: plain sum is the only meaningful reading, but we keep the reference's behaviour out of scope.
        raise Exception("single-site addition is not defined by the reference")
    if first:
        out[..., :ra] = A
        out[..., ra:] = B
    elif last:
        out[:la] = A
        out[la:] = B
    else:
        out[:la, ..., :ra] = A
        out[la:, ..., ra:] = B
    return out


def site_kron(A, B):
    """K10, matrix_product_operator.py:140-152: per (in,out) Kronecker product of the bond matrices."""
    la, i, o, ra = A.shape
    lb, _, _, rb = B.shape
    C = np.einsum("aiob,ciod->aciobd", A, B)
    return C.reshape(la * lb, i, o, ra * rb)


def _flat3(core):
    """View a core as (l, phys, r) with all physical legs flattened."""
    return core.reshape(core.shape[0], -1, core.shape[-1])


# --------------------------------------------------------------------------------------
# sweeps on bare core lists
# --------------------------------------------------------------------------------------
def round_qr(cores, q):
    """The strict `>>` sweep (matrix_product_state.py:432-468, matrix_product_operator.py:544-580) on a list of
    cores; returns new cores.  No canonicalisation before or after."""
    cores = [np.array(c, copy=True) for c in cores]
    n = len(cores)
    for k in range(n - 1):
        cur, nxt = cores[k], cores[k + 1]
        phys = cur.shape[1:-1]
        L = cur.reshape(-1, cur.shape[-1])
        Q, S = qrt(L, q)
        kept = Q.shape[1]
        W = S @ nxt.reshape(nxt.shape[0], -1)
        cores[k] = Q.reshape((cur.shape[0],) + phys + (kept,))
        cores[k + 1] = W.reshape((kept,) + nxt.shape[1:])
    return cores


def left_orthonormalize(cores):
    """matrix_product_state.py:554-566 / matrix_product_operator.py:673-694: reduced QR sweep, left to right."""
    cores = [np.array(c, copy=True) for c in cores]
    for k in range(len(cores) - 1):
        cur, nxt = cores[k], cores[k + 1]
        L = cur.reshape(-1, cur.shape[-1])
        U, B = np.linalg.qr(L)
        W = B @ nxt.reshape(nxt.shape[0], -1)
        cores[k] = U.reshape(cur.shape[:-1] + (U.shape[1],))
        cores[k + 1] = W.reshape((U.shape[1],) + nxt.shape[1:])
    return cores


def right_orthonormalize(cores):
    """matrix_product_state.py:568-580 / matrix_product_operator.py:696-719: QR of R^T, right to left."""
    cores = [np.array(c, copy=True) for c in cores]
    for k in range(len(cores) - 1, 0, -1):
        cur, prv = cores[k], cores[k - 1]
        R = cur.reshape(cur.shape[0], -1)
        V, U = np.linalg.qr(R.T)
        W = prv.reshape(-1, prv.shape[-1]) @ U.T
        cores[k] = V.T.reshape((V.shape[1],) + cur.shape[1:])
        cores[k - 1] = W.reshape(prv.shape[:-1] + (V.shape[1],))
    return cores


def overlap(A, B):
    """`A | B`, matrix_product_state.py:116-129: bilinear (NO conjugation) transfer-matrix contraction."""
    E = np.ones((1, 1), dtype=np.result_type(A[0].dtype, B[0].dtype))
    for a, b in zip(A, B):
        a3, b3 = _flat3(a), _flat3(b)
        T = np.tensordot(E, a3, axes=(0, 0))            # (lb, d, ra)
        E = np.tensordot(T, b3, axes=([0, 1], [0, 1]))  # (ra, rb)
    return E.reshape(1)[0]


def to_dense(cores):
    """Dense tensor of a chain: MPS -> (d_0..d_{N-1}); MPO -> (in_0..in_{N-1}, out_0..out_{N-1}).
    Same numbers as the reference's per-index `retrieve` loops (matrix_product_state.py:265-273, :546-551;
    matrix_product_operator.py:405-416, :663-670), evaluated by one left-to-right chain."""
    T = cores[0].reshape(-1, cores[0].shape[-1])
    dims = list(cores[0].shape[1:-1])
    for c in cores[1:]:
        T = T @ c.reshape(c.shape[0], -1)
        T = T.reshape(-1, c.shape[-1])
        dims += list(c.shape[1:-1])
    T = T.reshape(dims)
    if cores[0].ndim == 4:
        n = len(cores)
        T = T.transpose(list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))
    return T


def retrieve(cores, idx_in, idx_out=None):
    """One amplitude as the product of the sliced bond matrices (matrix_product_state.py:546-551,
    matrix_product_operator.py:663-670); returns a (1,1) array like the reference."""
    M = None
    for k, c in enumerate(cores):
        m = c[:, idx_in[k], :] if idx_out is None else c[:, idx_in[k], idx_out[k], :]
        M = m if M is None else M @ m
    return M


def mps_apply(cores, gate, index):
    """`MatrixProductState.apply(gate, index)` (matrix_product_state.py:487-534): contract a dense m-site gate
    (legs out_0..out_{m-1}, in_0..in_{m-1}; column-vector convention) into cores index..index+m-1, then re-split with the
    qrt step keeping the EXISTING bonds (the bond never grows; entanglement beyond it is projected away)."""
    cores = [np.array(c, copy=True) for c in cores]
    gate = np.asarray(gate)
    m = gate.ndim // 2
    T = cores[index]
    for k in range(index + 1, index + m):
        T = np.tensordot(T, cores[k], axes=(T.ndim - 1, 0))            # (l, in_0..in_j, r)
    l, r = T.shape[0], T.shape[-1]
    din = int(np.prod(T.shape[1:-1]))
    G = gate.reshape(int(np.prod(gate.shape[:m])), din)
    T = np.einsum("oi,lir->lor", G, T.reshape(l, din, r)).reshape((l,) + tuple(gate.shape[:m]) + (r,))
    for k in range(index, index + m - 1):
        lr, d = T.shape[0], T.shape[1]
        L = T.reshape(lr * d, -1)
        Q, R = qrt(L, cores[k].shape[2])
        kept = Q.shape[1]
        cores[k] = Q.reshape(lr, d, kept)
        T = R.reshape(kept, -1, cores[k + 1].shape[2]) if k + 1 == index + m - 1 else R.reshape((kept,) + T.shape[2:])
    cores[index + m - 1] = T.reshape(T.shape[0], -1, T.shape[-1])
    return cores


def mpo_apply(cores, op_cores, indices):
    """`MatrixProductOperator.apply(operator, indices)` for an MPO operator (matrix_product_operator.py:582-626), literal:
    contract the target cores `indices` with the operator cores (target OUT leg with operator IN leg, both bond chains summed,
    the operator's outer bonds must be 1: np.squeeze at :602), giving T[l, in_0..in_{m-1}, out'_0..out'_{m-1}, r]; then re-split
    left to right with the qrt step, keeping the target's existing right bonds.  The reference reshapes T as
    (l * in_k * out'_k, -1) although its legs are NOT interleaved (:609), so for m >= 2 the result is a fixed scramble of the
    product, not the product -- replicated as is (pinned by tests/golden/mpo_apply.npz).  Returns the new list of cores."""
    cores = [np.array(c, copy=True) for c in cores]
    m = len(indices)
    T = None
    for idx, jdx in enumerate(indices):
        c = np.einsum("lixr,mxon->lmiorn", cores[jdx], op_cores[idx])           # (l, l', in, out', r, r')
        if T is None:
            if c.shape[1] != 1:
                raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
            T = c[:, 0]                                                          # (l, in, out', r, r')
        else:
            T = np.tensordot(T, c, axes=([T.ndim - 2, T.ndim - 1], [0, 1]))     # (..., in, out', r, r')
    if T.shape[-1] != 1:
        raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
    T = T[..., 0]                                                                # (l, in_0, out_0, in_1, out_1, ..., r)
    perm = [0] + [1 + 2 * k for k in range(m)] + [2 + 2 * k for k in range(m)] + [T.ndim - 1]
    T = np.ascontiguousarray(T.transpose(perm))                                  # (l, in..., out'..., r) as contract() returns it
    for idx in range(m - 1):
        jdx = indices[idx]
        l, i = cores[jdx].shape[0], cores[jdx].shape[1]
        o = op_cores[idx].shape[2]
        L = T.reshape(l * i * o, -1)
        Q, R = qrt(L, cores[jdx].shape[3])
        cores[jdx] = Q.reshape(l, i, o, Q.shape[1])
        T = R
    jdx = indices[-1]
    cores[jdx] = T.reshape(cores[jdx].shape[0], cores[jdx].shape[1], op_cores[m - 1].shape[2], cores[jdx].shape[3])
    return cores


def dmrg_right_blocks(state, operator):
    """DMRG.__right_blocks (variational/dmrg.py:65-87): R_k[a, w, a'] = sum S_k[a,o,b] W_k[w,i,o,v] S_k[a',i,b'] R_{k+1}[b,v,b'] for
    k = n-1 .. 2 (no conjugation; the FIRST state copy meets the operator's OUT leg); entries 0 and 1 stay None."""
    n = len(state)
    blocks = [None] * n
    R = np.ones((1, 1, 1))
    for k in range(n - 1, 1, -1):
        R = np.einsum("aob,wiov,cid,bvd->awc", state[k], operator[k], state[k], R)
        blocks[k] = R
    return blocks


def dmrg_left_blocks(state, operator):
    """DMRG.__left_blocks (variational/dmrg.py:90-112): L_k[b, v, b'] = sum L_{k-1}[a,w,a'] S_k[a,o,b] W_k[w,i,o,v] S_k[a',i,b'] for
    k = 0 .. n-3; the last two entries stay None."""
    n = len(state)
    blocks = [None] * n
    L = np.ones((1, 1, 1))
    for k in range(n - 2):
        L = np.einsum("awc,aob,wiov,cid->bvd", L, state[k], operator[k], state[k])
        blocks[k] = L
    return blocks


def dmpo_project(state, operator, index):
    """DifferentialMatrixProductOperator.project (tensor/differential_matrix_product_operator.py:79-173): the merged ("crumbled") pair
    of operator cores index, index+1 (:13-27), and the wings -- everything left of `index` / right of `index + 1` of W @ X contracted
    into (1, prod(out), a, w) / (a, w, prod(out), 1) tensors (:104-149; MPS bond before MPO bond, output legs flattened in site order).
    The reference's wing einsums have no operand when a wing is empty, so 1 <= index <= n - 3."""
    n = len(state)
    if not (0 <= index < n - 1):
        raise Exception("trying to project on non-existant site (site indices should be between 0 and the number of sites - 1)")
    if index < 1 or index > n - 3:
        raise Exception("project needs at least one site on each side of the merged pair (1 <= index <= n - 3)")
    wl, wr = operator[index], operator[index + 1]
    center = np.tensordot(wl, wr, axes=(3, 0))                                  # (l, i, o, i', o', r')
    left = None                                                                 # (prod(out), a, w)
    for k in range(index):
        C = np.einsum("aib,wiov->awobv", state[k], operator[k])               # MPS bond major, MPO bond minor on both sides
        a, w, o, b, v = C.shape
        if left is None:
            left = C.reshape(a * w * o, b, v)                                   # a = w = 1 at the chain's left end (np.squeeze, :122)
        else:
            left = np.tensordot(left, C, axes=([1, 2], [0, 1])).reshape(-1, b, v)
    right = None                                                                # (a, w, prod(out))
    for k in range(n - 1, index + 1, -1):
        C = np.einsum("aib,wiov->awobv", state[k], operator[k])
        a, w, o, b, v = C.shape
        if right is None:
            right = C.reshape(a, w, o * b * v)                                  # b = v = 1 at the chain's right end (:159)
        else:
            right = np.tensordot(C, right, axes=([3, 4], [0, 1])).reshape(a, w, -1)
    return {"center_site": center, "left_wing": left[None], "left_center": state[index], "right_center": state[index + 1],
            "right_wing": right[..., None]}


def decompose_left(T, shapes):
    """TT decomposition by the qrt step, left to right (matrix_product_state.py:298-319,
    matrix_product_operator.py:430-450).  `T` is the (interleaved, for an MPO) dense tensor and `shapes` the
    declared core shapes; the rank kept at bond k is shapes[k][-1] (clamped)."""
    cores = []
    n = len(shapes)
    l = 1
    for k in range(n - 1):
        phys = int(np.prod(shapes[k][1:-1]))
        L = T.reshape(l * phys, -1)
        Q, R = qrt(L, shapes[k][-1])
        kept = Q.shape[1]
        cores.append(Q.reshape((l,) + tuple(shapes[k][1:-1]) + (kept,)))
        T, l = R, kept
    cores.append(T.reshape((l,) + tuple(shapes[n - 1][1:-1]) + (1,)))
    return cores


def decompose_right(T, shapes):
    """MPS-only right-to-left variant (matrix_product_state.py:324-347): QR of the transposed right unfolding."""
    n = len(shapes)
    cores = [None] * n
    r = 1
    for k in range(n - 1, 0, -1):
        phys = int(np.prod(shapes[k][1:-1]))
        L = T.reshape(-1, phys * r).T
        Q, R = qrt(L, shapes[k][0])
        kept = Q.shape[1]
        cores[k] = Q.T.reshape((kept,) + tuple(shapes[k][1:-1]) + (r,))
        T, r = R.T, kept
    cores[0] = T.reshape((1,) + tuple(shapes[0][1:-1]) + (r,))
    return cores


# --------------------------------------------------------------------------------------
# the two container types, with the reference's metadata rules
# --------------------------------------------------------------------------------------
class _Chain:
    """Shared behaviour of the two containers.  Metadata rules restated from the reference:
      * `from_sites` derives shape / input_shape / bond_shape from the cores (MPS :191-214, MPO :357-380);
      * the result of `>>` keeps the PRE-truncation `bond_shape` (stale) while `.shape` is updated
        (MPS :433-468; MPO :545-580) -- later `min_bond` computations and `>>` guards depend on it;
      * `>>` with q >= min(bond_shape) returns the operand itself (MPS :367-368; MPO :473-474).
    Not replicated: the MPO copy()/from_sites list aliasing that overwrites the source."""

    PHYS = 1

    def __init__(self):
        self.sites = []
        self.sites_number = 0
        self.shape = []
        self.bond_shape = ()
        self.input_shape = ()
        self.decomposed = False
        self.orthonormalized = None
        self.parameters_number = 0
        self.real_parameters_number = 0
        self.verbose = 0

    @classmethod
    def from_sites(cls, sites, orthogonality=None, real_parameters_number=None):
        mp = cls()
        mp.sites = list(sites)
        mp.sites_number = len(sites)
        mp.decomposed = True
        mp.orthonormalized = orthogonality
        mp.parameters_number = int(sum(int(np.prod(s.shape)) for s in sites))
        mp.real_parameters_number = real_parameters_number
        mp.shape = [tuple(s.shape) for s in sites]
        mp.input_shape = tuple(s.shape[1] for s in sites)
        if cls.PHYS == 2:
            mp.output_shape = tuple(s.shape[2] for s in sites)
        mp.bond_shape = tuple(s.shape[-1] for s in sites[:-1])
        return mp

    def copy(self):
        return type(self).from_sites(self.sites)

    # `>>`
    def __rshift__(self, dim):
        if not isinstance(dim, int):
            raise Exception("dimension should be an integer")
        return self.compress(dim, strict=True)

    def compress(self, dim, mode="left", strict=False):
        if dim >= min(self.bond_shape):
            return self
        if strict:
            that = self.copy()                       # bond_shape of `that` = actual bonds before truncation
            that.sites = round_qr(self.sites, dim)
            that.shape = [tuple(s.shape) for s in that.sites]
            that.parameters_number = int(sum(int(np.prod(s.shape)) for s in that.sites))
            return that
        # non-strict: canonicalise first, then the same qrt sweep from that side, in place
        # (MPS :370-430; MPO :480-541).  Returns None like the reference.
        if self.PHYS == 2:
            self.bond_shape = (dim,) * (self.sites_number - 1)      # MPO :477
        if mode == "left":
            if self.orthonormalized != "left":
                self.left_orthonormalization()
            self.sites = round_qr(self.sites, dim)
        elif mode == "right":
            if self.orthonormalized != "right":
                self.right_orthonormalization()
            rev = [np.swapaxes(_flat3(s), 0, 2) for s in reversed(self.sites)]
            # right sweep uses the *reduced* QR of R^T (MPS :410, MPO :520) -- same projection as qrt
            out = round_qr(rev, dim)
            shapes = [s.shape for s in self.sites]
            new = []
            for s, ref_shape in zip(reversed(out), shapes):
                t = np.swapaxes(s, 0, 2)
                new.append(t.reshape((t.shape[0],) + tuple(ref_shape[1:-1]) + (t.shape[2],)))
            self.sites = new
        self.shape = [tuple(s.shape) for s in self.sites]
        self.parameters_number = int(sum(int(np.prod(s.shape)) for s in self.sites))
        return None

    def left_orthonormalization(self):
        self.sites = left_orthonormalize(self.sites)
        self.shape = [tuple(s.shape) for s in self.sites]
        self.orthonormalized = "left"
        return self if self.PHYS == 2 else None

    def right_orthonormalization(self):
        self.sites = right_orthonormalize(self.sites)
        self.shape = [tuple(s.shape) for s in self.sites]
        self.orthonormalized = "right"
        return self if self.PHYS == 2 else None

    def left_orthogonality(self, k):
        L = self.sites[k].reshape(-1, self.sites[k].shape[-1])
        return L.T @ L

    def right_orthogonality(self, k):
        R = self.sites[k].reshape(self.sites[k].shape[0], -1)
        return R @ R.T

    def to_tensor(self):
        return to_dense(self.sites)


class MPS(_Chain):
    PHYS = 1

    @classmethod
    def dense(cls, tensor, bond_shape, mode="left"):
        """`MatrixProductState(tensor, bond_shape).decompose(mode)` (matrix_product_state.py:30-61, :296-350)."""
        tensor = np.asarray(tensor)
        n = len(bond_shape) + 1
        if tensor.ndim != n:
            raise Exception("dimensions of bond indices do not match order - 1")
        b = (1,) + tuple(bond_shape) + (1,)
        shapes = [(b[k], tensor.shape[k], b[k + 1]) for k in range(n)]
        cores = decompose_left(tensor, shapes) if mode == "left" else decompose_right(tensor, shapes)
        mp = cls.from_sites(cores, real_parameters_number=int(np.prod(tensor.shape)))
        mp.bond_shape = tuple(bond_shape)            # declared, not derived (matrix_product_state.py:38)
        mp.shape = shapes
        mp.input_shape = tuple(tensor.shape)
        return mp

    @classmethod
    def zeros(cls, input_shape, bond_shape):
        n = len(input_shape)
        b = (1,) + tuple(bond_shape) + (1,)
        return cls.from_sites([np.zeros((b[k], input_shape[k], b[k + 1])) for k in range(n)])

    def __add__(self, other):
        if not (self.decomposed and other.decomposed):
            raise Exception("Both Matrix Product Operator must be in canonical form (use .decompose()")
        n = self.sites_number
        return MPS.from_sites([site_add(self.sites[k], other.sites[k], k == 0, k == n - 1) for k in range(n)])

    def __or__(self, other):
        if not isinstance(other, MPS):
            raise Exception("right-hand site must be a MatrixProductState")
        return overlap(self.sites, other.sites)

    def dot(self):
        return np.sqrt(self | self)

    def normalize(self):
        """matrix_product_state.py:252-256: divide the LAST core by its Frobenius norm, in place."""
        self.sites[-1] = self.sites[-1] / np.linalg.norm(self.sites[-1].reshape(self.sites[-1].shape[0], -1))
        return self

    def __getitem__(self, key):
        if len(key) != self.sites_number:
            raise Exception("input indices do not match the number of sites")
        return retrieve(self.sites, key)


class MPO(_Chain):
    PHYS = 2

    @classmethod
    def dense(cls, tensor, bond_shape):
        """`MatrixProductOperator(tensor, bond_shape).decompose()` (matrix_product_operator.py:25-64, :418-467)."""
        tensor = np.asarray(tensor)
        bond_shape = tuple(bond_shape)
        n = len(bond_shape) + 1
        inp, out = tuple(tensor.shape[:n]), tuple(tensor.shape[n:])
        if len(inp) != len(out):
            raise Exception("input_shape and output_shape of the tensor must have the same length")
        if len(inp) != n:
            raise Exception("dimensions of bond indices do not match input dimension - 1")
        if bond_shape == ():
            mp = cls.from_sites([tensor.reshape(1, inp[0], out[0], 1)])
            return mp
        inter = tensor.transpose(sum(zip(range(n), range(n, 2 * n)), ()))
        b = (1,) + bond_shape + (1,)
        shapes = [(b[k], inp[k], out[k], b[k + 1]) for k in range(n)]
        cores = decompose_left(inter, shapes)
        mp = cls.from_sites(cores, real_parameters_number=int(np.prod(tensor.shape)))
        mp.bond_shape = bond_shape
        mp.shape = shapes
        return mp

    @classmethod
    def zeros(cls, input_shape, output_shape, bond_shape):
        n = len(input_shape)
        b = (1,) + tuple(bond_shape) + (1,)
        return cls.from_sites([np.zeros((b[k], input_shape[k], output_shape[k], b[k + 1])) for k in range(n)])

    def __add__(self, other):
        m = min(min(self.bond_shape), min(other.bond_shape))          # matrix_product_operator.py:80 (metadata!)
        if not (self.decomposed and other.decomposed):
            raise Exception("Both Matrix Product Operator must be in canonical form (use .decompose()")
        n = self.sites_number
        sites = [site_add(self.sites[k], other.sites[k], k == 0, k == n - 1) for k in range(n)]
        return MPO.from_sites(sites) >> m

    def __mul__(self, other):
        m = min(min(self.bond_shape), min(other.bond_shape))          # :128
        if not isinstance(other, MPO):
            raise Exception("left hand-side must be either a MatrixProductState or a MatrixProductOperator")
        sites = [site_kron(a, b) for a, b in zip(self.sites, other.sites)]
        return MPO.from_sites(sites) >> m

    def __matmul__(self, other):
        m = min(min(self.bond_shape), min(other.bond_shape))          # :176 (metadata!)
        if isinstance(other, MPS):
            sites = [site_mpo_mps(x, w) for x, w in zip(other.sites, self.sites)]
            return MPS.from_sites(sites) >> m                          # :192
        if isinstance(other, MPO):
            sites = [site_mpo_mpo(a, b) for a, b in zip(self.sites, other.sites)]
            return MPO.from_sites(sites) >> m                          # :289
        return None

    def apply(self, operator, indices):
        """In place, returns None (matrix_product_operator.py:582-626)."""
        self.sites = mpo_apply(self.sites, operator.sites, list(indices))
        return None

    def __getitem__(self, key):
        kin, kout = key
        if len(kin) != self.sites_number:
            raise Exception("input indices do not match the number of sites")
        if len(kout) != self.sites_number:
            raise Exception("output indices do not match the number of sites")
        return retrieve(self.sites, kin, kout)


def mul(op1, op2, mode="standard"):
    """`syn.mul`, tensor/utils.py:8-73.  NOTE (MPO,MPO) contracts op2's OUTPUT with op1's INPUT (:62) --
    the reverse of `@` -- with op2's bond major."""
    if op1.sites_number != op2.sites_number:
        raise Exception("both operator do not have the same number of sites")
    m = min(min(op1.bond_shape), min(op2.bond_shape))
    if mode != "standard":
        return None
    if isinstance(op1, MPS) and isinstance(op2, MPO):
        if not op1.decomposed or not op2.decomposed:
            raise Exception("Operators and States must be decomposed")
        return MPS.from_sites([site_mpo_mps(x, w) for x, w in zip(op1.sites, op2.sites)]) >> m
    if isinstance(op1, MPO) and isinstance(op2, MPS):
        if not op1.decomposed or not op2.decomposed:
            raise Exception("Operators and States must be decomposed")
        return MPS.from_sites([site_mpo_mps(x, w) for x, w in zip(op2.sites, op1.sites)]) >> m
    if isinstance(op1, MPS) and isinstance(op2, MPS):
        return op1 | op2
    if isinstance(op1, MPO) and isinstance(op2, MPO):
        return MPO.from_sites([site_mpo_mpo(b, a) for a, b in zip(op1.sites, op2.sites)]) >> m
    raise Exception("`syn.mul` should be provided MatrixProductState or MatrixProductOperator objects only")


# --------------------------------------------------------------------------------------
# quantum/ (caller of the path for BASELINE configs[2]) -- real gates only
# --------------------------------------------------------------------------------------
class Gates:
    """quantum/gate.py:4-61 (legs of multi-qubit gates: out..., in...)."""
    I = np.eye(2)
    X = np.array([[0., 1.], [1., 0.]])
    Z = np.array([[1., 0.], [0., -1.]])
    H = np.array([[1., 1.], [1., -1.]]) / np.sqrt(2)
    CX = np.array([[1., 0, 0, 0], [0, 1., 0, 0], [0, 0, 0, 1.], [0, 0, 1., 0]]).reshape(2, 2, 2, 2)
    SWAP = np.array([[1., 0, 0, 0], [0, 0, 1., 0], [0, 1., 0, 0], [0, 0, 0, 1.]]).reshape(2, 2, 2, 2)
    TOFFOLI = np.eye(8)[[0, 1, 2, 3, 4, 5, 7, 6]].reshape((2,) * 6)


class Qbit:
    """quantum/qbit.py:11-150: |0..0> register with every bond fixed at 2; gates through mps_apply."""

    def __init__(self, size, cores=None):
        self.size = size
        if cores is None:
            b = [1] + [2] * (size - 1) + [1]
            cores = []
            for k in range(size):
                c = np.zeros((b[k], 2, b[k + 1]))
                if k < size - 1:
                    c[0] = np.eye(2)[:, : b[k + 1]]
                else:
                    c[0, 0, 0] = 1.0
                cores.append(c)
        self.cores = cores

    def _apply(self, g, i):
        return Qbit(self.size, mps_apply(self.cores, g, i))

    def __matmul__(self, op):
        if len(op) == 2:
            return self._apply(*op)
        g, a, b = op
        lo, hi = min(a, b), max(a, b)
        q = self
        if hi - lo != 1:                                   # qbit.py:42-50
            for i in range(lo, hi - 1):
                q = q._apply(Gates.SWAP, i)
            if lo != a:
                q = q._apply(Gates.SWAP, hi - 1)
        q = q._apply(g, hi - 1)
        if hi - lo != 1:                                   # qbit.py:57-58, swap_out :91-106
            for i in range(hi - 1, lo - 1, -1):
                q = q._apply(Gates.SWAP, i)
        return q

    def to_tensor(self):
        return to_dense(self.cores).reshape(-1)
