"""Backend-agnostic restatement of the reference's own test scenarios against the committed goldens.

Every case takes a `backend` (tests/backends.py): the numpy oracle on the CPU, or the CUDA product
through the `syngular.tensor` drop-in API on the GPU.  The SAME assertions run for both, so the oracle is
pinned to the reference (goldens were produced by the unmodified reference, oracle/gen_golden.py) and the
CUDA path is pinned to both.  Only gauge-invariant quantities are compared (dense tensors, sampled
elements, overlaps, Gram matrices, shapes and the reference's -- deliberately stale -- bond metadata).
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-10          # north_star: gauge-invariant results within 1e-10 relative in FP64


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def chain(store, prefix):
    n = int(store[prefix + "/n"])
    return [store["%s/site%d" % (prefix, k)] for k in range(n)]


def meta(store, prefix):
    return tuple(int(v) for v in store[prefix + "/bond_shape"]), [tuple(int(x) for x in r) for r in store[prefix + "/shape"]]


def close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=np.float64 if not np.iscomplexobj(a) else np.complex128)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= rtol, "%s: max rel err %.3e > %.1e" % (what, err, rtol)


def same_meta(be, obj, store, prefix):
    bond, shape = meta(store, prefix)
    assert tuple(int(b) for b in obj.bond_shape) == bond, "%s bond_shape %s vs reference %s" % (prefix, obj.bond_shape, bond)
    assert [tuple(int(x) for x in s) for s in obj.shape] == shape, "%s shape %s vs reference %s" % (prefix, obj.shape, shape)
    assert [tuple(s.shape) for s in be.sites(obj)] == shape


# ------------------------------------------------------------------------------------------------
def case_matmul_known_answers(be):
    g = load("known_answers")
    x = np.arange(4).reshape(2, 2).astype(float)
    w = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    X = be.mps_dense(x, (2,))
    W = be.mpo_dense(w, (2,))
    Y = W @ X
    close(be.dense(Y).reshape(-1), [56, 62, 68, 74], what="W@X (test_syn.py:10-27)")
    close(be.dense(Y), g["matmul/WX_dense"], what="W@X golden")
    close(be.scalar(X | X), 14.0, what="X|X")
    Z = W @ W @ X
    # the reference's lossy QR-truncation is pinned by this vector (an SVD `>>` would fail it)
    close(be.dense(Z).reshape(-1), [1509.179143, 1742.53175013, 2253.42694889, 2522.46360351], rtol=1e-8,
          what="W@W@X (test_mpo.py:177-196)")
    close(be.dense(Z), g["matmul/WWX_dense"], what="W@W@X golden")
    same_meta(be, Z, g, "matmul/WWX")
    WW = W @ W
    close(be.dense(WW), g["matmul/WW_dense"], what="W@W dense")
    same_meta(be, WW, g, "matmul/WW")
    close(be.dense(be.mul(W, X)), g["syn/mul_WX"], what="syn.mul(W,X)")
    close(be.dense(be.mul(X, W)), g["syn/mul_XW"], what="syn.mul(X,W)")
    close(be.scalar(be.mul(X, X)), g["syn/mul_XX"], what="syn.mul(X,X)")
    close(be.dense(be.mul(W, W)), g["syn/mul_WW"], what="syn.mul(W,W)")


def case_mps_dot_compress_normalize(be):
    g = load("known_answers")
    x = np.arange(64).reshape(4, 4, 4).astype(float)
    X = be.mps_dense(x, (4, 4))
    close(be.dense(X), x, what="decompose reproduces x")
    close(be.scalar(X | X), 85344.0, what="X|X (test_mps.py:100-110)")
    close(be.scalar(X.dot()), 292.1369541841634, what="X.dot()")
    Z = X >> 2
    assert [tuple(s.shape) for s in be.sites(Z)] == [(1, 4, 2), (2, 4, 2), (2, 4, 1)]
    close(be.dense(Z), g["compress/Z_dense"], what="X>>2 dense")
    close(be.dense(Z), x, rtol=1e-12, what="X>>2 lossless (test_mps.py:72-85)")
    same_meta(be, Z, g, "compress/Z")
    Xn = be.mps_dense(x, (4, 4)).normalize()
    close(be.scalar(Xn.dot()), 1.0, what="normalize (test_mps.py:113-126)")
    for mode in ("left", "right"):
        Xm = be.mps_dense(x, (2, 2), mode=mode)
        close(be.dense(Xm), x, rtol=1e-12, what="decompose %s (test_mps.py:129-157)" % mode)
        same_meta(be, Xm, g, "decompose_%s/X" % mode)


def case_mps_add_augment(be):
    g = load("known_answers")
    x8 = np.arange(8).reshape(2, 2, 2).astype(float)
    X = be.mps_dense(x8, (2, 2))
    Y0 = be.mps_zeros((2, 2, 2), (2, 2))
    Z = X
    for _ in range(31):
        Z = Z + Y0
    assert tuple(Z.bond_shape) == (64, 64)
    same_meta(be, Z, g, "augment/Z")
    close(be.dense(Z >> 2), x8, rtol=1e-12, what="(X+31*0)>>2 (test_mps.py:22-38)")
    S = X + X
    close(be.dense(S), 2 * x8, rtol=1e-12, what="X+X (test_mps.py:51-70)")
    same_meta(be, S, g, "mps_add/Z")


def case_mpo_add_mul(be):
    g = load("known_answers")
    xa = np.arange(1, 17).reshape(2, 2, 2, 2).astype(float)
    ya = np.arange(18, 34).reshape(2, 2, 2, 2).astype(float)
    X = be.mpo_dense(xa, (3,))
    Y = be.mpo_dense(ya, (3,))
    Z = X + Y
    close(be.dense(Z), xa + ya, rtol=1e-11, what="X+Y (test_mpo.py:256-281)")
    close(be.dense(Z), g["mpo_add/dense"], what="X+Y golden")
    same_meta(be, Z, g, "mpo_add/Z")
    Z9 = X
    for _ in range(9):
        Z9 = Z9 + Y
    close(be.dense(Z9), g["mpo_add/dense9"], what="X+9Y golden")
    same_meta(be, Z9, g, "mpo_add/Z9")
    xm = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    XM = be.mpo_dense(xm, (2,))
    YM = be.mpo_dense(xm, (2,))
    ZM = XM * YM
    # Hadamard product is LOSSY in the reference (Kron bond 4 truncated to 2; test_mpo.py:211-233)
    close(be.dense(ZM), g["mpo_mul/dense"], what="X*Y golden")
    assert np.max(np.abs(be.dense(ZM) - xm * xm)) > 1.0
    same_meta(be, ZM, g, "mpo_mul/Z")


def case_mpo_orthogonality(be):
    g = load("known_answers")
    wo = np.arange(4 ** 6).reshape((4,) * 6).astype(float)
    W = be.mpo_dense(wo, (2, 2))
    W.left_orthonormalization()
    close(np.diag(be.arr(W.left_orthogonality(0))), [1, 1], what="left_orthogonality(0) (test_mpo.py:66-112)")
    close(np.diag(be.arr(W.left_orthogonality(1))), [1, 1], what="left_orthogonality(1)")
    close(be.arr(W.left_orthogonality(0)), np.eye(2), rtol=1e-12, what="left gram 0")
    close(be.arr(W[(1, 0, 0), (1, 0, 0)]).reshape(-1), [1040.0], what="element")
    close(be.arr(W[(1, 0, 0), (1, 0, 0)]).reshape(-1), g["orth/elem"], what="element golden")
    W.right_orthonormalization()
    close(be.arr(W.right_orthogonality(1)), np.eye(2), rtol=1e-12, what="right gram 1")
    close(be.arr(W.right_orthogonality(2)), np.eye(2), rtol=1e-12, what="right gram 2")
    close(be.arr(W[(1, 0, 0), (1, 0, 0)]).reshape(-1), [1040.0], what="element after right")
    yd = np.zeros((4, 4)); np.fill_diagonal(yd, [1, 2, 3, 4]); yd = yd.reshape(2, 2, 2, 2)
    Y = be.mpo_dense(yd, (4,)).right_orthonormalization()
    close(be.dense(Y), yd, rtol=1e-12, what="diag MPO (test_mpo.py:296-327)")
    close(be.arr(Y.right_orthogonality(1)), np.eye(4), rtol=1e-12, what="right_orthogonality(1)=I")


def case_mpo_decompose_compress(be):
    g = load("known_answers")
    wd = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    W = be.mpo_dense(wd, (4,))
    close(be.arr(W[(1, 1), (1, 1)]).reshape(-1), [wd[1, 1, 1, 1]], what="W[(1,1),(1,1)] (test_mpo.py:44-64)")
    zd = np.arange(4 ** 3 * 3 ** 3).reshape(4, 4, 4, 3, 3, 3).astype(float)
    Z = be.mpo_dense(zd, (3, 3))
    close(be.arr(Z[(3, 1, 0), (1, 2, 1)]).reshape(-1), [zd[3, 1, 0, 1, 2, 1]], what="Z[(3,1,0),(1,2,1)]")
    same_meta(be, Z, g, "decompose/Z")
    wc = np.arange(16 ** 4).reshape(16, 16, 16, 16).astype(float)
    WC = be.mpo_dense(wc, (8,))
    same_meta(be, WC, g, "mpo_compress/W8")
    assert WC.compress(4, mode="left") is None          # non-strict: in place, returns None (test_mpo.py:115-158)
    same_meta(be, WC, g, "mpo_compress/W4")
    close(be.dense(WC), g["mpo_compress/W4_dense"], what="compress(4) dense")
    Zc = WC >> 2
    assert [tuple(s.shape) for s in be.sites(Zc)] == [(1, 16, 16, 2), (2, 16, 16, 1)]
    close(be.dense(Zc), wc, rtol=1e-11, what=">>2 lossless on rank-2 arange")


def case_random_chain_r0(be):
    g = load("random_chains")
    A = be.mpo_sites(chain(g, "r0/A")); B = be.mpo_sites(chain(g, "r0/B"))
    for name, res in (("AB", A @ B), ("ApB", A + B), ("AhB", A * B), ("mulAB", be.mul(A, B))):
        close(be.dense(res), g["r0/%s_dense" % name], what="r0 " + name)
        same_meta(be, res, g, "r0/" + name)


def case_random_chain_r1(be):
    g = load("random_chains")
    X = be.mps_sites(chain(g, "r1/X")); W = be.mpo_sites(chain(g, "r1/W"))
    V = be.mps_sites(chain(g, "r1/V")); W2 = be.mpo_sites(chain(g, "r1/W2"))
    probes = g["r1/probes"]

    def sample(mp):
        return np.array([be.arr(mp[tuple(int(i) for i in p[0]), tuple(int(i) for i in p[1])]).reshape(-1)[0] for p in probes])

    Y = W @ X
    close(be.dense(Y), g["r1/WX_dense"], what="r1 W@X dense")
    same_meta(be, Y, g, "r1/WX")
    close(be.scalar(Y | V), g["r1/WX_V"], what="r1 (W@X)|V")
    close(be.scalar(X | V), g["r1/X_V"], what="r1 X|V")
    WW = W @ W2
    close(sample(WW), g["r1/WW2_elems"], what="r1 W@W2 elements"); same_meta(be, WW, g, "r1/WW2")
    S = W + W2
    close(sample(S), g["r1/WpW2_elems"], what="r1 W+W2 elements"); same_meta(be, S, g, "r1/WpW2")
    H = W * W2
    close(sample(H), g["r1/WhW2_elems"], what="r1 W*W2 elements"); same_meta(be, H, g, "r1/WhW2")
    A = X + V
    same_meta(be, A, g, "r1/XpV")
    close(be.dense(A), be.dense(X) + be.dense(V), rtol=1e-12, what="r1 X+V")
    for q in (1, 2, 3):
        Zq = V >> q
        same_meta(be, Zq, g, "r1/V_rs%d" % q)
        close(be.dense(Zq), be.dense_of(chain(g, "r1/V_rs%d" % q)), what="r1 V>>%d" % q)
    assert (X >> 7) is X                                   # guard returns the operand itself
    Yi = Y >> 4                                            # stale metadata lets `>>` inflate 3 -> 4
    same_meta(be, Yi, g, "r1/WX_rs4")
    close(be.dense(Yi), g["r1/WX_rs4_dense"], what="r1 (W@X)>>4 (inflation)")
    L = X.copy(); L.left_orthonormalization()
    close(be.dense(L), be.dense(X), rtol=1e-12, what="left canonical keeps the state")
    for k in range(4):
        close(be.arr(L.left_orthogonality(k)), np.eye(be.sites(L)[k].shape[-1]), rtol=1e-12, what="left gram")
    R = X.copy(); R.right_orthonormalization()
    close(be.dense(R), be.dense(X), rtol=1e-12, what="right canonical keeps the state")
    for k in range(1, 5):
        close(be.arr(R.right_orthogonality(k)), np.eye(be.sites(R)[k].shape[0]), rtol=1e-12, what="right gram")
    Lw = be.mpo_sites(chain(g, "r1/W")).left_orthonormalization()
    close(sample(Lw), sample(W), what="MPO left canonical keeps the operator")
    Rw = be.mpo_sites(chain(g, "r1/W")).right_orthonormalization()
    close(sample(Rw), sample(W), what="MPO right canonical keeps the operator")
    C1 = V.copy(); C1.compress(3, mode="left")
    close(be.dense(C1), be.dense_of(chain(g, "r1/V_c3left")), what="non-strict compress left")
    C2 = V.copy(); C2.compress(3, mode="right")
    close(be.dense(C2), be.dense_of(chain(g, "r1/V_c3right")), what="non-strict compress right")
    close(be.dense(be.mul(X, W)), g["r1/mul_XW"], what="syn.mul(X,W)")
    M2 = be.mul(W, W2)
    close(sample(M2), g["r1/mul_WW2_elems"], what="syn.mul(W,W2) elements"); same_meta(be, M2, g, "r1/mul_WW2")
    close(be.arr(X[(1, 2, 3, 0, 1)]).reshape(-1), g["r1/X_elem"], what="X[idx]")
    close(be.arr(W[(1, 2, 3, 0, 1), (3, 2, 1, 0, 2)]).reshape(-1), g["r1/W_elem"], what="W[idx]")


def case_random_chain_r2(be):
    g = load("random_chains")
    X = be.mps_sites(chain(g, "r2/X")); W = be.mpo_sites(chain(g, "r2/W")); V = be.mps_sites(chain(g, "r2/V"))
    Y = W @ X
    same_meta(be, Y, g, "r2/WX")
    close(be.scalar(Y | V), g["r2/WX_V"], what="r2 (W@X)|V")
    close(be.scalar(Y | Y), g["r2/WX_WX"], what="r2 (W@X)|(W@X)")
    Y2 = W @ (W @ X)
    same_meta(be, Y2, g, "r2/WWX")
    close(be.scalar(Y2 | V), g["r2/WWX_V"], what="r2 (W@W@X)|V")
    close(be.dense(Y2), be.dense_of(chain(g, "r2/WWX")), what="r2 W@W@X dense")


def case_readme_chain(be, from_dense=False):
    """readme.md:42-74 with np.random.seed(0).  Starts from the reference's decomposed
    cores (W0, X0, T0, U0) so that legacy-RNG streams and the 134 MB dense tensors are not needed."""
    g = load("readme_chain")
    W = be.mpo_sites(chain(g, "W0")); X = be.mps_sites(chain(g, "X0"))
    T = be.mpo_sites(chain(g, "T0")); U = be.mps_sites(chain(g, "U0"))
    # objects built by the dense ctor carry the DECLARED bonds as metadata; from_sites derives the same here
    W = W >> 4
    T = T >> 2
    same_meta(be, W, g, "W4"); same_meta(be, T, g, "T2")          # bond_shape stale: (16,16) and (8,8)
    TW = T + W
    same_meta(be, TW, g, "TpW")                                   # bonds 6, `>> 8` is a no-op
    TWT = TW @ T
    same_meta(be, TWT, g, "TpW_T")                                # actual 6, metadata (12,12)
    Z = TWT @ X
    same_meta(be, Z, g, "Z")                                      # actual 4, metadata (24,24)
    close(be.scalar(X | U), g["X_U"], what="X|U")
    close(be.scalar(Z | X), g["Z_X"], rtol=1e-9, what="Z|X")
    close(be.dense(Z), be.dense_of(chain(g, "Z")), rtol=1e-9, what="Z dense")
    Z = Z >> 16
    same_meta(be, Z, g, "Z16")                                    # inflated 4 -> 16, metadata (4,4)
    close(be.scalar(Z | X), g["Z16_X"], rtol=1e-9, what="(Z>>16)|X")
    Z.left_orthonormalization()
    close(np.diag(be.arr(Z.left_orthogonality(0))), np.ones(16), rtol=1e-12, what="diag left 0")
    close(np.diag(be.arr(Z.left_orthogonality(1))), np.ones(16), rtol=1e-12, what="diag left 1")
    close(be.scalar(Z | X), g["Z16_X"], rtol=1e-9, what="(Z>>16 canonical)|X")


QUANTUM_CIRCUITS = {
    "bell": (2, [("H", 0), ("CX", 0, 1)]),
    "ghz4": (4, [("H", 0), ("CX", 0, 1), ("CX", 1, 2), ("CX", 2, 3)]),
    "x_chain": (5, [("X", 0), ("X", 3), ("CX", 3, 4), ("SWAP", 1)]),
    "cx_far_02": (3, [("X", 0), ("CX", 0, 2)]),
    "cx_far_20": (3, [("X", 2), ("CX", 2, 0)]),
    "cx_far_03": (4, [("X", 0), ("CX", 0, 3)]),
    "h_layer": (4, [("H", 0), ("H", 1), ("H", 2), ("H", 3), ("Z", 1), ("CX", 1, 2)]),
    "toffoli_110": (3, [("X", 0), ("X", 1), ("TOFFOLI", 0)]),
    "toffoli_100": (3, [("X", 0), ("TOFFOLI", 0)]),
}


def case_mps_apply_gates(be):
    """MatrixProductState.apply (MPS:487-534): the bond is frozen at its existing value; goldens from the reference."""
    g = load("quantum")
    cores = chain(g, "apply/X")
    for name, gname, i in (("cx0", "CX", 0), ("swap2", "SWAP", 2), ("h4", "H", 4), ("tof1", "TOFFOLI", 1), ("x0", "X", 0),
                           ("cx3", "CX", 3), ("tof2", "TOFFOLI", 2), ("z2", "Z", 2)):
        dense, bonds = be.apply_gate(cores, gname, i)
        close(dense, g["apply/%s_dense" % name], what="apply " + name)
        assert list(bonds) == [int(b) for b in g["apply/%s_bonds" % name]]


def case_quantum_circuits(be):
    """Qbit registers (quantum/qbit.py) on real gates: Bell (test_quantum.py:155-163), GHZ-4, non-adjacent CX through the
    reference's swap_in / swap_out logic, Toffoli truth table rows (readme.md:97-117)."""
    g = load("quantum")
    for name, (size, ops) in QUANTUM_CIRCUITS.items():
        close(be.run_circuit(size, ops), g["circuit/%s" % name], rtol=1e-12, what="circuit " + name)
    close(be.run_circuit(2, QUANTUM_CIRCUITS["bell"][1]), np.array([1, 0, 0, 1]) / np.sqrt(2), rtol=1e-12, what="Bell state")
    close(be.run_circuit(3, QUANTUM_CIRCUITS["toffoli_110"][1]), np.eye(8)[7], rtol=1e-12, what="Toffoli |110> -> |111>")


def case_mpo_apply_range(be):
    """MatrixProductOperator.apply(operator, indices) (MPO:582-626): exact for one site, the reference's literal re-split beyond."""
    g = load("mpo_apply")
    for name in [str(n) for n in g["names"]]:
        A = be.mpo_sites(chain(g, name + "/A"))
        O = be.mpo_sites(chain(g, name + "/op"))
        ret = A.apply(O, [int(i) for i in g[name + "/indices"]])
        assert ret is None
        assert [tuple(int(x) for x in s.shape) for s in be.sites(A)] == [tuple(int(x) for x in r) for r in g[name + "/after_shapes"]]
        T = None
        for c in be.sites(A):                              # raw chain product (l, in_0, out_0, in_1, out_1, ..., r), as the golden
            T = c if T is None else np.tensordot(T, c, axes=(T.ndim - 1, 0))
        close(T, g[name + "/after_dense"], what="apply " + name)


ALL_CASES = [
    case_mpo_apply_range,
    case_matmul_known_answers,
    case_mps_dot_compress_normalize,
    case_mps_add_augment,
    case_mpo_add_mul,
    case_mpo_orthogonality,
    case_mpo_decompose_compress,
    case_random_chain_r0,
    case_random_chain_r1,
    case_random_chain_r2,
    case_readme_chain,
    case_mps_apply_gates,
    case_quantum_circuits,
]
