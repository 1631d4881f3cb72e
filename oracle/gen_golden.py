"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden
The fixtures are committed; the GPU box has neither the reference nor numpy's legacy RNG stream
guaranteed, so every input the reference consumed is stored next to every output it produced.

Each fixture is a flat npz: "<case>/<object>/site<k>" core arrays, "<case>/<object>/bond_shape"
(the reference's -- possibly stale -- metadata), "<case>/<object>/shape", and scalars / dense tensors
under "<case>/<name>".  Cases follow the reference's own tests (test/test_mpo.py, test/test_mps.py,
test/test_syn.py), the README chain (readme.md:42-74, np.random.seed(0)) and seeded random chains.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import shim
    syn = shim.load_reference()
    from syngular.tensor import MatrixProductState as MPS, MatrixProductOperator as MPO

    def quiet(fn, *a, **k):
        with contextlib.redirect_stdout(io.StringIO()):
            return fn(*a, **k)

    def put(store, prefix, mp):
        for k, s in enumerate(mp.sites):
            store["%s/site%d" % (prefix, k)] = np.array(s)
        store[prefix + "/bond_shape"] = np.array(mp.bond_shape, dtype=np.int64)
        store[prefix + "/shape"] = np.array([tuple(s) for s in mp.shape], dtype=np.int64)
        store[prefix + "/n"] = np.array(len(mp.sites))

    # ------------------------------------------------------------------ known-answer cases
    ka = {}
    x = np.arange(4).reshape(2, 2).astype(float)
    w = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    X = MPS(x, bond_shape=(2,)).decompose()
    W = quiet(MPO(w, bond_shape=(2,)).decompose)
    put(ka, "matmul/X", X); put(ka, "matmul/W", W)
    Y = W @ X
    put(ka, "matmul/WX", Y)
    ka["matmul/WX_dense"] = Y.to_tensor().real
    ka["matmul/XX"] = np.array(X | X)
    Z = W @ W @ X
    put(ka, "matmul/WWX", Z)
    ka["matmul/WWX_dense"] = Z.to_tensor().real
    WW = W @ W
    put(ka, "matmul/WW", WW)
    ka["matmul/WW_dense"] = WW.to_tensor()
    ka["syn/mul_WX"] = syn.mul(W, X).to_tensor().real
    ka["syn/mul_XW"] = syn.mul(X, W).to_tensor().real
    ka["syn/mul_XX"] = np.array(syn.mul(X, X))
    ka["syn/mul_WW"] = syn.mul(W, W).to_tensor()

    x = np.arange(64).reshape(4, 4, 4).astype(float)
    X = MPS(x, bond_shape=(4, 4)).decompose()
    put(ka, "dot/X", X)
    ka["dot/XX"] = np.array(X | X)
    ka["dot/norm"] = np.array(X.dot())
    Z = X >> 2
    put(ka, "compress/Z", Z)
    ka["compress/Z_dense"] = Z.to_tensor().real
    Xn = MPS(x, bond_shape=(4, 4)).decompose().normalize()
    put(ka, "normalize/X", Xn)
    ka["normalize/norm"] = np.array(Xn.dot())
    for mode in ("left", "right"):
        Xm = MPS(x, bond_shape=(2, 2)).decompose(mode=mode)
        put(ka, "decompose_%s/X" % mode, Xm)
        ka["decompose_%s/dense" % mode] = Xm.to_tensor().real

    x8 = np.arange(8).reshape(2, 2, 2).astype(float)
    X8 = MPS(x8, bond_shape=(2, 2)).decompose()
    Y0 = MPS.zeros((2, 2, 2), (2, 2))
    Z = X8
    for _ in range(31):
        Z = Z + Y0
    put(ka, "augment/Z", Z)
    ka["augment/Z2_dense"] = (Z >> 2).to_tensor().real
    put(ka, "augment/X", X8)
    Zadd = X8 + X8
    put(ka, "mps_add/Z", Zadd)
    ka["mps_add/dense"] = Zadd.to_tensor().real

    xa = np.arange(1, 17).reshape(2, 2, 2, 2).astype(float)
    ya = np.arange(18, 34).reshape(2, 2, 2, 2).astype(float)
    XA = quiet(MPO(xa, bond_shape=(3,)).decompose)
    YA = quiet(MPO(ya, bond_shape=(3,)).decompose)
    put(ka, "mpo_add/X", XA); put(ka, "mpo_add/Y", YA)
    ZA = XA + YA
    put(ka, "mpo_add/Z", ZA)
    ka["mpo_add/dense"] = ZA.to_tensor()
    Z9 = XA
    for _ in range(9):
        Z9 = Z9 + YA
    put(ka, "mpo_add/Z9", Z9)
    ka["mpo_add/dense9"] = Z9.to_tensor()

    xm = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    XM = quiet(MPO(xm, bond_shape=(2,)).decompose)
    YM = quiet(MPO(xm, bond_shape=(2,)).decompose)
    ZM = XM * YM
    put(ka, "mpo_mul/X", XM)
    put(ka, "mpo_mul/Z", ZM)
    ka["mpo_mul/dense"] = ZM.to_tensor()

    yd = np.zeros((4, 4)); np.fill_diagonal(yd, [1, 2, 3, 4]); yd = yd.reshape(2, 2, 2, 2)
    YD = quiet(MPO(yd, bond_shape=(4,)).decompose).right_orthonormalization()
    put(ka, "compare/Y", YD)
    ka["compare/dense"] = YD.to_tensor()
    ka["compare/right_orth1"] = YD.right_orthogonality(1)

    wo = np.arange(4 ** 6).reshape((4,) * 6).astype(float)
    WO = quiet(MPO(wo, bond_shape=(2, 2)).decompose)
    put(ka, "orth/W0", WO)
    WO.left_orthonormalization()
    put(ka, "orth/Wleft", WO)
    ka["orth/left0"] = WO.left_orthogonality(0)
    ka["orth/left1"] = WO.left_orthogonality(1)
    ka["orth/elem"] = np.array(WO[(1, 0, 0), (1, 0, 0)]).reshape(-1)
    WO.right_orthonormalization()
    put(ka, "orth/Wright", WO)
    ka["orth/right1"] = WO.right_orthogonality(1)
    ka["orth/right2"] = WO.right_orthogonality(2)

    wd = np.arange(16).reshape(2, 2, 2, 2).astype(float)
    WD = quiet(MPO(wd, bond_shape=(4,)).decompose)
    zd = np.arange(4 ** 3 * 3 ** 3).reshape(4, 4, 4, 3, 3, 3).astype(float)
    ZD = quiet(MPO(zd, bond_shape=(3, 3)).decompose)
    put(ka, "decompose/W", WD); put(ka, "decompose/Z", ZD)
    ka["decompose/W1111"] = np.array(WD[(1, 1), (1, 1)]).reshape(-1)
    ka["decompose/Z310121"] = np.array(ZD[(3, 1, 0), (1, 2, 1)]).reshape(-1)

    wc = np.arange(16 ** 4).reshape(16, 16, 16, 16).astype(float)
    WC = quiet(MPO(wc, bond_shape=(8,)).decompose)
    put(ka, "mpo_compress/W8", WC)
    quiet(WC.compress, 4, mode="left")                # non-strict, in place
    put(ka, "mpo_compress/W4", WC)
    ka["mpo_compress/W4_dense"] = to_dense_mpo(WC.sites)
    np.savez_compressed(os.path.join(OUT, "known_answers.npz"), **ka)
    print("known_answers:", len(ka), "arrays")

    # ------------------------------------------------------------------ seeded random chains
    rnd = {}
    rng = np.random.default_rng(12345)

    def rand_mps(n, d, chi):
        b = [1] + [chi] * (n - 1) + [1]
        return MPS.from_sites([rng.normal(size=(b[k], d, b[k + 1])) / np.sqrt(b[k] * d) for k in range(n)])

    def rand_mpo(n, di, do, chi):
        b = [1] + [chi] * (n - 1) + [1]
        return MPO.from_sites([rng.normal(size=(b[k], di, do, b[k + 1])) / np.sqrt(b[k] * di) for k in range(n)])

    probe_rng = np.random.default_rng(777)
    probes5 = probe_rng.integers(0, 4, size=(48, 2, 5))
    rnd["r1/probes"] = probes5

    def sample(mp):
        """gauge-invariant fingerprint of an MPO too large to store densely: 48 fixed elements"""
        return np.array([np.array(mp[tuple(p[0]), tuple(p[1])]).reshape(-1)[0] for p in probes5])

    # case r0: N=3, d=3 -- small enough to keep dense MPO results
    n, d = 3, 3
    A0 = rand_mpo(n, d, d, 3); B0 = rand_mpo(n, d, d, 2)
    put(rnd, "r0/A", A0); put(rnd, "r0/B", B0)
    for name, res in (("AB", A0 @ B0), ("ApB", A0 + B0), ("AhB", A0 * B0), ("mulAB", syn.mul(A0, B0))):
        put(rnd, "r0/" + name, res)
        rnd["r0/%s_dense" % name] = to_dense_mpo(res.sites)

    # case r1: N=5, d=4: MPO(chi 3) @ MPS(chi 4) -> bonds 12 -> >> 3 ; lossy, well conditioned
    n, d = 5, 4
    X = rand_mps(n, d, 4); W = rand_mpo(n, d, d, 3); V = rand_mps(n, d, 4); W2 = rand_mpo(n, d, d, 2)
    put(rnd, "r1/X", X); put(rnd, "r1/W", W); put(rnd, "r1/V", V); put(rnd, "r1/W2", W2)
    Y = W @ X
    put(rnd, "r1/WX", Y)
    rnd["r1/WX_dense"] = Y.to_tensor().real
    rnd["r1/WX_V"] = np.array(Y | V)
    rnd["r1/X_V"] = np.array(X | V)
    WW = W @ W2
    put(rnd, "r1/WW2", WW)
    rnd["r1/WW2_elems"] = sample(WW)
    S = W + W2
    put(rnd, "r1/WpW2", S)
    rnd["r1/WpW2_elems"] = sample(S)
    H = W * W2
    put(rnd, "r1/WhW2", H)
    rnd["r1/WhW2_elems"] = sample(H)
    A = X + V
    put(rnd, "r1/XpV", A)
    for q in (1, 2, 3):
        Zq = V >> q
        put(rnd, "r1/V_rs%d" % q, Zq)
    Zi = X >> 7            # guard: 7 >= 4 -> returns self
    rnd["r1/X_rs7_is_self"] = np.array(Zi is X)
    Yi = Y >> 4            # Y has stale bond_shape (12,..): 4 < 12 -> runs and INFLATES 3 -> 4 (l*d = 4 at site 0)
    put(rnd, "r1/WX_rs4", Yi)
    rnd["r1/WX_rs4_dense"] = Yi.to_tensor().real
    L = X.copy(); quiet(L.left_orthonormalization)
    put(rnd, "r1/X_left", L)
    R = X.copy(); R.right_orthonormalization()
    put(rnd, "r1/X_right", R)
    Lw = MPO.from_sites([s.copy() for s in W.sites]).left_orthonormalization()
    put(rnd, "r1/W_left", Lw)
    Rw = MPO.from_sites([s.copy() for s in W.sites]).right_orthonormalization()
    put(rnd, "r1/W_right", Rw)
    C1 = V.copy(); quiet(C1.compress, 3, mode="left")
    put(rnd, "r1/V_c3left", C1)
    C2 = V.copy(); quiet(C2.compress, 3, mode="right")
    put(rnd, "r1/V_c3right", C2)
    rnd["r1/mul_XW"] = syn.mul(X, W).to_tensor().real
    M2 = syn.mul(W, W2)
    put(rnd, "r1/mul_WW2", M2)
    rnd["r1/mul_WW2_elems"] = sample(M2)
    rnd["r1/X_elem"] = np.array(X[(1, 2, 3, 0, 1)]).reshape(-1)
    rnd["r1/W_elem"] = np.array(W[(1, 2, 3, 0, 1), (3, 2, 1, 0, 2)]).reshape(-1)

    # case r2: longer chain, d=3, N=8, chi 6 x chiW 3 -> 18 -> >> 3
    n, d = 8, 3
    X = rand_mps(n, d, 3); W = rand_mpo(n, d, d, 3); V = rand_mps(n, d, 2)
    put(rnd, "r2/X", X); put(rnd, "r2/W", W); put(rnd, "r2/V", V)
    Y = W @ X
    put(rnd, "r2/WX", Y)
    rnd["r2/WX_V"] = np.array(Y | V)
    rnd["r2/WX_WX"] = np.array(Y | Y)
    Y2 = W @ (W @ X)
    put(rnd, "r2/WWX", Y2)
    rnd["r2/WWX_V"] = np.array(Y2 | V)
    np.savez_compressed(os.path.join(OUT, "random_chains.npz"), **rnd)
    print("random_chains:", len(rnd), "arrays")

    # ------------------------------------------------------------------ README chain (readme.md:42-74)
    rd = {}
    np.random.seed(0)
    tensor_W = np.arange(16 ** 6).reshape((16,) * 6)
    tensor_X = np.arange(16 ** 3).reshape((16,) * 3)
    W = MPO(tensor_W, bond_shape=(16, 16)); quiet(W.decompose)
    X = MPS(tensor_X, bond_shape=(4, 4)); X.decompose()
    T = quiet(MPO.random, (16, 16, 16), (16, 16, 16), (8, 8))
    U = MPS.random((16, 16, 16), (8, 8))
    # the reference's MPO `>>` overwrites the SOURCE's cores through list aliasing: snapshot before
    put(rd, "W0", MPO.from_sites([s.copy() for s in W.sites]))
    put(rd, "X0", X); put(rd, "T0", MPO.from_sites([s.copy() for s in T.sites])); put(rd, "U0", U)
    W = W >> 4
    T = T >> 2
    put(rd, "W4", W); put(rd, "T2", T)
    TW = T + W
    put(rd, "TpW", TW)
    TWT = TW @ T
    put(rd, "TpW_T", TWT)
    Z = TWT @ X
    put(rd, "Z", Z)
    rd["X_U"] = np.array(X | U)
    rd["Z_X"] = np.array(Z | X)
    Z = Z >> 16
    put(rd, "Z16", Z)
    rd["Z16_X"] = np.array(Z | X)
    quiet(Z.left_orthonormalization)
    put(rd, "Z16_left", Z)
    rd["diag_left0"] = np.diag(Z.left_orthogonality(0))
    rd["diag_left1"] = np.diag(Z.left_orthogonality(1))
    np.savez_compressed(os.path.join(OUT, "readme_chain.npz"), **rd)
    qd, _ = quantum_goldens(MPS, quiet)
    np.savez_compressed(os.path.join(OUT, "quantum.npz"), **qd)
    print("quantum:", len(qd), "arrays; bell =", qd["circuit/bell"], " toffoli|110> =", qd["circuit/toffoli_110"])
    print("readme_chain:", len(rd), "arrays; X|U =", repr(float(rd["X_U"])), " Z|X =", repr(float(rd["Z_X"])))


def quantum_goldens(MPS, quiet):
    """MatrixProductState.apply and Qbit circuits with REAL gates (the CUDA library is FP64-real), run by the reference."""
    from syngular.quantum import Qbit, gate
    q = {}
    rng = np.random.default_rng(99)
    b = [1, 2, 4, 4, 2, 1]
    cores = [rng.normal(size=(b[k], 2, b[k + 1])) for k in range(5)]
    Xr = MPS.from_sites(cores)
    for k, c in enumerate(cores):
        q["apply/X/site%d" % k] = c
    q["apply/X/n"] = np.array(5)
    for name, g, i in (("cx0", gate.CX, 0), ("swap2", gate.SWAP, 2), ("h4", gate.H, 4), ("tof1", gate.TOFFOLI, 1), ("x0", gate.X, 0),
                       ("cx3", gate.CX, 3), ("tof2", gate.TOFFOLI, 2), ("z2", gate.Z, 2)):
        Y = Xr.apply(g, i)
        q["apply/%s_dense" % name] = Y.to_tensor().real
        q["apply/%s_bonds" % name] = np.array([s.shape[2] for s in Y.sites[:-1]])
    circuits = {
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
    for name, (size, ops) in circuits.items():
        qb = Qbit(size)
        for op in ops:
            g = getattr(gate, op[0])
            qb = quiet(qb.__matmul__, (g,) + tuple(op[1:]))
        q["circuit/%s" % name] = np.asarray(qb.to_tensor()).real
    q["circuit/binary_101"] = np.asarray(Qbit.from_binary("101").to_tensor()).real
    return q, circuits


def to_dense_mpo(sites):
    """(reference `to_tensor` for an MPO is a d^2N-iteration Python loop; same numbers by one chain product)"""
    n = len(sites)
    T = sites[0].reshape(-1, sites[0].shape[-1])
    dims = list(sites[0].shape[1:3])
    for c in sites[1:]:
        T = (T @ c.reshape(c.shape[0], -1)).reshape(-1, c.shape[-1])
        dims += list(c.shape[1:3])
    T = T.reshape(dims)
    return T.transpose(list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2)))


if __name__ == "__main__":
    main()
