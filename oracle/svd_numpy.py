"""TEST INFRASTRUCTURE ONLY -- textbook SVD bond rounding (CPU, numpy/LAPACK).

PARITY UNPINNED BY THE REFERENCE: antoine311200/Syngular has no SVD on its live path (its `>>` is the
QR-truncation restated in ref_numpy.py; the only SVD splits are dead code in trash/mpo.py:59-190 and
experimental/layers.py:241-325, and the density-matrix `@` of matrix_product_operator.py:193-260 is unfinished
and returns None).  BASELINE.json's north_star nevertheless asks for "QR left/right-orthonormalization followed
by SVD truncation to the target bond dimension", so this file states that algorithm in its canonical textbook
form and the CUDA path is checked against it on gauge-invariant quantities.  It is cross-checked only by
invariants (tests/test_oracle_svd.py): kept cores left-orthonormal, spectra equal to the dense SVD of the
unfolding for small N, discarded weight == squared error, optimality versus the QR-truncation.

Algorithm (Schollwoeck 2011, sec. 4.5; Oseledets 2011 "TT-rounding"):
  1. right-to-left QR sweep  -> all cores k>=1 right-orthonormal
  2. left-to-right sweep: M = core_k.reshape(l*d, r); U,S,Vt = svd(M); keep the leading
     k = min(chi_max, #{s_i > cutoff * s_0}, ) (at least 1) triplets; core_k <- U_k; core_{k+1} <- diag(S_k) Vt_k @ core_{k+1}.
The result is left-canonical, with the norm carried by the last core.
"""
from __future__ import annotations

import numpy as np

from .ref_numpy import right_orthonormalize, site_mpo_mps, site_mpo_mpo


def keep_count(S, chi_max, cutoff):
    k = int(np.count_nonzero(S > cutoff * S[0])) if S[0] > 0 else 1
    return max(1, min(int(chi_max), k, len(S)))


def round_svd(cores, chi_max, cutoff=0.0, canonicalize=True):
    """Returns (new cores, list of per-bond singular value arrays (all of them, before the cut),
    list of per-bond discarded weights sum_{i>=k} s_i^2)."""
    cores = right_orthonormalize(cores) if canonicalize else [np.array(c, copy=True) for c in cores]
    spectra, discarded = [], []
    for k in range(len(cores) - 1):
        cur, nxt = cores[k], cores[k + 1]
        M = cur.reshape(-1, cur.shape[-1])
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        keep = keep_count(S, chi_max, cutoff)
        spectra.append(S.copy())
        discarded.append(float(np.sum(S[keep:] ** 2)))
        carry = (S[:keep, None] * Vt[:keep]) @ nxt.reshape(nxt.shape[0], -1)
        cores[k] = U[:, :keep].reshape(cur.shape[:-1] + (keep,))
        cores[k + 1] = carry.reshape((keep,) + nxt.shape[1:])
    return cores, spectra, discarded


def apply_round_svd(X, W, chi_max, cutoff=0.0):
    """MPO x MPS (ref site contraction K1) followed by round_svd.  X: MPS cores, W: MPO cores."""
    prod = [site_mpo_mps(x, w) for x, w in zip(X, W)]
    return round_svd(prod, chi_max, cutoff)


def apply_round_svd_mpo(A, B, chi_max, cutoff=0.0):
    prod = [site_mpo_mpo(a, b) for a, b in zip(A, B)]
    return round_svd(prod, chi_max, cutoff)


def apply_round_density_matrix(X, W, chi_max, cutoff=0.0):
    """The same truncation computed the way the CUDA fast path does it (never forming the D x D factors):
    right environments E_k = Gram matrix of the product chain right of bond k, then a left-to-right sweep
    diagonalising M E M^T.  Mathematically identical to apply_round_svd (in exact arithmetic, up to gauge);
    this is the finished form of the reference's abandoned MATMUL_MODE == "opti" branch
    (matrix_product_operator.py:193-260).  Used in tests to separate algorithmic from kernel error."""
    n = len(X)
    prod = [site_mpo_mps(x, w) for x, w in zip(X, W)]
    E = [None] * (n + 1)
    E[n] = np.ones((1, 1))
    for k in range(n - 1, 0, -1):
        C = prod[k]
        T = np.tensordot(C, E[k + 1], axes=(2, 0))             # (l, o, r')
        E[k] = np.tensordot(T, C, axes=([1, 2], [1, 2]))        # (l, l')
    carry = np.ones((1, 1))
    out, spectra, discarded = [], [], []
    for k in range(n - 1):
        C = prod[k]
        M = np.tensordot(carry, C, axes=(1, 0))                 # (s, o, r)
        s, o, r = M.shape
        M2 = M.reshape(s * o, r)
        A = M2 @ E[k + 1] @ M2.T
        A = 0.5 * (A + A.T)
        lam, U = np.linalg.eigh(A)
        lam, U = lam[::-1], U[:, ::-1]
        S = np.sqrt(np.clip(lam, 0.0, None))
        keep = keep_count(S, chi_max, cutoff)
        spectra.append(S.copy())
        discarded.append(float(np.sum(S[keep:] ** 2)))
        out.append(U[:, :keep].reshape(s, o, keep))
        carry = U[:, :keep].T @ M2
    C = prod[n - 1]
    out.append(np.tensordot(carry, C, axes=(1, 0)))
    return out, spectra, discarded


class StructuredDensityMatrixSweep:
    """apply_round_density_matrix, with every contraction done through the (X, W) structure so that no product core
    (D x d x D) is formed -- the exact sequence of contractions the CUDA fast path executes, in numpy/BLAS -- and cut into its
    2n - 1 unit operations so that a caller can time any contiguous part of ONE real sweep:

        env_step(k)   k = n-1 .. 1   right environment E[k] from E[k+1]            (phase 1, right to left)
        site_step(k)  k = 0 .. n-1   carry . C_k, Gram matrix, eigh, truncation    (phase 2, left to right)

    `operations()` lists them in execution order; running all of them IS the sweep (apply_round_density_matrix_structured
    below does exactly that).  bench.py times this on the host cores, whole (cpu_baseline) or in K consecutive chunks
    (--impl reference), and tests/test_gpu_fullsize.py compares the CUDA sweep with it at BASELINE configs[1]'s full size."""

    def __init__(self, X, W, chi_max, cutoff=0.0):
        self.X, self.W, self.chi_max, self.cutoff = X, W, chi_max, cutoff
        self.n = len(X)
        self.E = [None] * (self.n + 1)
        self.E[self.n] = np.ones((1, 1))
        self.carry = np.ones((1, 1, 1))                                 # (s, a, l)
        self.out, self.spectra, self.discarded = [], [], []

    def operations(self):
        return [("env", k) for k in range(self.n - 1, 0, -1)] + [("site", k) for k in range(self.n)]

    def run(self, op):
        kind, k = op
        (self.env_step if kind == "env" else self.site_step)(k)

    def op_flops(self, op):
        """Dominant-term flop count of one unit operation (used only to balance chunks of a sweep, never to scale a timing)."""
        kind, k = op
        a, i, b = self.X[k].shape
        l, _, o, r = self.W[k].shape
        D, Dl = b * r, a * l
        if kind == "env":
            return 2.0 * a * i * b * r * D + 2.0 * l * o * a * D * i * r + 2.0 * l * a * b * l * i * o * r + 2.0 * Dl * Dl * i * b
        s = min(self.chi_max, 2 ** min(k, 40))
        return 2.0 * s * a * l * i * b + 2.0 * s * o * D * D + 2.0 * (s * o) ** 2 * D + 10.0 * (s * o) ** 3

    def env_step(self, k):
        Xk, Wk = self.X[k], self.W[k]
        a, i, b = Xk.shape
        l, _, o, r = Wk.shape
        En = self.E[k + 1].reshape(b, r, b * r)
        P1 = np.tensordot(Xk, En, axes=(2, 0))                         # (a, i, r, y)
        P2 = np.tensordot(Wk, P1, axes=([1, 3], [1, 2]))               # (l, o, a, y)
        P2 = P2.reshape(l, o, a, b, r)
        Z = np.tensordot(P2, Wk, axes=([1, 4], [2, 3]))                # (l, a, b', l', i')
        Ek = np.tensordot(Z, Xk, axes=([4, 2], [1, 2]))                # (l, a, l', a')
        self.E[k] = Ek.transpose(1, 0, 3, 2).reshape(a * l, a * l)

    def site_step(self, k):
        Xk, Wk = self.X[k], self.W[k]
        a, i, b = Xk.shape
        l, _, o, r = Wk.shape
        T1 = np.tensordot(self.carry, Xk, axes=(1, 0))                 # (s, l, i, b)
        M = np.tensordot(T1, Wk, axes=([1, 2], [0, 1]))                # (s, b, o, r)
        M = M.transpose(0, 2, 1, 3)                                    # (s, o, b, r)
        s = M.shape[0]
        if k == self.n - 1:
            self.out.append(M.reshape(s, o, b * r))
            return
        M2 = M.reshape(s * o, b * r)
        A = M2 @ self.E[k + 1] @ M2.T
        A = 0.5 * (A + A.T)
        lam, U = np.linalg.eigh(A)
        lam, U = lam[::-1], U[:, ::-1]
        S = np.sqrt(np.clip(lam, 0.0, None))
        keep = keep_count(S, self.chi_max, self.cutoff)
        self.spectra.append(S.copy())
        self.discarded.append(float(np.sum(S[keep:] ** 2)))
        self.out.append(U[:, :keep].reshape(s, o, keep))
        self.carry = (U[:, :keep].T @ M2).reshape(keep, b, r)
        self.E[k + 1] = None                                           # consumed: release the host memory


def apply_round_density_matrix_structured(X, W, chi_max, cutoff=0.0):
    """The whole structured sweep (see StructuredDensityMatrixSweep): returns (cores, spectra, discarded) like the others."""
    sweep = StructuredDensityMatrixSweep(X, W, chi_max, cutoff)
    for op in sweep.operations():
        sweep.run(op)
    return sweep.out, sweep.spectra, sweep.discarded
