"""Behaviour shared by MatrixProductState and MatrixProductOperator (host logic only).

Metadata rules follow the reference exactly where later results depend on them:
  * from_sites derives shape / input_shape / bond_shape from the cores (MPS:191-214, MPO:357-380);
  * the object returned by `>>` keeps the PRE-truncation bond_shape (stale) while .shape is updated (MPS:433-468, MPO:545-580);
  * `>> q` with q >= min(bond_shape) returns the operand itself (MPS:367-368, MPO:473-474).
Deliberately NOT replicated: the MPO copy()/from_sites list aliasing that lets `>>` overwrite its source,
the rewrite of the source's bond_shape (MPO:477) in strict mode, and the hot-path prints (MPO:442, MPS:556).
"""
import numpy as np
import torch

from syngular.tensor import _sweeps as sw


class _MatrixProduct:
    PHYS = 1                 # number of physical legs per core
    ROUNDING = "qr"          # "qr": reference semantics ; "svd": optimal truncation
    SVD_CUTOFF = 0.0

    # ---- construction -------------------------------------------------------------------------------------
    def _init_empty(self, verbose=0):
        self.parameters_number = 0
        self.real_parameters_number = 0
        self.verbose = verbose
        self.decomposed = False
        self.orthonormalized = None
        self.truncation = None      # SVD mode: per-bond singular values / kept ranks / discarded weights of the last rounding

    @classmethod
    def from_sites(cls, sites, orthogonality=None, real_parameters_number=None):
        mp = cls()
        mp.sites = [sw.as_core(s) for s in sites]
        mp.sites_number = len(mp.sites)
        mp.decomposed = True
        mp.orthonormalized = orthogonality
        mp.real_parameters_number = real_parameters_number
        mp._refresh_from_cores(bonds=True)
        return mp

    def _refresh_from_cores(self, bonds):
        self.shape = [tuple(int(x) for x in s.shape) for s in self.sites]
        self.parameters_number = int(sum(int(np.prod(s)) for s in self.shape))
        self.input_shape = tuple(s[1] for s in self.shape)
        if self.PHYS == 2:
            self.output_shape = tuple(s[2] for s in self.shape)
        if bonds:
            self.bond_shape = tuple(s[-1] for s in self.shape[:-1])

    @staticmethod
    def _chain_shapes(phys, bond_shape):
        b = (1,) + tuple(int(x) for x in bond_shape) + (1,)
        return [(b[k],) + tuple(phys[k]) + (b[k + 1],) for k in range(len(phys))]

    def copy(self):
        """New container over the same (immutable-by-convention) cores; sweeps never write cores in place."""
        mp = type(self).from_sites(self.sites)
        return mp

    # ---- `>>` ---------------------------------------------------------------------------------------------
    def __rshift__(self, dim):
        if isinstance(dim, int) and not isinstance(dim, bool):
            return self.compress(dim, strict=True)
        raise Exception("dimension should be an integer")

    def _round(self, sites, dim):
        if self.ROUNDING == "svd":
            out, trunc = sw.round_svd(sites, dim, self.SVD_CUTOFF)
            return out, trunc
        return sw.round_qr(sites, dim), None

    def compress(self, dim, mode="left", strict=False):
        if dim >= min(self.bond_shape):
            return self
        if strict:
            that = self.copy()                       # bond_shape := actual bonds BEFORE truncation (stays stale)
            that.sites, that.truncation = self._round(self.sites, dim)
            that._refresh_from_cores(bonds=False)
            return that
        # non-strict (MPS:370-430, MPO:480-541): canonicalise unless flagged, then the same sweep from that side, in place
        if self.PHYS == 2:
            self.bond_shape = (dim,) * (self.sites_number - 1)           # MPO:477
        if mode == "left":
            if self.orthonormalized != "left":
                self.left_orthonormalization()
            self.sites = sw.round_qr(self.sites, dim)
        elif mode == "right":
            if self.orthonormalized != "right":
                self.right_orthonormalization()
            # mirror image of the left sweep: reverse the chain and swap the bond legs (views), sweep, mirror back
            self.sites = sw.unmirror(sw.round_qr(sw.mirror(self.sites), dim), self.sites)
        self._refresh_from_cores(bonds=False)
        return None

    # ---- canonical forms ----------------------------------------------------------------------------------
    def left_orthonormalization(self, bond_shape=()):
        self._require_decomposed("Cannot orthonormalize an undecomposed MatrixProductOperator")
        self.sites = sw.left_orthonormalize(self.sites)
        self._refresh_from_cores(bonds=False)
        self.orthonormalized = "left"
        return self if self.PHYS == 2 else None      # MPO returns self (MPO:694), MPS returns None (MPS:566)

    def right_orthonormalization(self, bond_shape=()):
        self._require_decomposed("Cannot orthonormalize an undecomposed MatrixProductOperator")
        self.sites = sw.right_orthonormalize(self.sites)
        self._refresh_from_cores(bonds=False)
        self.orthonormalized = "right"
        return self if self.PHYS == 2 else None

    def _require_decomposed(self, msg):
        if not self.decomposed:
            raise Exception(msg)

    # ---- matricisations (views) and Gram matrices (MPS:583-630, MPO:721-777) -------------------------------
    def left_site_matricization(self, index):
        return self.left_matricization(self.sites[index], index)

    def right_site_matricization(self, index):
        return self.right_matricization(self.sites[index], index)

    def left_matricization(self, matrix=None, index=0):
        m = self.sites[index] if matrix is None else matrix
        rows = int(np.prod(self.shape[index][:-1]))
        return m.reshape(rows, -1)

    def right_matricization(self, matrix=None, index=0):
        m = self.sites[index] if matrix is None else matrix
        cols = int(np.prod(self.shape[index][1:]))
        return m.reshape(-1, cols)

    def tensoricization(self, matrix, index):
        return matrix.reshape(self.shape[index])

    def left_orthogonality(self, index):
        L = self.sites[index].reshape(-1, self.sites[index].shape[-1])
        return sw.gram(L, "left").cpu().numpy()

    def right_orthogonality(self, index):
        R = self.sites[index].reshape(self.sites[index].shape[0], -1)
        return sw.gram(R, "right").cpu().numpy()

    def grad(self, index):
        return self.sites[:index] + self.sites[index + 2:]

    def _repr(self, title):
        txt = "<%s> \n> Sites shape" % title + str(self.shape) + "\n"
        txt += "\t" + "|   " * self.sites_number + "\n"
        txt += "\t" + ("O---" * (self.sites_number - 1)) + "O" + "\n"
        return txt
