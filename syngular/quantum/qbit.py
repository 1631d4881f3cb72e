"""Qbit -- a qubit register on a matrix-product state (reference: quantum/qbit.py:11-150), cores on the GPU.

Like the reference, the register starts in |0...0> with every bond fixed at 2 and gates are applied with
`MatrixProductState.apply` (QR re-split, the bond never grows -- entanglement beyond chi=2 is projected away).  Pass
`chi_max` to switch to the growing-bond SVD split."""
import numpy as np

from syngular.quantum import gate
from syngular.tensor.matrix_product_state import MatrixProductState


class Qbit:
    VERBOSE = 0
    LSB = True

    def __init__(self, size, init=True, chi_max=None, cutoff=0.0):
        self.size = size
        self.dim = 2 ** size
        self.state = None
        self.chi_max, self.cutoff = chi_max, cutoff
        if init:
            b = [1] + [2] * (size - 1) + [1]
            cores = []
            for k in range(size):
                c = np.zeros((b[k], 2, b[k + 1]))
                if k < size - 1:
                    c[0] = gate.I[:, : b[k + 1]]                    # qbit.py:26-27
                else:
                    c[0, 0, 0] = 1.0                                # qbit.py:28
                cores.append(c)
            self.state = MatrixProductState.from_sites(cores, real_parameters_number=self.dim)

    @staticmethod
    def from_mps(mps, chi_max=None, cutoff=0.0):
        q = Qbit(mps.sites_number, init=False, chi_max=chi_max, cutoff=cutoff)
        q.state = mps
        return q

    @staticmethod
    def from_binary(bits):
        q = Qbit(len(bits))
        for i, ch in enumerate(bits):
            if ch == "1":
                q @= (gate.X, i)
        return q

    def _apply(self, g, index):
        if self.chi_max is not None:
            new = self.state.apply(g, index, mode="svd", chi_max=self.chi_max, cutoff=self.cutoff)
        else:
            new = self.state.apply(g, index)
        return Qbit.from_mps(new, self.chi_max, self.cutoff)

    def __matmul__(self, operator):
        if isinstance(operator, tuple):
            if len(operator) == 2:                                  # (gate, index): qbit.py:34-35
                return self._apply(*operator)
            g, a, b = operator                                      # (gate, control, target): qbit.py:36-60
            lo, hi = min(a, b), max(a, b)
            q = self
            if hi - lo != 1:
                q = q.swap_in(lo, hi - 1)
                if lo != a:
                    q = q @ (gate.SWAP, hi - 1)
            q = q._apply(g, hi - 1)
            if hi - lo != 1:
                q = q.swap_out(lo, hi - 1)
            return q
        if self.size == 1:                                          # bare matrix: only meaningful on one qubit
            return self._apply(np.asarray(operator), 0)
        raise Exception("bare-matrix gates on a multi-qubit register are undefined in the reference (qbit.py:62-63 crashes)")

    def __imatmul__(self, operator):
        return self @ operator

    def swap(self, idx1, idx2):
        return self.swap_in(idx1, idx2).swap_out(idx1, idx2)

    def swap_in(self, idx1, idx2):
        q = self
        for i in range(min(idx1, idx2), max(idx1, idx2)):
            q = q @ (gate.SWAP, i)
        return q

    def swap_out(self, idx1, idx2):
        q = self
        for i in range(max(idx1, idx2), min(idx1, idx2) - 1, -1):
            if i + 2 > q.size:
                continue                                            # the reference indexes past the chain here
            q = q @ (gate.SWAP, i)
        return q

    def apply(self, g):
        return self @ g

    def to_tensor(self):
        return self.state.to_tensor().reshape(self.dim)

    def to_binary(self):
        tensor = np.rint(self.to_tensor().real).astype(int)
        s = bin(int(np.where(tensor == 1)[0][0]))[2:].zfill(self.size)
        return s[::-1] if Qbit.LSB else s
