"""Adapters so the same golden cases run against the numpy oracle (CPU) and the CUDA product (GPU)."""
import numpy as np


class OracleBackend:
    name = "oracle"

    def __init__(self):
        from oracle import ref_numpy as R
        self.R = R

    def mps_dense(self, t, bonds, mode="left"):
        return self.R.MPS.dense(t, bonds, mode=mode)

    def mpo_dense(self, t, bonds):
        return self.R.MPO.dense(t, bonds)

    def mps_sites(self, sites):
        return self.R.MPS.from_sites([np.array(s) for s in sites])

    def mpo_sites(self, sites):
        return self.R.MPO.from_sites([np.array(s) for s in sites])

    def mps_zeros(self, inp, bonds):
        return self.R.MPS.zeros(inp, bonds)

    def mul(self, a, b):
        return self.R.mul(a, b)

    def apply_gate(self, cores, gname, index):
        out = self.R.mps_apply([np.array(c) for c in cores], getattr(self.R.Gates, gname), index)
        return np.real(self.R.to_dense(out)), [c.shape[2] for c in out[:-1]]

    def run_circuit(self, size, ops):
        q = self.R.Qbit(size)
        for op in ops:
            q = q @ ((getattr(self.R.Gates, op[0]),) + tuple(op[1:]))
        return np.real(q.to_tensor())

    def sites(self, obj):
        return [np.asarray(s) for s in obj.sites]

    def dense(self, obj):
        return np.real(self.R.to_dense(obj.sites))

    def dense_of(self, sites):
        return np.real(self.R.to_dense([np.asarray(s) for s in sites]))

    def arr(self, x):
        return np.asarray(x)

    def scalar(self, x):
        return float(np.real(np.asarray(x)).reshape(-1)[0])


class ProductBackend:
    """The drop-in `syngular.tensor` API backed by the sm_100a library (cores are torch CUDA tensors)."""
    name = "cuda"

    def __init__(self):
        import torch
        from syngular.tensor import MatrixProductOperator, MatrixProductState
        import syngular
        from oracle import ref_numpy as R
        self.torch, self.MPS, self.MPO, self.syn, self.R = torch, MatrixProductState, MatrixProductOperator, syngular, R

    def mps_dense(self, t, bonds, mode="left"):
        return self.MPS(np.asarray(t, dtype=np.float64), bond_shape=bonds).decompose(mode=mode)

    def mpo_dense(self, t, bonds):
        return self.MPO(np.asarray(t, dtype=np.float64), bond_shape=bonds).decompose()

    def mps_sites(self, sites):
        return self.MPS.from_sites([np.array(s) for s in sites])

    def mpo_sites(self, sites):
        return self.MPO.from_sites([np.array(s) for s in sites])

    def mps_zeros(self, inp, bonds):
        return self.MPS.zeros(inp, bonds)

    def mul(self, a, b):
        return self.syn.mul(a, b)

    def apply_gate(self, cores, gname, index):
        from syngular.quantum import gate
        Y = self.MPS.from_sites([np.array(c) for c in cores]).apply(getattr(gate, gname), index)
        return np.real(self.arr(Y.to_tensor())), [s.shape[2] for s in Y.sites[:-1]]

    def run_circuit(self, size, ops):
        from syngular.quantum import Qbit, gate
        q = Qbit(size)
        for op in ops:
            q @= (getattr(gate, op[0]),) + tuple(op[1:])
        return np.real(q.to_tensor())

    def sites(self, obj):
        return [s.detach().cpu().numpy() for s in obj.sites]

    def dense(self, obj):
        return np.real(self.arr(obj.to_tensor()))

    def dense_of(self, sites):
        return np.real(self.R.to_dense([np.asarray(s) for s in sites]))

    def arr(self, x):
        if isinstance(x, self.torch.Tensor):
            return x.detach().cpu().numpy()
        return np.asarray(x)

    def scalar(self, x):
        return float(np.real(self.arr(x)).reshape(-1)[0])
