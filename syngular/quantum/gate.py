"""Gate tensors (reference: quantum/gate.py:4-61).  Multi-qubit gates are stored with legs (out..., in...).
Real gates run on the FP64 kernels directly; Y, T, S (and any complex unitary) make the touched cores complex128 and run
planar on the same kernels (syngular_b200/cplx.py)."""
import numpy as np

I = np.array([[1., 0.], [0., 1.]])
X = np.array([[0., 1.], [1., 0.]])
Y = np.array([[0., -1.j], [1.j, 0.]])
Z = np.array([[1., 0.], [0., -1.]])
H = 1 / np.sqrt(2) * np.array([[1., 1.], [1., -1.]])
T = np.array([[1., 0.], [0., np.exp(1.j * np.pi / 4)]])
S = np.array([[1., 0.], [0., np.exp(1.j * np.pi / 2)]])


def _perm(n, mapping):
    m = np.zeros((2 ** n, 2 ** n))
    for src in range(2 ** n):
        m[mapping(src), src] = 1.0
    return m.reshape((2,) * (2 * n))


CX = _perm(2, lambda b: b ^ 1 if b & 2 else b)                 # control = first qubit (most significant)
SWAP = _perm(2, lambda b: ((b & 1) << 1) | (b >> 1))
TOFFOLI = _perm(3, lambda b: b ^ 1 if (b & 6) == 6 else b)
CZ = np.diag([1., 1., 1., -1.]).reshape(2, 2, 2, 2)
