"""GPU: gate application with the growing-bond SVD split (BASELINE configs[2] regime, real gates):
a brick-wall circuit of random orthogonal two-qubit gates against an exact dense state-vector simulation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_orthogonal(rng, n):
    q, r = np.linalg.qr(rng.normal(size=(n, n)))
    return q * np.sign(np.diag(r))


def dense_apply(psi, g4, i, n):
    psi = psi.reshape((2,) * n)
    psi = np.tensordot(g4.reshape(2, 2, 2, 2), psi, axes=([2, 3], [i, i + 1]))
    return np.moveaxis(psi, [0, 1], [i, i + 1]).reshape(-1)


def test_brickwall_circuit_exact_when_bond_is_large_enough():
    from syngular.quantum import Circuit
    rng = np.random.default_rng(3)
    n, depth = 8, 6
    structure, psi = [], np.zeros(2 ** n)
    psi[0] = 1.0
    for layer in range(depth):
        for i in range(layer % 2, n - 1, 2):
            g = random_orthogonal(rng, 4)
            structure.append((g.reshape(2, 2, 2, 2), i))
            psi = dense_apply(psi, g, i, n)
    c = Circuit(n, structure=structure, chi_max=16)
    c.run()
    got = c.get().to_tensor().real
    assert np.max(np.abs(got - psi)) < 1e-11
    bonds = [s.shape[2] for s in c.get().state.sites[:-1]]
    assert max(bonds) == 16 and bonds == [2, 4, 8, 16, 8, 4, 2]
    # truncated run: norm can only decrease, and the kept state stays close for a mild cut
    c2 = Circuit(n, structure=structure, chi_max=8)
    c2.run()
    got2 = c2.get().to_tensor().real
    assert np.linalg.norm(got2) <= 1.0 + 1e-12 and abs(np.dot(got2, psi)) > 0.5


def test_reference_mode_freezes_the_bond():
    from syngular.quantum import Qbit, gate
    q = Qbit(4)
    for g in ((gate.H, 0), (gate.CX, 0, 1), (gate.H, 2), (gate.CX, 2, 3), (gate.CX, 1, 2)):
        q @= g
    assert [s.shape[2] for s in q.state.sites[:-1]] == [2, 2, 2]          # qbit.py:23 / MPS:516: never grows
    y = (Qbit(2) @ (gate.Y, 0)).to_tensor()                               # complex gates run planar on the same kernels
    assert np.allclose(y, [0, 0, 1j, 0], atol=1e-14)
