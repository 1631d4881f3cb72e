"""GPU: complex128 cores and gates (BASELINE configs[2]) on the real FP64 kernels -- the scenarios of
tests/complex_cases.py through the CUDA product, plus a larger brick-wall circuit whose bond SVDs run the multi-CTA Jacobi."""
import numpy as np
import pytest

import complex_cases as cc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(cc.CASES))
def test_complex_case(name):
    cc.CASES[name]()


def test_brickwall_12_qubits_chi_64_against_state_vector():
    """12 qubits, depth 8, Haar-random gates: exact up to bond 64 (embedded eigenproblems up to 256 x 256: Cholesky + ring Jacobi)."""
    from syngular.quantum import Circuit
    rng = np.random.default_rng(11)
    n, depth = 12, 8
    structure, psi = [], np.zeros(2 ** n, dtype=complex)
    psi[0] = 1.0
    for layer in range(depth):
        for i in range(layer % 2, n - 1, 2):
            g = cc.haar(rng, 4).reshape(2, 2, 2, 2)
            structure.append((g, i))
            psi = cc._dense_apply(psi, g, i, n)
    c = Circuit(n, structure=structure, chi_max=64)
    c.run()
    st = c.get().state
    assert max(s.shape[2] for s in st.sites[:-1]) == 64
    assert np.max(np.abs(c.get().to_tensor() - psi)) < 1e-10
    assert abs((st.conj() | st) - 1.0) < 1e-10
    # truncated: chi 16 keeps the norm below one and the fidelity with the exact state high but not perfect
    c2 = Circuit(n, structure=structure, chi_max=16)
    c2.run()
    got = c2.get().to_tensor()
    f = abs(np.vdot(psi, got)) ** 2
    assert np.linalg.norm(got) <= 1.0 + 1e-12 and 0.05 < f < 1.0


def test_brickwall_16_qubits_truncated_to_chi_128_matches_numpy_tebd():
    """Saturated complex bonds (unfoldings 256 x 256, embedded 512 x 512) are cut by the fused spectral-projection kernel (complex-structured mode); the whole
    truncated evolution is compared with the same sequence of truncations done by numpy SVDs."""
    from syngular.quantum import Circuit
    from syngular_b200 import ops
    rng = np.random.default_rng(13)
    n, depth, chi = 16, 9, 128
    structure = [(cc.haar(rng, 4).reshape(2, 2, 2, 2), i) for layer in range(depth) for i in range(layer % 2, n - 1, 2)]
    calls = [0]
    orig, orig_c = ops.dominant_subspace, ops.dominant_subspace_c128

    def spy(*a, **k):
        calls[0] += 1
        return orig(*a, **k)

    def spy_c(*a, **k):
        calls[0] += 1
        return orig_c(*a, **k)
    ops.dominant_subspace, ops.dominant_subspace_c128 = spy, spy_c
    try:
        c = Circuit(n, structure=structure, chi_max=chi)
        c.run()
    finally:
        ops.dominant_subspace, ops.dominant_subspace_c128 = orig, orig_c
    st = c.get().state
    assert max(s.shape[2] for s in st.sites[:-1]) == chi and calls[0] >= 1
    ref = cc._tebd_numpy(n, structure, chi)
    got = c.get().to_tensor()
    assert np.max(np.abs(got - ref)) < 1e-9 * np.max(np.abs(ref))
    assert abs((st.conj() | st) - np.vdot(ref, ref)) < 1e-9


def test_identity_basis_at_untruncated_bonds_gives_the_same_state():
    """Complex bonds that keep their whole space take the identity basis instead of an eigen-solve (cplx.IDENTITY_MIN_M): the split is exact
    for any unitary and later truncations only see that unitary on the bond index, so the final (truncated) state is the same."""
    from syngular.quantum import Circuit
    from syngular_b200 import cplx
    rng = np.random.default_rng(21)
    n, depth, chi = 14, 10, 64
    structure = [(cc.haar(rng, 4).reshape(2, 2, 2, 2), i) for layer in range(depth) for i in range(layer % 2, n - 1, 2)]
    saved = cplx.IDENTITY_MIN_M
    try:
        cplx.IDENTITY_MIN_M = 16            # bonds of 16 rows and more: many identity splits before the truncating ones
        a = Circuit(n, structure=structure, chi_max=chi); a.run()
        cplx.IDENTITY_MIN_M = 0             # every split by an eigen-solve
        b = Circuit(n, structure=structure, chi_max=chi); b.run()
    finally:
        cplx.IDENTITY_MIN_M = saved
    ta, tb = a.get().to_tensor(), b.get().to_tensor()
    assert np.linalg.norm(tb) < 1.0 - 1e-6                                  # the run does truncate
    assert np.max(np.abs(ta - tb)) < 1e-10 * np.max(np.abs(tb))
