"""CPU: host logic of the complex128 path (syngular_b200/cplx.py and the complex-aware sweeps) against numpy, with the kernel
wrappers swapped for their numpy restatements (tests/cpu_ops.py).  This checks the ORCHESTRATION -- planar 4-GEMM complex
products with the real descriptor, the interleaved embedding behind the complex qrt, the pairing of the doubled Jacobi spectrum
(including degenerate singular values) -- not the kernels; tests/test_gpu_complex.py runs the same scenarios on the B200."""
import numpy as np
import pytest
import torch

import cpu_ops
import complex_cases as cc


@pytest.fixture(autouse=True)
def _emulated_kernels():
    with cpu_ops.patched():
        yield


def test_emulated_gemm_follows_the_two_level_index_contract():
    # the emulation itself: the site contraction issued by the product equals the oracle's einsum
    from oracle import ref_numpy as R
    from syngular.tensor import _sweeps as sw
    rng = np.random.default_rng(0)
    X, W = rng.normal(size=(3, 2, 4)), rng.normal(size=(2, 2, 3, 5))
    got = sw.site_mpo_mps(torch.from_numpy(X), torch.from_numpy(W)).numpy()
    assert np.allclose(got, R.site_mpo_mps(X, W), atol=1e-13)


@pytest.mark.parametrize("name", sorted(cc.CASES))
def test_complex_case(name):
    cc.CASES[name]()


def test_real_golden_cases_through_the_emulated_kernels():
    """The reference-pinned golden scenarios (tests/golden_cases.py) drive the same host code on the CPU: guards the refactor
    that made the sweeps operand-generic (real tensors / planar complex)."""
    import backends
    import golden_cases
    be = backends.ProductBackend()
    for case in golden_cases.ALL_CASES:
        case(be)
