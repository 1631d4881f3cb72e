"""GPU: the drop-in `syngular.tensor` API (sm_100a kernels through the C ABI) on the reference's own scenarios, checked
against the goldens produced by the unmodified reference AND (same cases) the numpy oracle."""
import pytest

import golden_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases.ALL_CASES, ids=lambda c: c.__name__)
def test_cuda_matches_reference_golden(case):
    from backends import ProductBackend
    case(ProductBackend())


@pytest.mark.parametrize("case", [golden_cases.case_matmul_known_answers, golden_cases.case_random_chain_r1,
                                  golden_cases.case_random_chain_r2, golden_cases.case_readme_chain], ids=lambda c: c.__name__)
def test_cuda_unfused_path_matches_too(case):
    """The literal materialise-then-round `@` must give the same numbers as the fused default."""
    from backends import ProductBackend
    import syngular.tensor.matrix_product_operator as M
    old = M._FUSE_STANDARD
    M._FUSE_STANDARD = False
    try:
        case(ProductBackend())
    finally:
        M._FUSE_STANDARD = old
