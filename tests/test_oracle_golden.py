"""CPU: the numpy oracle (oracle/ref_numpy.py) against fixtures produced by the unmodified reference."""
import pytest

from backends import OracleBackend
import golden_cases


@pytest.mark.parametrize("case", golden_cases.ALL_CASES, ids=lambda c: c.__name__)
def test_oracle_matches_reference_golden(case):
    case(OracleBackend())
