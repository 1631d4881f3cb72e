"""Drop-in `syngular` package: the reference's `syngular.tensor` API (MatrixProductState, MatrixProductOperator, syn.mul)
backed by the sm_100a library `syngular_b200`.  Mirrors /__init__.py:1 of the reference (`from syngular.tensor.utils import mul`)."""
from syngular.tensor.utils import mul  # noqa: F401
from syngular import tensor  # noqa: F401
