"""`syngular.tensor` -- same exports as the reference's tensor/__init__.py:1-5 for the hot-path types."""
from syngular.tensor.matrix_product_state import MatrixProductState
from syngular.tensor.matrix_product_operator import MatrixProductOperator
from syngular.tensor.differential_matrix_product_operator import DifferentialMatrixProductOperator


def set_rounding(mode, chi_cutoff=0.0):
    """Select what `>>` (and the `>> min_bond` inside `@`, `+`, `*`) does on both types:
    "qr"  -- the reference's QR truncation (default; bit-for-bit the reference's semantics),
    "svd" -- optimal SVD truncation to the same target bond (north_star), relative singular-value cutoff `chi_cutoff`."""
    if mode not in ("qr", "svd"):
        raise Exception("rounding mode should be 'qr' or 'svd'")
    MatrixProductState.ROUNDING = mode
    MatrixProductOperator.ROUNDING = mode
    MatrixProductState.SVD_CUTOFF = float(chi_cutoff)
    MatrixProductOperator.SVD_CUTOFF = float(chi_cutoff)


__all__ = ["MatrixProductState", "MatrixProductOperator", "DifferentialMatrixProductOperator", "set_rounding"]
