"""`syn.mul` -- functional alias of `@` / `|` (reference: tensor/utils.py:8-99)."""
from syngular.tensor import _sweeps as sw
from syngular.tensor.matrix_product_operator import MatrixProductOperator, _apply_to_state
from syngular.tensor.matrix_product_state import MatrixProductState


def mul(op1, op2, mode="standard", bond=None, host_out=None):
    """mode="standard": the reference's path (contraction + `>> min_bond`, QR truncation unless set_rounding("svd")).
    mode="optimized": the density-matrix algorithm the reference sketched but never finished (MPO:193-260, utils.py:84-91):
    MPO x MPS with optimal (SVD) truncation, product cores never formed.
    `bond` (extension): target bond dimension; defaults to the reference's min_bond rule.
    `host_out` (extension, MPO x MPS only): one pinned 1-D float64 host tensor per site; the result cores are also copied into them --
    from inside the sweep, overlapped with it, on the optimized route (a host-resident caller gets its arrays back without a
    separate download phase); the returned chain still holds the device cores."""
    res = _mul(op1, op2, mode, bond, host_out)
    if host_out is not None:
        if not isinstance(res, MatrixProductState) or any(not isinstance(c, sw.torch.Tensor) or c.is_complex() for c in res.sites):
            raise Exception("host_out needs a real MPO x MPS product")
        if not getattr(res, "_streamed", False):
            sw.copy_to_host(res.sites, host_out)
    return res


def _mul(op1, op2, mode, bond, host_out):
    n, m = op1.sites_number, op2.sites_number
    min_bond = min(min(op1.bond_shape), min(op2.bond_shape))                 # utils.py:12
    guard = bond is None
    if bond is not None:
        if not isinstance(bond, int):
            raise Exception("dimension should be an integer")
        min_bond = bond
    if n != m:
        raise Exception("both operator do not have the same number of sites")
    if mode == "optimized":
        if isinstance(op1, MatrixProductState) and isinstance(op2, MatrixProductOperator):
            op1, op2 = op2, op1
        if not (isinstance(op1, MatrixProductOperator) and isinstance(op2, MatrixProductState)):
            raise Exception("`syn.mul` mode 'optimized' needs one MatrixProductOperator and one MatrixProductState")
        if not op1.decomposed or not op2.decomposed:
            raise Exception("Operators and States must be decomposed")
        return _apply_to_state(op1, op2, min_bond, rounding="svd", guard=guard, host_out=host_out)
    if mode == "standard":
        if isinstance(op1, MatrixProductState) and isinstance(op2, MatrixProductOperator):
            if not op1.decomposed or not op2.decomposed:
                raise Exception("Operators and States must be decomposed")
            return _apply_to_state(op2, op1, min_bond, guard=guard)          # utils.py:30-43
        if isinstance(op1, MatrixProductOperator) and isinstance(op2, MatrixProductState):
            if not op1.decomposed or not op2.decomposed:
                raise Exception("Operators and States must be decomposed")
            return _apply_to_state(op1, op2, min_bond, guard=guard)          # utils.py:44-57
        if isinstance(op1, MatrixProductState) and isinstance(op2, MatrixProductState):
            return op1 | op2                                                 # utils.py:58-59
        if isinstance(op1, MatrixProductOperator) and isinstance(op2, MatrixProductOperator):
            # utils.py:60-71: op2's OUTPUT is contracted with op1's INPUT (the reverse of `@`), op2's bond major
            sites = [sw.site_mpo_mpo(b, a) for a, b in zip(op1.sites, op2.sites)]
            return MatrixProductOperator.from_sites(sites) >> min_bond
        raise Exception("`syn.mul` should be provided MatrixProductState or MatrixProductOperator objects only")
    if mode in ("variational", "fitup"):
        return None                                                          # empty stubs in the reference (utils.py:76-99)
    return None
