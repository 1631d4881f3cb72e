"""DMRG environment blocks -- reference: variational/dmrg.py:65-112, cores on the GPU.

Only the well-defined part of the reference's DMRG is provided: the left and right blocks of <S| W |S> (no conjugation, like the
reference), each a chain of three strided GEMMs per site.  The reference's `solve` / sweeps feed a contracted 4-tensor to an
experimental optimizer (dmrg.py:134-172) and are out of scope."""
from syngular.tensor import _sweeps as sw


class DMRG:
    @staticmethod
    def right_blocks(operator, state):
        """`DMRG.__right_blocks(operator, state)`: list of n entries, blocks (a, w, a') at k = 2 .. n-1 as CUDA tensors, None elsewhere."""
        return sw.dmrg_right_blocks(state.sites, operator.sites)

    @staticmethod
    def left_blocks(operator, state):
        """`DMRG.__left_blocks(operator, state)`: blocks (b, v, b') at k = 0 .. n-3, None elsewhere."""
        return sw.dmrg_left_blocks(state.sites, operator.sites)

    @staticmethod
    def solve(operator, optimizer=None):
        raise NotImplementedError("the reference's DMRG sweeps (variational/dmrg.py:16-63,134-172) are experimental and out of scope; "
                                  "the environment blocks are DMRG.left_blocks / DMRG.right_blocks")
