"""DifferentialMatrixProductOperator -- reference: tensor/differential_matrix_product_operator.py:9-173, cores on the GPU.

`project(index, state)` differentiates Y = W X with respect to the merged pair of operator cores (index, index + 1): it returns that
merged core and what multiplies it -- the left and right "wings" of W @ X and the two state cores under the pair.  The wings are chains
of the `@` site contraction (one strided GEMM per site, product bonds flattened MPS-bond major like everywhere on this path) multiplied
up with one GEMM per site; the reference builds them with one opt_einsum.contract over 2(index) operands (:104-149).  `gradient` is an
empty stub in the reference (:11-12) and stays one."""
from syngular.tensor import _sweeps as sw
from syngular.tensor.matrix_product_operator import MatrixProductOperator
from syngular.tensor.matrix_product_state import MatrixProductState


class DifferentialMatrixProductOperator(MatrixProductOperator):

    @staticmethod
    def from_sites(sites, orthogonality=None, real_parameters_number=None):
        mp = MatrixProductOperator.from_sites(sites, orthogonality, real_parameters_number)
        mp.__class__ = DifferentialMatrixProductOperator
        return mp

    def gradient(self, loss):
        pass                                                       # reference :11-12

    def crumble_site(self, index):
        """Merged core (l, i, o, i', o', r') of sites index and index + 1 (:13-27): one GEMM over the shared bond."""
        return sw.crumble(self.sites[index], self.sites[index + 1])

    def project(self, index, state):
        """{'center_site', 'left_wing' (1, prod(out), a, w), 'left_center', 'right_center', 'right_wing' (a, w, prod(out), 1)} (:79-173)."""
        if not isinstance(state, MatrixProductState):
            raise Exception("projected wings should come from a matrix product state input")
        if not (0 <= index < self.sites_number - 1):
            raise Exception("trying to project on non-existant site (site indices should be between 0 and the number of sites - 1)")
        if index < 1 or index > self.sites_number - 3:
            # the reference hands opt_einsum an operand-free expression here and crashes inside it (:104-112, :146-154)
            raise Exception("project needs at least one site on each side of the merged pair (1 <= index <= n - 3)")
        left, right = sw.project_wings(state.sites, self.sites, index)
        return {"center_site": self.crumble_site(index), "left_wing": left, "left_center": state.sites[index],
                "right_center": state.sites[index + 1], "right_wing": right}
