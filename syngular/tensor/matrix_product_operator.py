"""MatrixProductOperator -- drop-in for the reference's tensor/matrix_product_operator.py, cores on the GPU.

Cores are contiguous CUDA float64 tensors (l, in, out, r).  `@` (MPO x MPS, MPO x MPO), `+`, `*`, `>>`, canonical forms
and element access keep the reference's semantics, including the `>> min_bond` that ends `@`, `+`, `*` with min_bond taken
from the (possibly stale) bond_shape METADATA (MPO:80,128,176).
"""
import numpy as np
import torch

from syngular.tensor import _sweeps as sw
from syngular.tensor._chain import _MatrixProduct
from syngular.tensor.matrix_product_state import MatrixProductState


class MatrixProductOperator(_MatrixProduct):
    PHYS = 2

    COMPRESS = True          # class flags of the reference (MPO:20-23); only MATMUL_MODE is read
    DECOMPOSE = False
    MATMUL_MODE = "standard"  # "standard" (reference) | "fused" (same numbers, product cores never materialised)

    def __init__(self, tensor=None, bond_shape=(), verbose=0):
        self._init_empty(verbose)
        if tensor is not None:                                   # MPO:30-60
            t = sw.as_core(tensor)
            self.tensor_shape = tuple(int(x) for x in t.shape)
            self.bond_shape = tuple(int(b) for b in bond_shape)
            self.real_parameters_number = int(np.prod(self.tensor_shape))
            self.sites_number = n = len(self.bond_shape) + 1
            self.sites = [None] * n
            self.input_shape = tuple(self.tensor_shape[:n])
            self.output_shape = tuple(self.tensor_shape[n:])
            if len(self.input_shape) != len(self.output_shape):
                raise Exception("input_shape and output_shape of the tensor must have the same length")
            if len(self.input_shape) != len(self.bond_shape) + 1:
                raise Exception("dimensions of bond indices do not match input dimension - 1")
            self.shape = self._chain_shapes(list(zip(self.input_shape, self.output_shape)), self.bond_shape)
            self.shape_indices = [(2 * n + i, i, n + i, 2 * n + i + 1) for i in range(n)]
            # interleave (in_0, out_0, in_1, out_1, ...): MPO:60
            self.tensor = t.permute(sum(zip(range(n), range(n, 2 * n)), ())).contiguous()

    # ---- operators --------------------------------------------------------------------------------------
    def __add__(self, mpo):
        min_bond = min(min(self.bond_shape), min(mpo.bond_shape))            # MPO:80 (metadata)
        if self.decomposed and mpo.decomposed:
            n = self.sites_number
            sites = [sw.add_site(self.sites[k], mpo.sites[k], k == 0, k == n - 1) for k in range(n)]
            return MatrixProductOperator.from_sites(sites) >> min_bond      # MPO:110
        raise Exception("Both Matrix Product Operator must be in canonical form (use .decompose()")

    def __mul__(self, mp):
        min_bond = min(min(self.bond_shape), min(mp.bond_shape))             # MPO:128
        if not self.decomposed and mp.decomposed:
            raise Exception("Both Matrix Product Operator must be in canonical form (use .decompose()")
        if isinstance(mp, MatrixProductOperator):
            sites = [sw.ops.kron_site(a, b) for a, b in zip(self.sites, mp.sites)]
            return MatrixProductOperator.from_sites(sites) >> min_bond      # MPO:156
        if isinstance(mp, MatrixProductState):
            raise NotImplementedError("MPO * MPS is an empty stub in the reference (MPO:158-159)")
        raise Exception("left hand-side must be either a MatrixProductState or a MatrixProductOperator")

    def __matmul__(self, mp):
        min_bond = min(min(self.bond_shape), min(mp.bond_shape))             # MPO:176 (metadata)
        if isinstance(mp, MatrixProductState):
            if MatrixProductOperator.MATMUL_MODE in ("standard", "fused"):
                return _apply_to_state(self, mp, min_bond)
            raise NotImplementedError("MATMUL_MODE %r is unfinished in the reference (MPO:193-276)" % MatrixProductOperator.MATMUL_MODE)
        if isinstance(mp, MatrixProductOperator):
            sites = [sw.site_mpo_mpo(a, b) for a, b in zip(self.sites, mp.sites)]       # MPO:278-288
            return MatrixProductOperator.from_sites(sites) >> min_bond                  # MPO:289
        return None

    def __mod__(self, mode):
        raise NotImplementedError("`%` is an empty stub in the reference (MPO:291-315)")

    def __getitem__(self, key):
        key_inp, key_out = key[0], key[1]
        if len(key_inp) != self.sites_number:
            raise Exception("input indices do not match the number of sites")
        if len(key_out) != self.sites_number:
            raise Exception("output indices do not match the number of sites")
        return self.retrieve(key_inp, key_out)

    def __repr__(self):
        return self._repr("Matrix Product Operator") + "\t" + "|   " * self.sites_number + "\n"

    # ---- constructors -----------------------------------------------------------------------------------
    @staticmethod
    def random(input_shape, output_shape, bond_shape):
        tensor = np.random.normal(size=(*input_shape, *output_shape))        # MPO:353-355
        return MatrixProductOperator(tensor, bond_shape=bond_shape).decompose()

    @staticmethod
    def random_cores(input_shape, output_shape, bond_shape, seed=None):
        rng = np.random.default_rng(seed)
        shapes = _MatrixProduct._chain_shapes(list(zip(input_shape, output_shape)), bond_shape)
        return MatrixProductOperator.from_sites([rng.normal(size=s) / np.sqrt(s[0] * s[1]) for s in shapes])

    @staticmethod
    def empty():
        return MatrixProductOperator()

    @staticmethod
    def zeros(input_shape, output_shape, bond_shape):
        shapes = _MatrixProduct._chain_shapes(list(zip(input_shape, output_shape)), bond_shape)
        dev = sw.device()
        return MatrixProductOperator.from_sites([torch.zeros(s, dtype=torch.float64, device=dev) for s in shapes])

    # ---- dense <-> chain --------------------------------------------------------------------------------
    def to_tensor(self):
        return sw.to_dense(self.sites).cpu().numpy()

    def to_tensor_device(self):
        return sw.to_dense(self.sites)

    def retrieve(self, input_indices, output_indices):
        return sw.retrieve(self.sites, input_indices, output_indices).cpu().numpy()

    def decompose(self, mode="left"):
        if self.bond_shape == () and not self.decomposed:                    # MPO:419-425
            self.sites = [self.tensor.reshape(1, self.tensor_shape[0], self.tensor_shape[1], 1).contiguous()]
            self.decomposed = True
            del self.tensor
            return self
        if not self.decomposed:
            if mode != "left":
                raise NotImplementedError("MatrixProductOperator.decompose(mode='right') is wrong in the reference (MPO:455-463)")
            self.sites = sw.decompose_left(self.tensor, self.shape)
            self.parameters_number = int(sum(int(np.prod(s.shape)) for s in self.sites))
            self.shape = [tuple(int(x) for x in s.shape) for s in self.sites]
            del self.tensor
            self.decomposed = True
        return self

    def apply(self, operator, indices):
        """Apply an MPO `operator` to the sites `indices` of this chain, in place, returns None (MPO:582-626).  One site: the
        exact product.  Several sites: the reference's literal re-split (see _sweeps.mpo_apply_range)."""
        if not self.decomposed:
            raise Exception("MatrixProductState not decomposed")                     # MPO:650 (message as in the reference)
        if not isinstance(operator, MatrixProductOperator):
            raise NotImplementedError("MatrixProductOperator.apply with a dense operator is unfinished in the reference (MPO:627-647)")
        self.sites = sw.mpo_apply_range(self.sites, operator.sites, [int(i) for i in indices])
        self._refresh_from_cores(bonds=False)
        return None

    def transpose(self):
        return MatrixProductOperator.from_sites([s.permute(0, 2, 1, 3).contiguous() for s in self.sites])

    @staticmethod
    def split():
        raise NotImplementedError("empty stub in the reference (MPO:653-655)")


def _apply_to_state(W, X, min_bond, rounding=None, guard=True, host_out=None):
    """MPO x MPS then `>> min_bond` (MPO:181-192).  The guard of `>>` is evaluated on the bonds the product WOULD have
    (from_sites metadata: products of the actual bonds), exactly like the reference's from_sites + compress."""
    prod_bonds = tuple(x.shape[2] * w.shape[3] for x, w in zip(X.sites[:-1], W.sites[:-1]))
    rounding = rounding or MatrixProductState.ROUNDING
    guard_noop = guard and min_bond >= min(prod_bonds)      # explicit `bond=` targets (syn.mul extension) skip the guard
    if sw.any_complex(W.sites, X.sites) and not guard_noop:
        # complex128 cores: the literal route on the planar compositions (the fused sweeps are real-only)
        prod = MatrixProductState.from_sites([sw.site_mpo_mps(x, w) for x, w in zip(X.sites, W.sites)])
        if rounding == "svd":
            prod.sites, prod.truncation = sw.round_svd(prod.sites, min_bond, MatrixProductState.SVD_CUTOFF)
            prod._refresh_from_cores(bonds=False)
            return prod
        return prod >> min_bond
    if guard_noop or (MatrixProductOperator.MATMUL_MODE == "standard" and rounding == "qr" and not _FUSE_STANDARD):
        sites = [sw.site_mpo_mps(x, w) for x, w in zip(X.sites, W.sites)]
        return MatrixProductState.from_sites(sites) >> min_bond
    out = MatrixProductState()
    out.sites_number = X.sites_number
    out.decomposed = True
    out.orthonormalized = None
    out.real_parameters_number = None
    if rounding == "svd":
        out.sites, out.truncation = sw.apply_round_dm(X.sites, W.sites, min_bond, MatrixProductState.SVD_CUTOFF, host_out=host_out)
        out._streamed = host_out is not None          # the result cores were copied to the host from inside the sweep
    else:
        out.sites = sw.apply_round_qr(X.sites, W.sites, min_bond)
    out._refresh_from_cores(bonds=False)
    out.bond_shape = prod_bonds                       # stale pre-truncation bonds, as after from_sites(...) >> min_bond
    return out


# `@` always ends in `>> min_bond`; fusing the two is a pure re-association of the same GEMMs and projections, so it is on
# by default even in "standard" mode.  Set to False to force the literal materialise-then-round path (used by the tests to
# show both give the same numbers).
_FUSE_STANDARD = True
