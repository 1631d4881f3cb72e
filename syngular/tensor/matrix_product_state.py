"""MatrixProductState -- drop-in for the reference's tensor/matrix_product_state.py, cores on the GPU.

Same constructor, attributes, operators (`+`, `|`, `>>`, `[]`) and methods; `sites[k]` are contiguous CUDA float64
torch tensors of shape (l_k, d_k, r_k) whose `.cpu().numpy()` equals the reference's arrays up to gauge.  Scalars and
inspection results (`|`, `dot`, `[]`, `to_tensor`, Gram matrices) come back as numpy values like the reference's.
"""
import numpy as np
import torch

from syngular.tensor import _sweeps as sw
from syngular.tensor._chain import _MatrixProduct


class MatrixProductState(_MatrixProduct):
    PHYS = 1

    def __init__(self, tensor=None, bond_shape=(), verbose=0):
        self._init_empty(verbose)
        if tensor is not None:                                   # MPS:35-58
            self.tensor = sw.as_core(tensor)
            self.tensor_shape = tuple(int(x) for x in self.tensor.shape)
            self.bond_shape = tuple(int(b) for b in bond_shape)
            self.order = len(self.tensor_shape)
            self.real_parameters_number = int(np.prod(self.tensor_shape))
            self.input_shape = self.tensor_shape
            self.sites_number = len(self.bond_shape) + 1
            self.sites = [None] * self.sites_number
            self.norm = None
            if self.order != self.sites_number:
                raise Exception("dimensions of bond indices do not match order - 1")
            self.shape = self._chain_shapes([(d,) for d in self.tensor_shape], self.bond_shape)

    # ---- operators --------------------------------------------------------------------------------------
    def __add__(self, mps):
        if self.decomposed and mps.decomposed:                   # MPS:75-102 (no truncation for states)
            n = self.sites_number
            sites = [sw.add_site(self.sites[k], mps.sites[k], k == 0, k == n - 1) for k in range(n)]
            return MatrixProductState.from_sites(sites)
        raise Exception("Both Matrix Product Operator must be in canonical form (use .decompose()")

    def __or__(self, mp):
        if isinstance(mp, MatrixProductState):                   # MPS:116-129, bilinear, no conjugation
            v = sw.overlap(self.sites, mp.sites).item()
            return np.complex128(v) if isinstance(v, complex) else np.float64(v)
        raise Exception("right-hand site must be a MatrixProductState")

    def overlap_device(self, mp):
        """`self | mp` without the host synchronisation: a (1,1) CUDA tensor."""
        return sw.overlap(self.sites, mp.sites)

    def __getitem__(self, key):
        if len(key) != self.sites_number:
            raise Exception("input indices do not match the number of sites")
        return self.retrieve(key)

    def __repr__(self):
        return self._repr("Matrix Product State")

    # ---- constructors -----------------------------------------------------------------------------------
    @staticmethod
    def random(input_shape, bond_shape):
        tensor = np.random.normal(size=input_shape)              # MPS:176-178 (host RNG, like the reference)
        return MatrixProductState(tensor, bond_shape=bond_shape).decompose()

    @staticmethod
    def random_cores(input_shape, bond_shape, seed=None):
        """Direct core generator for chains too long to draw densely (SURVEY section 9 `random`): core k ~ N(0, 1/(l d))."""
        rng = np.random.default_rng(seed)
        shapes = _MatrixProduct._chain_shapes([(d,) for d in input_shape], bond_shape)
        return MatrixProductState.from_sites([rng.normal(size=s) / np.sqrt(s[0] * s[1]) for s in shapes])

    @staticmethod
    def empty():
        return MatrixProductState()

    @staticmethod
    def zeros(input_shape, bond_shape):
        shapes = _MatrixProduct._chain_shapes([(d,) for d in input_shape], bond_shape)
        dev = sw.device()
        return MatrixProductState.from_sites([torch.zeros(s, dtype=torch.float64, device=dev) for s in shapes])

    # ---- scalars ----------------------------------------------------------------------------------------
    def dot(self):
        return np.sqrt(self | self)

    def normalize(self):
        self.sites[-1] = sw.normalize_last(self.sites[-1])       # MPS:252-256 divides the last core by its Frobenius norm
        return self

    def conj(self):
        """Complex conjugate chain (extension: the reference's `|` is bilinear, so <psi|psi> is `psi.conj() | psi`)."""
        return MatrixProductState.from_sites([s.conj().resolve_conj() if s.is_complex() else s for s in self.sites])

    def to_tensor(self):
        return sw.to_dense(self.sites).cpu().numpy().astype(complex)     # the reference returns a complex array (MPS:266)

    def to_tensor_device(self):
        return sw.to_dense(self.sites)

    def retrieve(self, indices):
        return sw.retrieve(self.sites, indices).cpu().numpy()

    # ---- decomposition ----------------------------------------------------------------------------------
    def decompose(self, mode="left"):
        if not self.decomposed:
            if mode == "left":
                self.sites = sw.decompose_left(self.tensor, self.shape)
            elif mode == "right":
                self.sites = sw.decompose_right(self.tensor, self.shape)
            else:
                return self
            self.parameters_number = int(sum(int(np.prod(s.shape)) for s in self.sites))
            self.shape = [tuple(int(x) for x in s.shape) for s in self.sites]
            del self.tensor
            self.decomposed = True
        return self

    def apply(self, operator, index, strict=True, mode="compress", chi_max=None, cutoff=0.0):
        """Apply a dense m-site gate (ndarray with legs out_0..out_{m-1}, in_0..in_{m-1}) at sites index..index+m-1 (MPS:487-534).
        mode="compress": the reference's behaviour -- QR re-split keeping the existing bonds (never grows).
        mode="svd"     : extension -- local SVD split, bonds grow up to chi_max (relative singular-value cutoff)."""
        gate = sw.as_core(operator)
        m = gate.dim() // 2
        if gate.dim() != 2 * m or index < 0 or index + m > self.sites_number:
            raise Exception("gate does not fit the chain at this index")
        that = self.copy() if strict else self
        that.sites = sw.apply_gate(that.sites, gate, index, mode=mode, chi_max=chi_max, cutoff=cutoff)
        that._refresh_from_cores(bonds=(mode == "svd"))
        return that
