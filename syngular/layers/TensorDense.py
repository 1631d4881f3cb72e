"""TensorDense -- MPO-compressed dense layer, forward pass (reference: layers/TensorDense.py:17-142, a Keras layer).

Same constructor arguments and core layouts as the reference -- first core (i0, o0, b0), middle cores (i, o, b_l, b_r), last core
(i, o, b_last), bias of shape tt_output_shape, activation (default relu) -- but the weights are CUDA tensors and the forward
pass is a chain of N strided DMMA GEMMs (one per core, no transposition, the whole batch in the GEMM's N dimension) followed
by a fused bias + activation kernel.  The reference contracts sample by sample under tf.vectorized_map, and its einsum
specification is only valid for exactly 3 cores (TensorDense.py:110-114); this works for any number of cores.

Two arithmetic paths:
  precision="f64"  (default, the checked path): FP64 on the strided DMMA GEMM, any number of cores and any shapes;
  precision="tf32": float32 in / out, products on tcgen05.mma.kind::tf32 with FP32 accumulation -- the reference computes in float32
                    (Keras default dtype; TensorDense.py:50-71 add_weight) -- by ONE fused kernel per call that keeps both per-sample
                    intermediates in TMEM / shared memory (csrc/ttdense.cu).  Covered shape: three cores with every mode and bond 16
                    (BASELINE configs[4]); anything else raises.  Stated tolerance against the float64 restatement: 4e-3 of the output's
                    largest magnitude (TF32 keeps 10 mantissa bits of every operand, three chained contractions).
"""
import numpy as np
import torch

from syngular.tensor import _sweeps as sw
from syngular_b200 import ops


class TensorDense:
    def __init__(self, tt_input_shape, tt_output_shape, tt_bond_shape, activation="relu", use_bias=True, seed=None, precision="f64"):
        self.tt_input_shape = tuple(int(x) for x in tt_input_shape)
        self.tt_output_shape = tuple(int(x) for x in tt_output_shape)
        self.tt_bond_shape = tuple(int(x) for x in tt_bond_shape)
        n = len(self.tt_output_shape)
        if len(self.tt_input_shape) != n or n != len(self.tt_bond_shape) + 1:
            raise Exception("Incompatible shapes. Cannot create TensorDense with %d %d and %d " % (
                len(self.tt_input_shape), n, len(self.tt_bond_shape)))
        self.cores_number = n
        self.tt_input_shape_unfold = int(np.prod(self.tt_input_shape))
        self.tt_output_shape_unfold = int(np.prod(self.tt_output_shape))
        self.activation = activation
        self.use_bias = use_bias
        self.cores, self.bias = [], None
        self._seed = seed
        if precision not in ("f64", "tf32"):
            raise ValueError("precision must be 'f64' or 'tf32'")
        if precision == "tf32" and not ops.tt_dense3_fits(self.tt_input_shape, self.tt_output_shape, self.tt_bond_shape):
            raise NotImplementedError("the fused TF32 kernel covers three cores with every mode and bond equal to 16; use precision='f64'")
        if precision == "tf32" and activation not in ("relu", None, "linear", "identity"):
            raise NotImplementedError("the fused TF32 kernel applies relu or no activation")
        self.precision = precision
        self._packed = self._bias32 = None

    def core_shapes(self):
        n, i, o, b = self.cores_number, self.tt_input_shape, self.tt_output_shape, self.tt_bond_shape
        if n == 1:
            return [(i[0], o[0], 1)]
        return [(i[0], o[0], b[0])] + [(i[k], o[k], b[k - 1], b[k]) for k in range(1, n - 1)] + [(i[n - 1], o[n - 1], b[n - 2])]

    def build(self, cores=None, bias=None):
        """Create the weights (Keras "random_normal": N(0, 0.05^2); bias zeros) or adopt given arrays (reference layouts)."""
        if cores is None:
            rng = np.random.default_rng(self._seed)
            cores = [rng.normal(scale=0.05, size=s) for s in self.core_shapes()]
        assert [tuple(c.shape) for c in cores] == self.core_shapes(), "cores must be in the reference's layouts"
        self.cores = [sw.as_core(c) for c in cores]
        self.bias = sw.as_core(bias if bias is not None else np.zeros(self.tt_output_shape)) if self.use_bias else None
        if self.precision == "tf32":
            # the reference's weights are float32 (Keras default): round once, pack into the kernel's operand images
            g32 = [c.to(torch.float32).contiguous() for c in self.cores]
            self._packed = ops.tt_dense3_pack(*g32)
            self._bias32 = self.bias.to(torch.float32).reshape(-1).contiguous() if self.bias is not None else None
        return self

    def _dims(self, k):
        """(I, O, BL, BR) of core k."""
        n, b = self.cores_number, self.tt_bond_shape
        return self.tt_input_shape[k], self.tt_output_shape[k], (1 if k == 0 else b[k - 1]), (1 if k == n - 1 else b[k])

    def __call__(self, inputs, chunk=None):
        return self.call(inputs, chunk)

    def call(self, inputs, chunk=None):
        if not self.cores:
            self.build()
        if self.precision == "tf32":
            x32 = inputs if isinstance(inputs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(inputs, dtype=np.float32))
            x32 = x32.to(device=sw.device(), dtype=torch.float32).reshape(-1, self.tt_input_shape_unfold).contiguous()
            return ops.tt_dense3_tf32(x32, self._packed, self._bias32, relu=self.activation == "relu")
        x = sw.as_core(inputs).reshape(-1, self.tt_input_shape_unfold)
        batch = x.shape[0]
        out = torch.empty((batch, self.tt_output_shape_unfold), dtype=torch.float64, device=x.device)
        if chunk is None:                       # keep the largest intermediate around 2 GB
            widest = max(int(np.prod(self.tt_output_shape[:k + 1])) * self._dims(k)[3] * int(np.prod(self.tt_input_shape[k + 1:]))
                         for k in range(self.cores_number))
            chunk = max(1, min(batch, (1 << 28) // max(widest, 1)))
        for lo in range(0, batch, chunk):
            self._forward_chunk(x[lo:lo + chunk], out[lo:lo + chunk])
        return out

    def _forward_chunk(self, x, out):
        n = self.cores_number
        T, g = x, x.shape[0]                                        # T viewed as [g, (b_l, i_k), R_k]
        for k in range(n):
            I, O, BL, BR = self._dims(k)
            R = int(np.prod(self.tt_input_shape[k + 1:])) if k + 1 < n else 1
            M, K = O * BR, BL * I
            dst = out if k == n - 1 else torch.empty((g * M * R,), dtype=torch.float64, device=x.device)
            # C[(o,br), (g,rho)] = sum_{(bl,i)} G_k[i, o, bl, br] * T[g, (bl,i), rho]   -- the core is read in place
            ops.gemm(self.cores[k], T, dst, M=M, N=g * R, K=K,
                     a_m=(BL * BR, 1, BR), a_k=(BR, O * BL * BR, I),
                     b_k=R, b_n=(K * R, 1, R),
                     c_m=R, c_n=(M * R, 1, R))
            T, g = dst, g * O                                       # [(g, o), (br, i_{k+1}), R_{k+1}]
        ops.bias_act_(out, self.bias.reshape(-1) if self.bias is not None else None, self.activation)
        return out
