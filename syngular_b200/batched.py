"""Batches of independent matrix-product states on one GPU (BASELINE configs[3]: 8192 random MPS, N=32, d=2, chi=64).

The reference has no batching construct on this path: a batch is a Python loop over `A | B` (MPS:116-129) or `W @ X`
(MPO:181-192).  Here core k of all B states is ONE (B, l, d, r) CUDA tensor and every step of the transfer-matrix / apply+round
sweeps is a single batched launch of the same kernels the single-chain path uses.  Across GPUs the batch is sharded in
contiguous blocks (syngular_b200.parallel) and only the B result scalars are gathered.
"""
import numpy as np
import torch

from . import ops

F64 = torch.float64


class BatchedMatrixProductState:
    def __init__(self, cores):
        self.sites = [c.contiguous() for c in cores]
        self.batch = int(self.sites[0].shape[0])
        self.sites_number = len(self.sites)
        self.shape = [tuple(int(x) for x in c.shape[1:]) for c in self.sites]
        self.bond_shape = tuple(s[-1] for s in self.shape[:-1])
        self.input_shape = tuple(s[1] for s in self.shape)

    # ---- constructors ---------------------------------------------------------------------------------------
    @staticmethod
    def from_states(states):
        """Stack single MatrixProductState objects (or lists of cores) of identical shapes."""
        lists = [s.sites if hasattr(s, "sites") else s for s in states]
        n = len(lists[0])
        cores = []
        for k in range(n):
            cores.append(torch.stack([torch.as_tensor(l[k]).to(F64) if not isinstance(l[k], torch.Tensor) else l[k] for l in lists]).cuda())
        return BatchedMatrixProductState(cores)

    @staticmethod
    def random(batch, input_shape, bond_shape, seed=0, device=None):
        """Cores ~ N(0, 1/(l d)) drawn on the device (large batches: 8192 x 1.35 MB); use from_states for bit-identical host data."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=device).manual_seed(int(seed))
        b = (1,) + tuple(int(x) for x in bond_shape) + (1,)
        cores = []
        for k, d in enumerate(input_shape):
            c = torch.randn((batch, b[k], d, b[k + 1]), dtype=F64, device=device, generator=g)
            cores.append(c / float(np.sqrt(b[k] * d)))
        return BatchedMatrixProductState(cores)

    def state(self, b):
        from syngular.tensor import MatrixProductState
        return MatrixProductState.from_sites([c[b] for c in self.sites])

    # ---- `|` for the whole batch ----------------------------------------------------------------------------------
    def overlap(self, other):
        """(B,) tensor of <self_b | other_b> (bilinear, like MPS:116-129): two batched strided GEMMs per site."""
        assert other.batch == self.batch and other.sites_number == self.sites_number
        B = self.batch
        E = torch.ones((B, 1, 1), dtype=F64, device=self.sites[0].device)
        for a, b in zip(self.sites, other.sites):
            la, ra, lb, rb = a.shape[1], a.shape[-1], b.shape[1], b.shape[-1]
            T = ops.matmul(E.transpose(1, 2), a.reshape(B, la, -1))                  # (B, lb, d*ra)
            E = ops.matmul(T.reshape(B, -1, ra).transpose(1, 2), b.reshape(B, -1, rb))  # (B, ra, rb)
        return E.reshape(B)

    def norms2(self):
        return self.overlap(self)

    # ---- MPO application + rounding for the whole batch (shared operator) ---------------------------------------------
    def apply_round(self, mpo, dim):
        """`W @ X_b` followed by `>> dim` (reference QR-truncation semantics) for every state of the batch with ONE shared MPO;
        fused like the single-chain path (product cores never formed), every step batched over B."""
        W = mpo.sites if hasattr(mpo, "sites") else mpo
        B, n = self.batch, self.sites_number
        dev = self.sites[0].device
        T = torch.ones((B, 1, 1, 1), dtype=F64, device=dev)            # carry (B, s, l, a)
        out = []
        for k in range(n):
            X, Wk = self.sites[k], W[k]
            _, a, i, b = X.shape
            l, _, o, r = Wk.shape
            s = T.shape[1]
            T1 = torch.empty((B, s, l, i, b), dtype=F64, device=dev)
            ops.gemm(T, X, T1, M=s, N=i * b, K=a, a_m=l * a, a_k=1, b_k=i * b, b_n=1, c_m=l * i * b, c_n=1,
                     batch=B * l, a_b=(s * l * a, a, l), b_b=(a * i * b, 0, l), c_b=(s * l * i * b, i * b, l))
            M = torch.empty((B, s, o, b * r), dtype=F64, device=dev)
            ops.gemm(T1, Wk, M, M=s * b, N=o * r, K=l * i, a_m=(l * i * b, 1, b), a_k=b, b_k=o * r, b_n=1,
                     c_m=(o * b * r, r, b), c_n=(b * r, 1, r), batch=B, a_b=s * l * i * b, b_b=0, c_b=s * o * b * r)
            if k == n - 1:
                out.append(M)
                break
            L = M.reshape(B, s * o, b * r)
            Q, _ = ops.qrt(L, dim, want_S=False)
            kept = Q.shape[2]
            out.append(Q.reshape(B, s, o, kept))
            Tn = torch.empty((B, kept, r, b), dtype=F64, device=dev)
            ops.gemm(Q, L, Tn, M=kept, N=b * r, K=s * o, a_m=1, a_k=kept, b_k=b * r, b_n=1, c_m=r * b, c_n=(1, b, r),
                     batch=B, a_b=s * o * kept, b_b=s * o * b * r, c_b=kept * r * b)
            T = Tn
        return BatchedMatrixProductState(out)
