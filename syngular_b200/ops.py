"""Thin Python wrappers over the C ABI (one function per exported entry point).  No arithmetic happens here."""
import ctypes

import torch

from . import _lib
from ._lib import GemmDesc, check, ix, lib, ptr, require_cuda_f64, stream_ptr


def gemm(A, B, C, M, N, K, a_m, a_k, b_k, b_n, c_m, c_n, batch=1, a_b=0, b_b=0, c_b=0, alpha=1.0, beta=0.0):
    """C[m,n] = alpha * sum_k A[m,k] B[k,n] + beta * C[m,n] with two-level strided indices (see syngular_b200.h).
    A, B, C are CUDA float64 tensors used as base pointers (their own strides are ignored)."""
    require_cuda_f64(A, B, C)
    d = GemmDesc(int(M), int(N), int(K), int(batch), ix(a_m), ix(a_k), ix(a_b), ix(b_k), ix(b_n), ix(b_b),
                 ix(c_m), ix(c_n), ix(c_b), float(alpha), float(beta))
    check(lib.syn_gemm_f64(ctypes.byref(d), ptr(A), ptr(B), ptr(C), stream_ptr()), "syn_gemm_f64")
    return C


def matmul(a, b, out=None, alpha=1.0, beta=0.0):
    """out = alpha * a @ b + beta * out for 2-D (or batched 3-D) strided views; no copies are made."""
    require_cuda_f64(a, b)
    if a.dim() == 2 and b.dim() == 2:
        M, K = a.shape
        K2, N = b.shape
        assert K == K2, (a.shape, b.shape)
        if out is None:
            out = torch.empty((M, N), dtype=torch.float64, device=a.device)
        return gemm(a, b, out, M, N, K, a.stride(0), a.stride(1), b.stride(0), b.stride(1), out.stride(0), out.stride(1),
                    alpha=alpha, beta=beta)
    assert a.dim() == 3 and b.dim() == 3 and a.shape[0] == b.shape[0]
    nb, M, K = a.shape
    _, K2, N = b.shape
    assert K == K2
    if out is None:
        out = torch.empty((nb, M, N), dtype=torch.float64, device=a.device)
    return gemm(a, b, out, M, N, K, a.stride(1), a.stride(2), b.stride(1), b.stride(2), out.stride(1), out.stride(2),
                batch=nb, a_b=a.stride(0), b_b=b.stride(0), c_b=out.stride(0), alpha=alpha, beta=beta)
