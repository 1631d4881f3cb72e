"""Experiment: how many FP64 Jacobi sweeps remain after an FP32-accurate pre-diagonalisation?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops

captured = []
orig = ops.jacobi_rows
def spy(G, *a, **k):
    if G.shape[-1] >= 256 and len(captured) < 40:
        captured.append(G.clone())
    return orig(G, *a, **k)
X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
ops.jacobi_rows = spy
sw.apply_round_dm(Xd, Wd, 256)
ops.jacobi_rows = orig

def timed(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)

for idx in (5, 20, 35):
    A = captured[idx]; n = A.shape[0]
    A = 0.5 * (A + A.T)
    w0 = A.clone(); t_full = timed(lambda: ops.jacobi_rows(w0)); s_full = ops.jacobi_sweeps_used()
    lam32, U32 = torch.linalg.eigh(A.float())
    U = U32.double().T.contiguous()                       # rows = approximate eigenvectors
    for _ in range(2):                                    # Newton-Schulz orthonormalisation
        U = 1.5 * U - 0.5 * (U @ U.T) @ U
    print("n=%d  orth err %.1e" % (n, (U @ U.T - torch.eye(n, device=U.device, dtype=U.dtype)).abs().max().item()))
    Ap = U @ A @ U.T
    Ap = 0.5 * (Ap + Ap.T)
    off = (Ap - torch.diag(torch.diagonal(Ap))).abs().max().item() / Ap.abs().max().item()
    w1 = Ap.clone(); t_pre = timed(lambda: ops.jacobi_rows(w1)); s_pre = ops.jacobi_sweeps_used()
    Ut, sigma, info, winfo = ops.jacobi_finalize(w1, n, sqrt_mode=True)
    lam = torch.linalg.eigvalsh(A).flip(0).clamp_min(0).sqrt()
    print("   full jacobi: %.2f ms %s sweeps | preconditioned (offdiag %.1e): %.2f ms %s sweeps | sigma err %.1e" % (
        t_full, s_full, off, t_pre, s_pre, (sigma - lam).abs().max().item() / lam[0].item()))
    # looser FP32-like perturbation: emulate a 1e-5 accurate basis
    Un = U + 1e-5 * torch.randn_like(U)
    for _ in range(2):
        Un = 1.5 * Un - 0.5 * (Un @ Un.T) @ Un
    Ap = Un @ A @ Un.T; Ap = 0.5 * (Ap + Ap.T)
    w2 = Ap.clone(); t2 = timed(lambda: ops.jacobi_rows(w2)); print("   1e-5-perturbed basis: %.2f ms %s sweeps" % (t2, ops.jacobi_sweeps_used()))
