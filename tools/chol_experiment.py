"""Experiment: one-sided Jacobi on the Gram matrices of the C2 sweep directly versus on the transposed Cholesky factor
(G + delta I = L L^T; rows of L^T, csrc/chol.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops


def time_it(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


captured = []
orig = ops.jacobi_rows
def spy(G, *a, **k):
    if G.shape[-1] >= 64 and len(captured) < 64:
        captured.append(G.clone())
    return orig(G, *a, **k)
X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
ops.jacobi_rows = spy
sw.CHOLESKY_MIN_N = 0            # capture the Gram matrices themselves, not their factors
sw.apply_round_dm(Xd, Wd, 256)
sw.CHOLESKY_MIN_N = 129
ops.jacobi_rows = orig
for idx in (3, 8, 20, 32, 50):
    A = captured[idx]; n = A.shape[-1]
    lam = torch.linalg.eigvalsh(A).flip(0)
    work = A.clone()
    ms1 = time_it(lambda: (work.copy_(A), ops.jacobi_rows(work, null_rel=1e-13))); s1 = ops.jacobi_sweeps_used()
    B, shift = ops.chol_upper(A.clone()); delta = shift.item()
    ms_chol = time_it(lambda: (work.copy_(A), ops.chol_upper(work)))
    ms2 = time_it(lambda: (work.copy_(B), ops.jacobi_rows(work, null_rel=3.2e-7))); s2 = ops.jacobi_sweeps_used()
    ms2b = time_it(lambda: (work.copy_(B), ops.jacobi_rows(work, null_rel=0.0))); s2b = ops.jacobi_sweeps_used()
    ev = (work * work).sum(1).sort(descending=True).values - delta
    k = n // 2
    err = ((ev[:k] - lam[:k]).abs() / lam[:k]).max().item()
    print("matrix %2d n=%4d lam[n/2]/lam0 %.1e lam[-1]/lam0 %.1e | on G: %.3f ms %s sweeps | on chol^T: %.3f ms %s sweeps (no null floor: %.3f ms %s) | chol %.3f ms | top-half eig relerr %.1e"
          % (idx, n, (lam[k] / lam[0]).item(), (lam[-1] / lam[0]).item(), ms1, s1, ms2, s2, ms2b, s2b, ms_chol, err))
