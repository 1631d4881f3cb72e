"""Timing of the spectral-projection solver against Cholesky + Jacobi on the Gram matrices of the C2 sweep (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops


def time_it(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


captured = []
orig = sw.gram_with_environment
def spy(*a, **k):
    A = orig(*a, **k)
    captured.append(A.clone())
    return A
X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
sw.gram_with_environment = spy
sw.PURIFY_STATS.update(taken=0, fallback=0)
sw.apply_round_dm(Xd, Wd, 256)
sw.gram_with_environment = orig
print("sweep stats", sw.PURIFY_STATS, "sites", len(captured))
for idx in (7, 8, 9, 20, 32, 50, 55, 56, 57):
    A = captured[idx]; n = A.shape[0]
    if n <= 256:
        continue
    ne = 256
    lam = torch.linalg.eigvalsh(A).flip(0)
    gap = ((lam[ne - 1] - lam[ne]) / lam[0]).item() if ne < n else float("nan")
    line = "site %2d n=%4d lam[ne]/lam0 %.1e gap/lam0 %.1e |" % (idx, n, (lam[ne - 1] / lam[0]).item(), gap)
    for sp2, ns in ((52, 26),):
        U, info = ops.dominant_subspace(A, ne, sp2, ns, fused=False)
        h = info.cpu().numpy()
        ms = time_it(lambda: ops.dominant_subspace(A, ne, sp2, ns, fused=False))
        line += " [%d/%d: %.3f ms tr-ne %.1e idem %.1e dev %.1e]" % (sp2, ns, ms, h[0] - ne, h[6], h[4])
    Uf, info = ops.dominant_subspace(A, ne, fused=True)
    h = info.cpu().numpy()
    ms = time_it(lambda: ops.dominant_subspace(A, ne, fused=True))
    pdiff = (Uf @ Uf.t() - U @ U.t()).abs().max().item()
    line += " [FUSED: %.3f ms iters %d tr-ne %.1e idem %.1e dev %.1e |P-P_unfused| %.1e]" % (ms, int(h[7]), h[0] - ne, h[6], h[4], pdiff)
    work = A.clone()
    def jac():
        work.copy_(A)
        B, shift = ops.chol_upper(work)
        ops.jacobi_rows(B, null_rel=0.0)
        ops.jacobi_finalize(B, ne, 0.0, rank_tol=3.2e-7, sqrt_mode=2, shift=shift)
    line += " | chol+jacobi %.3f ms" % time_it(jac, 3)
    print(line)
# phases of one call at n = 512
A = captured[32]
for sp2, ns in ((44, 0), (1, 22), (1, 0)):
    print("n=512 sp2=%d ns=%d: %.3f ms" % (sp2, ns, time_it(lambda: ops.dominant_subspace(A, 256, sp2, ns, fused=False))))
