"""Time the one-sided Jacobi kernel on (a) the actual density-matrix Gram matrices of the C2 sweep and (b) synthetic matrices."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops


def time_it(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


captured = []
orig = ops.jacobi_rows


def spy(G, *a, **k):
    if G.shape[-1] >= 64 and len(captured) < 64:
        captured.append(G.clone())
    return orig(G, *a, **k)


chi = int(sys.argv[1]) if len(sys.argv) > 1 else 256
X, W = bench.make_chain(2, chi=chi)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
ops.jacobi_rows = spy
sw.apply_round_dm(Xd, Wd, chi)
ops.jacobi_rows = orig
print("captured", len(captured), "gram matrices; sizes", sorted(set(g.shape[-1] for g in captured)))
for idx in (2, 5, len(captured) // 2, len(captured) - 3):
    A = captured[idx]
    n = A.shape[-1]
    work = A.clone()
    ms = time_it(lambda: (work.copy_(A), ops.jacobi_rows(work)))
    ms_copy = time_it(lambda: work.copy_(A))
    sweeps = ops.jacobi_sweeps_used()
    lam = torch.linalg.eigvalsh(A)
    print("site-matrix %2d n=%4d  jacobi %.3f ms  sweeps %s  cond(A)=%.2e" % (idx, n, ms - ms_copy, sweeps, (lam[-1] / lam[0].abs().clamp_min(1e-300)).item()))
rng = np.random.default_rng(0)
for n in (128, 256, 512, 1024):
    B = torch.from_numpy(rng.normal(size=(n, 4 * n))).cuda()
    A = B @ B.T
    R = torch.linalg.qr(torch.from_numpy(rng.normal(size=(2 * n, n))).cuda(), mode="r")[1].contiguous()
    for name, M0 in (("wishart-gram", A), ("triangular-R", R)):
        work = M0.clone()
        ms = time_it(lambda: (work.copy_(M0), ops.jacobi_rows(work)))
        print("%-13s n=%4d  jacobi %.3f ms  sweeps %s" % (name, n, ms, ops.jacobi_sweeps_used()))
# batched small problems (C4-like): 1024 x (128 x 128)
B = torch.from_numpy(rng.normal(size=(1024, 128, 256))).cuda()
A = B @ B.transpose(1, 2)
work = A.clone()
ms = time_it(lambda: (work.copy_(A), ops.jacobi_rows(work)))
print("batched 1024 x (128x128): %.3f ms total, %.2f us per problem, sweeps[0:4] %s" % (ms, 1e3 * ms / 1024, ops.jacobi_sweeps_used()[:4]))
