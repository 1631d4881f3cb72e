"""How well conditioned are the blocks the QR sweep orthonormalises (C2), and would a diagonal column scaling help the Newton-Schulz start?
Probe (diagnostic only: torch.linalg.svdvals): singular values of L[:, :q] and of L[:, :q] D, D = 1 / column norms, at a few plateau sites."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
import bench
from syngular.tensor import _sweeps as sw

X, W = bench.make_chain(2)
Xd, Wd = [sw.as_core(x) for x in X], [sw.as_core(w) for w in W]
n = len(Xd)
T = torch.ones((1, 1, 1), dtype=torch.float64, device=Xd[0].device)
for k in range(n - 1):
    M = sw.contract_carry(T, Xd[k], Wd[k])
    s, o, _ = M.shape
    b, r = Xd[k].shape[2], Wd[k].shape[3]
    L = M.reshape(s * o, b * r)
    if k in (10, 20, 32, 45):
        A = L[:, :256]
        sv = torch.linalg.svdvals(A)
        d = A.norm(dim=0)
        sv2 = torch.linalg.svdvals(A / d)
        def steps(sv):
            x = sv / sv.max()
            return math.ceil(math.log2(float(1.0 / x.min()))) + 4
        print("site %2d L %s: cond %.1e -> with column scaling %.1e; column norms max/min %.1f; ~NS steps %d -> %d" % (
            k, tuple(L.shape), float(sv[0] / sv[-1]), float(sv2[0] / sv2[-1]), float(d.max() / d.min()), steps(sv), steps(sv2)))
    Q, _ = sw.qrt_step(L, 256, want_S=False)
    kept = Q.shape[1]
    T = sw._carry_from(Q, L, kept, b, r, transposed_basis=False)
