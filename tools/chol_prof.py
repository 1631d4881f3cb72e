"""One call of the shifted Cholesky + Jacobi + finalize on a 512 x 512 Gram matrix (run under ncu for a per-kernel launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
F = rng.normal(size=(n, 4 * n))
G = torch.from_numpy(F @ F.T).cuda()
for _ in range(2):
    A = G.clone()
    B, shift = ops.chol_upper(A)
    torch.cuda.synchronize()
# event timing (the per-launch times under ncu are cold-cache and serialised)
best = 1e9
for _ in range(5):
    A = G.clone(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.chol_upper(A); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("chol_upper n=%d: %.3f ms" % (n, best))
