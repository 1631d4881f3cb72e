"""Phase clocks of the Jacobi kernel (build with SYN_NVCC_EXTRA=-DSYN_JACOBI_TIMING): prints cycles per phase for CTA 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
F = rng.normal(size=(n, 4 * n)) * np.exp(-np.arange(n) / 40.0)[:, None]
G = torch.from_numpy(F @ F.T).cuda()
for _ in range(2):
    A = G.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.jacobi_rows(A); e1.record(); torch.cuda.synchronize()
    print("n=%d  %.3f ms  sweeps %s" % (n, e0.elapsed_time(e1), ops.jacobi_sweeps_used()))
