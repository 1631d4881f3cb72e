import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
B = torch.from_numpy(rng.normal(size=(n, 4 * n))).cuda()
A = B @ B.T
# the production pipeline of the density-matrix rounding: shifted Cholesky factor, then Jacobi on its rows
for _ in range(2):
    Bf, shift = ops.chol_upper(A.clone()); ops.jacobi_rows(Bf, null_rel=0.0); torch.cuda.synchronize()
print("sweeps", ops.jacobi_sweeps_used())
