"""One call of the spectral-projection solver on a C2-like Gram matrix (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from syngular_b200 import ops

n, ne = 512, 256
rng = np.random.default_rng(0)
q, _ = np.linalg.qr(rng.normal(size=(n, n)))
lam = np.exp(-12.0 * np.arange(n) / n)
A = torch.from_numpy((q * lam) @ q.T).cuda()
sp2, ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 3
U, info = ops.dominant_subspace(A, ne, sp2, ns)
torch.cuda.synchronize()
print(info.cpu().numpy())
