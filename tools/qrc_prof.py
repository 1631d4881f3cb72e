import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
A = torch.from_numpy(np.random.default_rng(0).normal(size=(512, 256))).cuda()
for _ in range(3):
    ops.qrt(A, 256, want_S=False)
torch.cuda.synchronize()
