import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
rng = np.random.default_rng(0)
def T(fn, reps=5):
    fn(); best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for (m, n, q) in ((512, 4096, 256), (512, 256, 256), (256, 2048, 128), (128, 1024, 64)):
    A = torch.from_numpy(rng.normal(size=(m, n))).cuda()
    print("qrt %4d x %4d q=%3d : %.3f ms (no S: %.3f ms)" % (m, n, q, T(lambda: ops.qrt(A, q)), T(lambda: ops.qrt(A, q, want_S=False))))
