"""Per-step cost of the batched projection kernel: time differences between runs with different iteration limits (one problem per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
n, ne, B = 128, 64, 148
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
lam = np.concatenate([np.exp(-rng.uniform(0, 4, ne)), 1e-3 * np.exp(-rng.uniform(0, 6, n - ne))])
A1 = (Q * lam) @ Q.T
A = torch.from_numpy(np.broadcast_to(0.5 * (A1 + A1.T), (B, n, n)).copy()).cuda()

def t(sp2, ns, reps=5):
    ops.dominant_subspace_batched(A, ne, sp2_max=sp2, ns_max=ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        U, info = ops.dominant_subspace_batched(A, ne, sp2_max=sp2, ns_max=ns)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, info[0, 7].item()

a, ia = t(5, 0); b, ib = t(25, 0)
print("SP2 step: %.2f us  (5 steps %.1f us, 25 steps %.1f us; info %d %d)" % ((b - a) / 20, a, b, ia, ib))
c, ic = t(90, 0); d, id_ = t(90, 8)
print("converged SP2 + 0 NS: %.1f us (info %d); + 8 NS: %.1f us (info %d): NS step %.2f us" % (c, ic, d, id_, (d - c) / 8))
e, ie = t(90, 60)
print("full: %.1f us (info %d)" % (e, ie))
print("DMMA floor of one SP2 step at 37.1 TFLOP/s / 148 SMs: %.2f us" % (2.0 * n ** 3 / (37.1e12 / 148) * 1e6))
