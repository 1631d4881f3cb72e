"""Quick timing of the fused projection solver (n = 512 -> 256) on Gram matrices of the C2 sweep and of the NS-only orthonormalisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops


def time_it(fn, reps=20):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


captured = {}
X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
cap = {20: None, 32: None, 50: None}
sw.apply_round_dm(Xd, Wd, 256, capture=cap)
for k, (A, U) in cap.items():
    Uf, info = ops.dominant_subspace(A, 256, 52, 26, sp2_max=90, ns_max=60)
    h = info.cpu().numpy()
    lam, V = np.linalg.eigh(A.cpu().numpy())
    Vk = V[:, -256:]
    Ug = Uf.cpu().numpy()
    ms = time_it(lambda: ops.dominant_subspace(A, 256, 52, 26, sp2_max=90, ns_max=60))
    print("site %d: fused solver %.3f ms, sp2 %d ns %d lift %d, dev %.1e, |P - P_eigh| %.2e" % (
        k, ms, int(h[7]) % 1000, (int(h[7]) // 1000) % 1000, int(h[7]) // 1000000, h[4], np.linalg.norm(Ug @ Ug.T - Vk @ Vk.T)))
L = torch.randn(512, 4096, dtype=torch.float64, device="cuda")
Q, info = ops.orthonormalize_columns(L[:, :256])
h = info.cpu().numpy()
ms = time_it(lambda: ops.orthonormalize_columns(L[:, :256]))
print("orthonormalize_columns 512 x 256: %.3f ms, ns %d, dev %.1e, |QtQ-I| %.1e" % (
    ms, (int(h[7]) // 1000) % 1000, h[4], float((Q.t() @ Q - torch.eye(256, dtype=torch.float64, device="cuda")).abs().max())))
