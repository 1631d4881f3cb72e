"""Per-bond record of the projection solver over one C2 sweep: size, target, steps, lift (log2 |A|_F / lambda_cut), accepted or not."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops

X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
orig = ops.dominant_subspace
log = []
def spy(A, ne, *a, **k):
    U, info = orig(A, ne, *a, **k)
    h = info.cpu().numpy()
    lam = torch.linalg.eigvalsh(A).flip(0)
    log.append((A.shape[0], ne, int(h[7]) % 1000, (int(h[7]) // 1000) % 1000, int(h[7]) // 1000000, (lam[ne - 1] / lam[0]).item(),
                ((lam[ne - 1] - lam[ne]) / lam[0]).item(), (h[3] / lam[0].item())))
    return U, info
ops.dominant_subspace = spy
sw.PURIFY_STATS.update(taken=0, fallback=0)
sw.apply_round_dm(Xd, Wd, 256)
print(sw.PURIFY_STATS)
for k, l in enumerate(log):
    print("call %2d n=%d ne=%d sp2=%d ns=%d lift=%d lam_cut/lam0 %.1e gap/lam0 %.1e |A|_F/lam0 %.2f" % ((k,) + l))
