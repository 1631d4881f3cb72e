import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import capped_bonds
from syngular_b200 import batched, ops
from syngular_b200.batched import BatchedMatrixProductState as BMPS
from syngular.tensor import MatrixProductOperator as MPO
orig = ops.dominant_subspace_batched
def spy(A, ne, **kw):
    U, info = orig(A, ne, **kw)
    h = info.cpu().numpy()
    sp2 = h[:, 7] % 1000; ns = (h[:, 7] // 1000) % 1000; lift = h[:, 7] // 1000000
    bad = batched._rejected(h, ne, False)
    print("n %3d ne %3d: sp2 mean %.1f max %d | ns mean %.1f max %d | lift mean %.1f max %d | bad %d: tr-ne %s dev %s idem %s" % (
        A.shape[1], ne, sp2.mean(), sp2.max(), ns.mean(), ns.max(), lift.mean(), lift.max(), len(bad),
        np.abs(h[bad, 0] - ne)[:3], h[bad, 4][:3], h[bad, 6][:3]))
    return U, info
ops.dominant_subspace_batched = spy
n, B = 32, 256
dev = torch.device("cuda")
bonds = capped_bonds(n, 2, 64)[1:-1]
A = BMPS.random(B, (2,) * n, bonds, seed=1000, device=dev)
W = MPO.random_cores((2,) * n, (2,) * n, capped_bonds(n, 4, 4)[1:-1], seed=7).sites
A.apply_round_svd(W, 64, chunk=256)
