"""GEMM shapes of the batched SVD apply+round (C4(ii)): event time per shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import defaultdict
from bench import capped_bonds
from syngular_b200 import ops
from syngular_b200.batched import BatchedMatrixProductState as BMPS
from syngular.tensor import MatrixProductOperator as MPO
n, B = 32, 256
dev = torch.device("cuda")
A = BMPS.random(B, (2,) * n, capped_bonds(n, 2, 64)[1:-1], seed=1000, device=dev)
W = MPO.random_cores((2,) * n, (2,) * n, capped_bonds(n, 4, 4)[1:-1], seed=7).sites
A.apply_round_svd(W, 64, chunk=256)
acc = defaultdict(lambda: [0, 0.0])
orig = ops.gemm
import inspect
sig = inspect.signature(orig)
def timed(*a, **k):
    b = sig.bind(*a, **k); b.apply_defaults(); k2 = b.arguments
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(*a, **k); e1.record(); torch.cuda.synchronize()
    key = (k2["M"], k2["N"], k2["K"], k2.get("batch", 1))
    acc[key][0] += 1; acc[key][1] += e0.elapsed_time(e1)
    return r
ops.gemm = timed
A.apply_round_svd(W, 64, chunk=256)
tot = sum(v[1] for v in acc.values())
print("gemm total %.1f ms" % tot)
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:12]:
    fl = 2.0 * k[0] * k[1] * k[2] * k[3]
    print("%-28s n %3d  %7.2f ms  %5.1f%%  %6.2f TFLOP/s" % (k, v[0], v[1], 100 * v[1] / tot, fl * v[0] / v[1] / 1e9))
