"""C4(ii) -- a shared MPO (chi_W = 4) applied to a batch of chi = 64 states and rounded back by the density-matrix SVD sweep: one call
for a launch list (run under ncu --metrics gpu__time_duration.sum) or, with `time`, the event-timed rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import capped_bonds
from syngular_b200.batched import BatchedMatrixProductState as BMPS
from syngular.tensor import MatrixProductOperator as MPO
n, B = 32, 256
dev = torch.device("cuda")
bonds = capped_bonds(n, 2, 64)[1:-1]
A = BMPS.random(B, (2,) * n, bonds, seed=1000, device=dev)
W = MPO.random_cores((2,) * n, (2,) * n, capped_bonds(n, 4, 4)[1:-1], seed=7).sites
A.apply_round_svd(W, 64, chunk=256)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        A.apply_round_svd(W, 64, chunk=256)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("apply_round_svd: %.2f ms per batch of %d = %.0f states/s" % (ms, B, B / ms * 1e3))
else:
    torch.cuda.profiler.start()
    A.apply_round_svd(W, 64, chunk=256)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
from syngular_b200 import batched
print("projection: taken %d, fallback %d" % (batched.PROJECTION_STATS["taken"], batched.PROJECTION_STATS["fallback"]))
