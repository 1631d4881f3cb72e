"""BASELINE configs[3] overlaps (8192 x N=32, d=2, chi=64): fused transfer-matrix kernel vs the GEMM-per-site route."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from syngular_b200 import ops
from syngular_b200.batched import BatchedMatrixProductState as BMPS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bonds = bench.capped_bonds(32, 2, 64)[1:-1]
A = BMPS.random(B, (2,) * 32, bonds, seed=1)
C = BMPS.random(B, (2,) * 32, bonds, seed=2)
full = [1] + list(bonds) + [1]
flops = sum(4.0 * 2 * a * b * max(a, b) for a, b in zip(full[:-1], full[1:])) * B
bytes_ = 2 * 8.0 * sum(a * 2 * b for a, b in zip(full[:-1], full[1:])) * B


def timed(fn, reps=3):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for fused in (True, False):
    ops.OVERLAP_FUSED = fused
    ms = timed(lambda: A.overlap(C))
    print("%s: %.3f ms per %d pairs = %.0f states/s, %.2f TFLOP/s FP64, %.0f GB/s of core reads"
          % ("fused kernel " if fused else "GEMM per site", ms, B, B / (ms * 1e-3), flops / (ms * 1e-3) / 1e12, bytes_ / (ms * 1e-3) / 1e9))
ops.OVERLAP_FUSED = True
r1 = A.overlap(C); ops.OVERLAP_FUSED = False; r2 = A.overlap(C); ops.OVERLAP_FUSED = True
print("max rel diff fused vs gemm: %.2e" % float(((r1 - r2).abs().max() / r2.abs().max()).item()))
# how much of the time is waiting for HBM?  one side shared by all pairs (its cores stay in L2): half the HBM traffic, the same arithmetic
Cs = [c[:1].contiguous() for c in C.sites]
ms = timed(lambda: ops.overlap_batched(A.sites, Cs))
print("fused kernel, B side shared (L2-resident): %.3f ms per %d pairs, %.2f TFLOP/s" % (ms, B, flops / (ms * 1e-3) / 1e12))
