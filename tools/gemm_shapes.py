"""Per-shape timing of every strided-GEMM launch in one C2 sweep (SVD/density-matrix mode and QR mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import defaultdict
import torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "svd"
X, W = bench.make_chain(2)
Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
fn = (lambda: sw.apply_round_dm(Xd, Wd, 256)) if mode == "svd" else (lambda: sw.apply_round_qr_steps(Xd, Wd, 256))
fn(); torch.cuda.synchronize()
ops.GEMM_PROFILE = []
fn(); torch.cuda.synchronize()
prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for e0, e1, fl, by, shape in prof:
    a = agg[shape]; a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl; a[3] += by
tot = sum(a[1] for a in agg.values())
print("mode %s: %d python-level gemm calls, %.2f ms total, %.2f TFLOP/s overall" % (mode, len(prof), tot, sum(a[2] for a in agg.values()) / tot / 1e9))
print("%-28s %5s %9s %7s %9s %9s" % ("(M, N, K, batch)", "count", "ms", "share", "TFLOP/s", "GB/s(alg)"))
for shape, (c, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-28s %5d %9.3f %6.1f%% %9.2f %9.0f" % (str(shape), c, ms, 100 * ms / tot, fl / ms / 1e9, by / ms / 1e6))
