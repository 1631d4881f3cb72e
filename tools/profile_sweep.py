"""One warm-up + one measured sweep of the C2 hot path, for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_sweep.py [svd|qr]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from syngular.tensor import _sweeps as sw

mode = sys.argv[1] if len(sys.argv) > 1 else "svd"
chi = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CHI
X, W = bench.make_chain(2, chi=chi)
Xd = [sw.as_core(x) for x in X]
Wd = [sw.as_core(w) for w in W]
fn = (lambda: sw.apply_round_dm(Xd, Wd, chi)) if mode == "svd" else (lambda: sw.apply_round_qr(Xd, Wd, chi))
fn()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measured")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("mode", mode, "chi", chi, "sweep ms", e0.elapsed_time(e1))
