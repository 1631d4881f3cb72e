"""One launch of the fused TF32 TensorDense kernel for ncu (batch from argv, default 16384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
dev = torch.device("cuda")
G = [torch.from_numpy(rng.normal(scale=0.05, size=s).astype(np.float32)).to(dev) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
packed = ops.tt_dense3_pack(*G)
x = torch.randn((B, 4096), dtype=torch.float32, device=dev)
bias = torch.zeros(4096, dtype=torch.float32, device=dev)
out = torch.empty_like(x)
for _ in range(3):
    ops.tt_dense3_tf32(x, packed, bias, relu=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.tt_dense3_tf32(x, packed, bias, relu=True, out=out); e1.record(); torch.cuda.synchronize()
print("batch %d: %.3f ms, %.2f M samples/s" % (B, e0.elapsed_time(e1), B / e0.elapsed_time(e1) / 1e3))
