import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syngular_b200 import ops
shapes = {"p1": (512, 65536, 256), "me": (512, 4096, 4096), "sq": (4096, 4096, 4096)}
which = sys.argv[1] if len(sys.argv) > 1 else "p1"
M, N, K = shapes[which]
a = torch.randn(M, K, dtype=torch.float64, device="cuda"); b = torch.randn(K, N, dtype=torch.float64, device="cuda")
out = torch.empty(M, N, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.matmul(a, b, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.matmul(a, b, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(which, (M, N, K), "%.3f ms  %.2f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
e0.record()
for _ in range(5):
    torch.matmul(a, b, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("cublas", "%.3f ms  %.2f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
