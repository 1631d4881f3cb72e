"""One GEMM shape, timed: python tools/gemm_one_shape.py M N K [ta]   (ta = 1: A is stored K x M, read transposed, as the carry product U^T M).
Combine with SYN_GEMM_CFG=L|W|S to force a tile configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from syngular_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
ta = len(sys.argv) > 4 and sys.argv[4] == "1"
dev = torch.device("cuda")
A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device=dev)
B = torch.randn((K, N), dtype=torch.float64, device=dev)
C = torch.empty((M, N), dtype=torch.float64, device=dev)
kw = dict(M=M, N=N, K=K, a_m=1 if ta else K, a_k=M if ta else 1, b_k=N, b_n=1, c_m=N, c_n=1)
for _ in range(5):
    ops.gemm(A, B, C, **kw)
ref = (A.t() if ta else A) @ B
err = float((C - ref).abs().max() / ref.abs().max())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.gemm(A, B, C, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("cfg %s  %d x %d x %d ta=%d: %.1f us, %.2f TFLOP/s (back to back, L2-warm), err %.1e" % (os.environ.get("SYN_GEMM_CFG", "auto"), M, N, K, ta, ms * 1e3, 2.0 * M * N * K / ms / 1e9, err))
