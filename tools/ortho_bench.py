import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
for (m, q, cond) in ((2048, 1024, 1e2), (2048, 1024, 1e6), (2048, 512, 1e2), (1024, 512, 1e2), (512, 256, 1e2)):
    rng = np.random.default_rng(0)
    U, _ = np.linalg.qr(rng.normal(size=(m, q)))
    V, _ = np.linalg.qr(rng.normal(size=(q, q)))
    s = np.exp(-np.linspace(0, np.log(cond), q))
    A = torch.from_numpy((U * s) @ V.T).cuda()
    Q, info = ops.orthonormalize_columns(A); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): Q, info = ops.orthonormalize_columns(A)
    e1.record(); torch.cuda.synchronize()
    h = info.cpu().numpy()
    print("m %d q %d cond %.0e: %.2f ms, ns steps %d, dev %.1e" % (m, q, cond, e0.elapsed_time(e1) / 3, (int(h[7]) // 1000) % 1000, h[4]))
