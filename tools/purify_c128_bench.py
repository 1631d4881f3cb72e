"""The complex Hermitian projection solver at the size of a saturated chi_max = 512 circuit bond (m = 1024, k = 512): time and iteration counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = m // 2
rng = np.random.default_rng(0)
Z = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
Q, _ = np.linalg.qr(Z)
lam = np.concatenate([np.exp(-rng.uniform(0, 6, k)), 1e-4 * np.exp(-rng.uniform(0, 8, m - k))])
H = (Q * lam) @ Q.conj().T
H = 0.5 * (H + H.conj().T)
Hre, Him = torch.from_numpy(np.ascontiguousarray(H.real)).cuda(), torch.from_numpy(np.ascontiguousarray(H.imag)).cuda()
for _ in range(2):
    Ure, Uim, info = ops.dominant_subspace_c128(Hre, Him, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Ure, Uim, info = ops.dominant_subspace_c128(Hre, Him, k)
e1.record(); torch.cuda.synchronize()
h = info.cpu().numpy()
print("m %d k %d: %.2f ms per call; sp2 %d, ns %d, lift %d; tr %.6f dev %.1e" % (m, k, e0.elapsed_time(e1) / 5, int(h[7]) % 1000, (int(h[7]) // 1000) % 1000, int(h[7]) // 1000000, h[0], h[4]))
