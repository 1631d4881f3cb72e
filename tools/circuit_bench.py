"""BASELINE configs[2]: 50-qubit brick-wall circuit of Haar-random two-qubit unitaries, depth 20, SVD truncation chi_max (default 512),
complex128 on the FP64 kernels.  Prints gates/s, the final norm and the bond profile.   python tools/circuit_bench.py [chi_max] [depth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from syngular.quantum import Circuit
from syngular_b200 import ops

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nq = 50
rng = np.random.default_rng(3)


def haar4():
    z = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    q, r = np.linalg.qr(z)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)


structure = [(haar4(), i) for layer in range(depth) for i in range(layer % 2, nq - 1, 2)]
Circuit(8, structure=[(g, i % 7) for g, i in structure[:40]], chi_max=16).run()
Circuit(nq, structure=structure, chi_max=chi).run()          # warm-up at full size: the caching allocator has seen every shape (cold run: ~1.4x slower)
torch.cuda.synchronize()
l0 = ops.lib.syn_launch_count()
t0 = time.perf_counter()
c = Circuit(nq, structure=structure, chi_max=chi)
c.run()
torch.cuda.synchronize()
sec = time.perf_counter() - t0
st = c.get().state
print("chi_max %d depth %d: %d gates in %.2f s = %.1f gates/s; kernels %d; max bond %d; norm2 %.6e; bonds %s" % (
    chi, depth, len(structure), sec, len(structure) / sec, ops.lib.syn_launch_count() - l0, max(s.shape[2] for s in st.sites[:-1]),
    float(np.real(st.conj() | st)), [int(s.shape[2]) for s in st.sites[:-1]]))
