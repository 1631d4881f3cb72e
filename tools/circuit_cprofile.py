import os, sys, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from syngular.quantum import Circuit
nq, depth, chi = 50, 14, 512
rng = np.random.default_rng(3)
def haar4():
    z = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    q, r = np.linalg.qr(z)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)
structure = [(haar4(), i) for layer in range(depth) for i in range(layer % 2, nq - 1, 2)]
Circuit(8, structure=[(g, i % 7) for g, i in structure[:40]], chi_max=16).run()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
Circuit(nq, structure=structure, chi_max=chi).run()
torch.cuda.synchronize()
print("first run (cold allocator): %.2f s; reserved %.1f GB" % (time.perf_counter() - t0, torch.cuda.memory_reserved() / 2**30))
t0 = time.perf_counter()
Circuit(nq, structure=structure, chi_max=chi).run()
torch.cuda.synchronize()
print("second run: %.2f s; reserved %.1f GB" % (time.perf_counter() - t0, torch.cuda.memory_reserved() / 2**30))
pr = cProfile.Profile()
pr.enable()
Circuit(nq, structure=structure, chi_max=chi).run()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(12)
