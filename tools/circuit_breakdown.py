"""Where the time of the complex128 circuit (BASELINE configs[2]) goes: CUDA-event time per library call of syngular_b200.ops, by wrapping
every public wrapper (synchronises around each call: the total is larger than the untimed run, the SHARES are what matters).
python tools/circuit_breakdown.py [chi_max] [depth]"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import defaultdict
from syngular_b200 import ops
from syngular.quantum import Circuit

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 512
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 12
acc = defaultdict(lambda: [0, 0.0])
depth_now = [0]


def wrap(name, fn):
    def inner(*a, **k):
        if depth_now[0]:
            return fn(*a, **k)
        depth_now[0] = 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.synchronize()
            key = name
            if name == "gemm":
                key = "gemm M%d N%d K%d b%d" % (k.get("M", 0), k.get("N", 0), k.get("K", 0), k.get("batch", 1))
            elif a and hasattr(a[0], "shape"):
                key = "%s %s" % (name, tuple(a[0].shape))
            acc[key][0] += 1
            acc[key][1] += time.perf_counter() - t0
            depth_now[0] = 0
    return inner


for name in dir(ops):
    fn = getattr(ops, name)
    if isinstance(fn, types.FunctionType) and fn.__module__ == ops.__name__ and not name.startswith("_") and name not in (
            "ptr", "check", "stream_ptr", "workspace", "require_cuda_f64") and not name.endswith("_fits"):
        setattr(ops, name, wrap(name, fn))

nq = 50
rng = np.random.default_rng(3)


def haar4():
    z = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    q, r = np.linalg.qr(z)
    return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)


structure = [(haar4(), i) for layer in range(depth) for i in range(layer % 2, nq - 1, 2)]
torch.cuda.synchronize()
t0 = time.perf_counter()
Circuit(nq, structure=structure, chi_max=chi).run()
torch.cuda.synchronize()
total = time.perf_counter() - t0
lib = sum(v[1] for v in acc.values())
print("chi_max %d depth %d: %d gates, %.2f s wall with per-call synchronisation, %.2f s inside library calls" % (chi, depth, len(structure), total, lib))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:28]:
    print("  %-44s n %5d  %8.1f ms  %5.1f %%   %7.3f ms each" % (k, v[0], v[1] * 1e3, 100 * v[1] / total, v[1] / v[0] * 1e3))
