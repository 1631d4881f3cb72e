"""Parity + timing of the CTA-pair TF32 layer kernel (syn_tt_dense3_tf32_pair) against the float64 restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
from oracle import tensordense_numpy as TD
rng = np.random.default_rng(0)
dev = torch.device("cuda")
G = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
packed = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in G])
I1 = np.zeros((16, 16, 16), np.float32); I2 = np.zeros((16, 16, 16, 16), np.float32); I3 = np.zeros((16, 16, 16), np.float32)
for i in range(16):
    I1[i, 15 - i, 3] = 1; I2[i, 15 - i, 3, 7] = 1; I3[i, 15 - i, 7] = 1
pk = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in (I1, I2, I3)])
xs = (np.arange(3 * 4096) % 1021).astype(np.float32).reshape(3, 4096)
ys = ops.tt_dense3_tf32(torch.from_numpy(xs).to(dev), pk, None, relu=False, pair=True).cpu().numpy()
want = xs.reshape(3, 16, 16, 16)[:, ::-1, ::-1, ::-1].reshape(3, 4096)
bad = np.argwhere(ys != want)
print("permutation layer: mismatches %d of %d" % (len(bad), xs.size))
for s, f in bad[:8]:
    print("   sample %d out (o1,o2,o3)=(%d,%d,%d): got %g expected %g" % (s, f // 256, (f // 16) % 16, f % 16, ys[s, f], want[s, f]))
for batch in (1, 2, 3, 75, 300, 1111):
    x = rng.normal(size=(batch, 4096)).astype(np.float32)
    bias = (0.01 * rng.normal(size=4096)).astype(np.float32)
    y = ops.tt_dense3_tf32(torch.from_numpy(x).to(dev), packed, torch.from_numpy(bias).to(dev), relu=True, pair=True).cpu().numpy()
    r = TD.forward(x.astype(np.float64), [g.astype(np.float64) for g in G], bias.astype(np.float64), "relu")
    print("batch %5d: max err %.3e (scale %.3e)" % (batch, np.max(np.abs(y - r)), np.max(np.abs(r))))
B = 65536
xb = torch.randn((B, 4096), dtype=torch.float32, device=dev); out = torch.empty_like(xb); bz = torch.zeros(4096, dtype=torch.float32, device=dev)
for _ in range(3):
    ops.tt_dense3_tf32(xb, packed, bz, relu=True, out=out, pair=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.tt_dense3_tf32(xb, packed, bz, relu=True, out=out, pair=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("batch %d: %.3f ms = %.2f M samples/s, %.1f TFLOP/s" % (B, ms, B / ms / 1e3, 3.7748736e7 * B / ms / 1e9))
