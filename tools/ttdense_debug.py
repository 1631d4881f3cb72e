"""Bring-up / timing of the fused tcgen05 TF32 TensorDense kernel (csrc/ttdense.cu): packed-image check against a numpy restatement of
the swizzled layouts, structured inputs that expose index permutations, random parity, and the BASELINE configs[4] timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from syngular_b200 import ops


def sw128(row, kb):
    return (row >> 3) * 1024 + (row & 7) * 128 + ((((kb >> 4) ^ (row & 7)) & 7) << 4) + (kb & 15)


def sw64(row, kb):
    return (row >> 3) * 512 + (row & 7) * 64 + ((((kb >> 4) ^ ((row & 7) >> 1)) & 3) << 4) + (kb & 15)


def pack_numpy(G1, G2, G3):
    img = np.zeros(8192 + 65536, dtype=np.float32)
    for t in range(2):
        for row in range(128):
            for i3 in range(16):
                img[(t * 8192 + sw64(row, i3 * 4)) // 4] = G3[i3, 8 * t + (row >> 4), row & 15]
    for a in range(8):
        for o1 in range(16):
            for kk in range(32):
                img[4096 + (a * 2048 + sw128(o1, kk * 4)) // 4] = G1[kk & 15, o1, 2 * a + (kk >> 4)]
    for h in range(2):
        for a in range(8):
            for n in range(128):
                for kk in range(32):
                    img[8192 + h * 32768 + (a * 16384 + sw128(n, kk * 4)) // 4] = G2[2 * a + (kk >> 4), 8 * h + (n & 7), n >> 3, kk & 15]
    return img


def ref(x, G1, G2, G3, bias, relu):
    """float64 restatement by three tensordot steps (np.einsum's greedy path can fall off BLAS for some batch sizes: minutes)."""
    from oracle import tensordense_numpy as TD
    return TD.forward(x.astype(np.float64), [g.astype(np.float64) for g in (G1, G2, G3)], None if bias is None else bias.astype(np.float64),
                      "relu" if relu else None)


def run(x, G1, G2, G3, bias, relu):
    dev = torch.device("cuda")
    packed = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in (G1, G2, G3)])
    y = ops.tt_dense3_tf32(torch.from_numpy(x).to(dev), packed, torch.from_numpy(bias).to(dev) if bias is not None else None, relu=relu)
    torch.cuda.synchronize()
    return y.cpu().numpy(), packed


rng = np.random.default_rng(0)
G1 = rng.normal(scale=0.05, size=(16, 16, 16)).astype(np.float32)
G2 = rng.normal(scale=0.05, size=(16, 16, 16, 16)).astype(np.float32)
G3 = rng.normal(scale=0.05, size=(16, 16, 16)).astype(np.float32)
x = rng.normal(size=(3, 4096)).astype(np.float32)
y, packed = run(x, G1, G2, G3, None, False)
print("packed image == numpy restatement:", bool(np.array_equal(packed.cpu().numpy(), pack_numpy(G1, G2, G3))))

# structured test: identity cores (bond index 0 only) -> y == x exactly (values are small integers: exact in TF32)
I1 = np.zeros((16, 16, 16), np.float32); I2 = np.zeros((16, 16, 16, 16), np.float32); I3 = np.zeros((16, 16, 16), np.float32)
for i in range(16):
    I1[i, i, 0] = 1; I2[i, i, 0, 0] = 1; I3[i, i, 0] = 1
xs = (np.arange(2 * 4096) % 1021).astype(np.float32).reshape(2, 4096)
ys, _ = run(xs, I1, I2, I3, None, False)
bad = np.argwhere(ys != xs)
print("identity cores: mismatches %d of %d" % (len(bad), xs.size))
if len(bad):
    for s, f in bad[:12]:
        print("   sample %d out (o1,o2,o3)=(%d,%d,%d): got %g expected %g" % (s, f // 256, (f // 16) % 16, f % 16, ys[s, f], xs[s, f]))

for batch in (1, 2, 3, 75, 300, 4096):
    x = rng.normal(size=(batch, 4096)).astype(np.float32)
    bias = rng.normal(size=4096).astype(np.float32) * 0.01
    y, _ = run(x, G1, G2, G3, bias, True)
    r = ref(x, G1, G2, G3, bias, True)
    lin, _ = run(x, G1, G2, G3, None, False)
    rl = ref(x, G1, G2, G3, None, False)
    print("batch %5d: relu+bias max err %.3e (scale %.3e)   linear max err %.3e (scale %.3e)" % (
        batch, np.max(np.abs(y - r)), np.max(np.abs(r)), np.max(np.abs(lin - rl)), np.max(np.abs(rl))))

# timing at BASELINE configs[4]
dev = torch.device("cuda")
B = 65536
xb = torch.randn((B, 4096), dtype=torch.float32, device=dev)
packed = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in (G1, G2, G3)])
bias_d = torch.zeros(4096, dtype=torch.float32, device=dev)
out = torch.empty_like(xb)
for _ in range(3):
    ops.tt_dense3_tf32(xb, packed, bias_d, relu=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.tt_dense3_tf32(xb, packed, bias_d, relu=True, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("batch %d: %.3f ms = %.2f M samples/s, %.1f TFLOP/s (3.78e7 flop/sample), %.0f GB/s of x + y" % (
    B, ms, B / ms / 1e3, 3.78e7 * B / ms / 1e9, 2 * B * 4096 * 4 / ms / 1e6))
