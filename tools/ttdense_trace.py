"""Timeline of CTA 0 of the TF32 layer kernel (library built with SYN_NVCC_EXTRA=-DSYN_TT_DEBUG): per-phase waits and durations in cycles."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import defaultdict
from syngular_b200 import ops
rng = np.random.default_rng(0)
dev = torch.device("cuda")
G = [torch.from_numpy(rng.normal(scale=0.05, size=s).astype(np.float32)).to(dev) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
packed = ops.tt_dense3_pack(*G)
B = 74 * 6
x = torch.randn((B, 4096), dtype=torch.float32, device=dev)
NW, LEN = 12, 8192
buf = torch.zeros((NW, LEN, 2), dtype=torch.int32, device=dev)
ops.tt_dense3_tf32(x, packed, None, relu=True)           # warm
torch.cuda.synchronize()
ops.lib.syn_tt_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.tt_dense3_tf32(x, packed, None, relu=True)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
names = {0: "tma+s1 t0", 1: "s23 t0", 2: "epi g0 q2", 6: "epi g1 q2", 10: "s1 t1", 11: "s23 t1"}
for w in (0, 1, 2, 6, 11):
    ev = [(int(v), int(c)) for v, c in t[w] if v != 0]
    if not ev:
        continue
    print("== warp %d (%s): %d events, span %d cycles" % (w, names[w], len(ev), (ev[-1][1] - ev[0][1]) & 0xFFFFFFFF))
    # durations between consecutive marks, grouped by (tag of previous, tag of next); sample j >= 2 only (steady state)
    agg = defaultdict(list)
    for (v0, c0), (v1, c1) in zip(ev[:-1], ev[1:]):
        j0 = (v0 >> 8) & 0xF
        if j0 < 2:
            continue
        agg[(v0 >> 12, v1 >> 12)].append((c1 - c0) & 0xFFFFFFFF)
    for k in sorted(agg):
        a = np.array(agg[k])
        print("   %3x -> %3x : n %4d  mean %7.0f  median %7.0f  max %7d   total/sample %8.0f" % (k[0], k[1], len(a), a.mean(), np.median(a), a.max(), a.sum() / 4.0))
