"""Timeline of the first CTA pair of the cta_group::2 TF32 layer kernel (library built with SYN_NVCC_EXTRA=-DSYN_TT_DEBUG, ops.tt_dense3_tf32(pair=True)): for each traced warp, the time between consecutive marks grouped by (tag -> tag), steady-state samples only."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from collections import defaultdict
from syngular_b200 import ops
rng = np.random.default_rng(0)
dev = torch.device("cuda")
G = [torch.from_numpy(rng.normal(scale=0.05, size=s).astype(np.float32)).to(dev) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
packed = ops.tt_dense3_pack(*G)
NS = 8
B = 74 * NS
x = torch.randn((B, 4096), dtype=torch.float32, device=dev)
NW, LEN = 12, 4096
buf = torch.zeros((2, NW, LEN, 2), dtype=torch.int32, device=dev)
ops.tt_dense3_tf32(x, packed, None, relu=True, pair=True)
torch.cuda.synchronize()
ops.lib.syn_ttp_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.tt_dense3_tf32(x, packed, None, relu=True, pair=True)
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
for cta, w in ((0, 0), (0, 1), (0, 2), (0, 6), (1, 2), (1, 6)):
    ev = [(int(v), int(c)) for v, c in t[cta, w] if v != 0]
    if not ev:
        continue
    span = (ev[-1][1] - ev[0][1]) & 0xFFFFFFFF
    print("== cta %d warp %d: %d events, span %d cycles (%.0f per sample)" % (cta, w, len(ev), span, span / NS))
    agg = defaultdict(list)
    for (v0, c0), (v1, c1) in zip(ev[:-1], ev[1:]):
        j0 = (v0 >> 8) & 0xF
        if j0 < 2 or j0 >= NS - 1:
            continue
        agg[(v0 >> 12, v1 >> 12)].append((c1 - c0) & 0xFFFFFFFF)
    for k in sorted(agg):
        a = np.array(agg[k])
        print("   %3x -> %3x : n %4d  mean %7.0f  median %7.0f  max %7d   total/sample %8.0f" % (k[0], k[1], len(a), a.mean(), np.median(a), a.max(), a.sum() / (NS - 3.0)))
