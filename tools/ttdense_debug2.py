import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
rng = np.random.default_rng(0)
G = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in ((16,16,16),(16,16,16,16),(16,16,16))]
dev = torch.device("cuda")
packed = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in G])
batch = int(sys.argv[1])
x = rng.normal(size=(batch, 4096)).astype(np.float32)
y = ops.tt_dense3_tf32(torch.from_numpy(x).to(dev), packed, None, relu=False)
torch.cuda.synchronize()
r = np.einsum("sijk,iax,jbxy,kcy->sabc", x.reshape(-1,16,16,16).astype(np.float64), *[g.astype(np.float64) for g in G], optimize=True).reshape(batch, -1)
err = np.abs(y.cpu().numpy() - r).max(axis=1)
print("batch", batch, "grid", os.environ.get("SYN_TT_GRID"), "max err", err.max(), "worst samples", np.argsort(-err)[:6], "scale", np.abs(r).max())
