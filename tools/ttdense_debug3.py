"""Progress-marker run of the TF32 kernel (library built with SYN_NVCC_EXTRA=-DSYN_TT_DEBUG): launch, sleep, dump where every warp is."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
from cuda import cudart
rng = np.random.default_rng(0)
G = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in ((16,16,16),(16,16,16,16),(16,16,16))]
dev = torch.device("cuda")
packed = ops.tt_dense3_pack(*[torch.from_numpy(g).to(dev) for g in G])
batch = int(sys.argv[1])
nwords = 148 * 8 * 4
err, hptr = cudart.cudaHostAlloc(nwords * 4, cudart.cudaHostAllocMapped)
err, dptr = cudart.cudaHostGetDevicePointer(hptr, 0)
host = (ctypes.c_uint32 * nwords).from_address(hptr)
for i in range(nwords): host[i] = 0
ops.lib.syn_tt_debug_buffer(ctypes.c_void_p(dptr))
x = torch.from_numpy(rng.normal(size=(batch, 4096)).astype(np.float32)).to(dev)
torch.cuda.synchronize()
y = ops.tt_dense3_tf32(x, packed, None, relu=False)
time.sleep(4)
arr = np.frombuffer(host, dtype=np.uint32).reshape(148, 8, 4)
grid = int(os.environ.get("SYN_TT_GRID", "148"))
for b in range(min(grid, 8)):
    print("block %d:" % b, " | ".join("w%d %x %x %x" % (w, arr[b, w, 0], arr[b, w, 1], arr[b, w, 2]) for w in range(6)))
sys.stdout.flush()
os._exit(0)
