import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from syngular_b200 import ops
n = 512
rng = np.random.default_rng(0)
B = torch.from_numpy(rng.normal(size=(n, 4 * n))).cuda()
A = B @ B.T
def T(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best, r
t, G32 = T(lambda: ops.cast_f32(A)); print("cast %.3f ms" % t)
def j32():
    g = G32.clone(); ops.jacobi_rows_f32(g); return g
t, g = T(j32); words = ops._workspaces[("jacobi_ctrl32", A.device)].view(torch.int32); stride = ops.lib.syn_jacobi_ctrl_stride(30)
print("jacobi32 %.3f ms (incl clone), sweeps %d" % (t, int(words[30 + 1].item())))
t, U = T(lambda: ops.rows_to_basis(g)); print("basis %.3f ms, orth err %.2e" % (t, (U @ U.T - torch.eye(n, device=U.device, dtype=U.dtype)).abs().max().item()))
def ns(U):
    for _ in range(2):
        X = ops.matmul(U, U.t()); Un = ops.copy_strided(U); ops.matmul(X, U, out=Un, alpha=-0.5, beta=1.5); U = Un
    return U
t, U2 = T(lambda: ns(U)); print("newton-schulz x2 %.3f ms, orth err %.2e" % (t, ops.identity_deviation(ops.matmul(U2, U2.t())).item()))
t, Ap = T(lambda: ops.matmul(ops.matmul(U2, A), U2.t())); off = (Ap - torch.diag(torch.diagonal(Ap))).abs().max().item() / Ap.abs().max().item()
print("transform %.3f ms, offdiag %.2e" % (t, off))
def j64():
    w = Ap.clone(); ops.jacobi_rows(w, null_rel=1e-13); return w
t, w = T(j64); print("jacobi64 on A' %.3f ms, sweeps %s" % (t, ops.jacobi_sweeps_used()))
def j64full():
    w = A.clone(); ops.jacobi_rows(w, null_rel=1e-13); return w
t, w = T(j64full); print("jacobi64 on A  %.3f ms, sweeps %s" % (t, ops.jacobi_sweeps_used()))
