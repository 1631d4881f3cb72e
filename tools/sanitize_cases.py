"""Small invocations of every kernel that synchronises by hand (grid barriers, version words, DSMEM exchanges, mbarrier pipelines),
for compute-sanitizer (tools/sanitize.sh): purify / orthonormalise, Jacobi (single- and multi-CTA), cluster QR, Householder QR,
Cholesky, fused overlap, environment sandwich, strided GEMM (TMA and cp.async variants), the TF32 tcgen05 layer kernel and its CTA-pair
variant (cluster barriers, remote mbarrier arrivals), the batched projection kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from syngular_b200 import ops
from syngular.tensor import _sweeps as sw
import bench

which = sys.argv[1:] or ["gemm", "purify", "ortho", "jacobi", "qr", "chol", "overlap", "sweep", "ttdense", "ttpair", "purifyb", "c128", "smallcore"]
dev = torch.device("cuda")
rng = np.random.default_rng(0)


def spd(n, gap_at):
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    lam = np.concatenate([np.linspace(1.0, 0.5, gap_at), np.linspace(0.05, 0.001, n - gap_at)])
    return torch.from_numpy((q * lam) @ q.T).to(dev)


if "gemm" in which:
    a = torch.randn(256, 128, dtype=torch.float64, device=dev); b = torch.randn(128, 384, dtype=torch.float64, device=dev)
    assert torch.allclose(ops.matmul(a, b), a @ b)
    a = torch.randn(70, 33, dtype=torch.float64, device=dev); b = torch.randn(33, 45, dtype=torch.float64, device=dev)
    assert torch.allclose(ops.matmul(a, b), a @ b)
    # block-lower output mask (128 x 64 tiles), a skinny and a flat output
    a = torch.randn(256, 96, dtype=torch.float64, device=dev); b = torch.randn(96, 128, dtype=torch.float64, device=dev)
    c = torch.zeros(256, 128, dtype=torch.float64, device=dev)
    ops.gemm(a, b, c, M=256, N=128, K=96, a_m=96, a_k=1, b_k=128, b_n=1, c_m=128, c_n=1, mask=(128, 64))
    ref = a @ b
    assert torch.allclose(c[:128, :64], ref[:128, :64]) and torch.allclose(c[128:], ref[128:]) and float(c[:128, 64:].abs().max()) == 0.0
    a = torch.randn(512, 40, dtype=torch.float64, device=dev); b = torch.randn(40, 16, dtype=torch.float64, device=dev)
    assert torch.allclose(ops.matmul(a, b), a @ b)
    assert torch.allclose(ops.matmul(b.t().contiguous(), a.t().contiguous()), b.t() @ a.t())
if "purify" in which:
    A = spd(128, 64)
    U, info = ops.dominant_subspace(A, 64, 52, 26, sp2_max=90, ns_max=60)
    assert float((U.t() @ U - torch.eye(64, dtype=torch.float64, device=dev)).abs().max()) < 1e-10
if "ortho" in which:
    L = torch.randn(256, 64, dtype=torch.float64, device=dev)
    Q, info = ops.orthonormalize_columns(L)
    assert float((Q.t() @ Q - torch.eye(64, dtype=torch.float64, device=dev)).abs().max()) < 1e-10
if "jacobi" in which:
    for n in (48, 160):
        G = spd(n, n // 2).contiguous()
        Ut, sigma, info, winfo = ops.jacobi_solve(G.clone(), n, sqrt_mode=True)
        assert int(info[1]) == n
if "qr" in which:
    for m, n, q in ((96, 40, 24), (512, 128, 64)):
        L = torch.randn(m, n, dtype=torch.float64, device=dev)
        Q, S = ops.qrt(L, q)
        assert float((Q.t() @ Q - torch.eye(Q.shape[1], dtype=torch.float64, device=dev)).abs().max()) < 1e-10
if "chol" in which:
    G = spd(192, 96).contiguous()
    B, shift = ops.chol_upper(G.clone())
if "overlap" in which:
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    bonds = bench.capped_bonds(12, 2, 32)[1:-1]
    A = BMPS.random(5, (2,) * 12, bonds, seed=1); B = BMPS.random(5, (2,) * 12, bonds, seed=2)
    A.overlap(B)
if "sweep" in which:
    X, W = bench.make_chain(3, n=14, chi=64, chiw=16)
    Xd = [sw.as_core(x) for x in X]; Wd = [sw.as_core(w) for w in W]
    sw.apply_round_dm(Xd, Wd, 64)
    sw.apply_round_qr(Xd, Wd, 64)
if "ttdense" in which:
    from syngular.layers import TensorDense
    layer = TensorDense((16, 16, 16), (16, 16, 16), (16, 16), seed=1, precision="tf32").build()
    layer(np.random.default_rng(1).normal(size=(5, 4096)).astype(np.float32))
if "ttpair" in which:
    G = [torch.from_numpy(rng.normal(scale=0.05, size=sh).astype(np.float32)).to(dev) for sh in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
    packed = ops.tt_dense3_pack(*G)
    x = torch.randn((5, 4096), dtype=torch.float32, device=dev)
    y1 = ops.tt_dense3_tf32(x, packed, None, relu=True, pair=True)
    y0 = ops.tt_dense3_tf32(x, packed, None, relu=True)
    assert float((y1 - y0).abs().max()) < 1e-5
if "purifyb" in which:
    A = torch.stack([spd(64, 32) for _ in range(3)]).contiguous()
    U, info = ops.dominant_subspace_batched(A, 32)
    assert float((U[1].t() @ U[1] - torch.eye(32, dtype=torch.float64, device=dev)).abs().max()) < 1e-10
if "c128" in which:
    Z = rng.normal(size=(64, 64)) + 1j * rng.normal(size=(64, 64))
    Qc, _ = np.linalg.qr(Z)
    lam = np.concatenate([np.linspace(1.0, 0.5, 32), np.linspace(0.05, 0.001, 32)])
    H = (Qc * lam) @ Qc.conj().T
    H = 0.5 * (H + H.conj().T)
    Ure, Uim, info = ops.dominant_subspace_c128(torch.from_numpy(np.ascontiguousarray(H.real)).to(dev), torch.from_numpy(np.ascontiguousarray(H.imag)).to(dev), 32)
    U = torch.complex(Ure, Uim)
    assert float((U.conj().t() @ U - torch.eye(32, dtype=torch.complex128, device=dev)).abs().max()) < 1e-10
if "smallcore" in which:
    X = torch.randn((9, 8, 70), dtype=torch.float64, device=dev); Wm = torch.randn((8, 8), dtype=torch.float64, device=dev)
    Y = torch.empty((9, 8, 70), dtype=torch.float64, device=dev)
    ops.apply_small_core(X, Wm, Y, Q=9, L=70, x_q=560, x_r=70, x_l=1, y_q=560, y_ro=(0, 70, 8), y_l=1)
    assert torch.allclose(Y, torch.einsum("or,qrx->qox", Wm, X))
torch.cuda.synchronize()
print("sanitize cases ok:", " ".join(which))
